"""biobench2_b200 — B200-native engine for the HMMER3 hmmsearch acceleration pipeline.

The product is libp7gpu.so (biobench2_b200/csrc, C ABI in include/p7gpu.h): hand-written sm_100a
kernels for the SSV/MSV filter, the Viterbi filter and the Forward/Backward parsers of
p7_Pipeline() (reiverjohn/biobench2 HMMER/src/p7_pipeline.c:665-901).  This Python package is only
the thin host-side binding used by tests and bench.py; it has no CPU fallback.
"""
__version__ = "0.1.0"
