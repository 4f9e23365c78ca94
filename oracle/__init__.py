"""Test oracle for the HMMER3 p7_Pipeline() hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under biobench2_b200/ imports it (tests/test_layout.py checks that).

  oracle.port : ctypes binding of liboracle.so — our plain-C k-space restatement (oracle/p7_oracle.c)
  oracle.ref  : ctypes binding of _ref/libp7ref.so — the UNMODIFIED reference HMMER 3.1b2 compiled
                from /root/reference by oracle/Makefile (`make -C oracle ref`), plus ref_harness.c
"""
