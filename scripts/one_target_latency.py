#!/usr/bin/env python
"""Latency of the one-target drop-in calls (p7gpu_MSVFilter / ViterbiFilter / ForwardParser): what the reference's
calibration loops (evalues.c:264,337,442: 200 random targets each) pay per call.  Usage: one_target_latency.py [profile]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biobench2_b200 import engine
from biobench2_b200.oprofile import OProfile

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tufa420"
    op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % name))
    t0 = time.perf_counter(); ctx = engine.Context(0); t1 = time.perf_counter(); prof = engine.Profile(ctx, op); t2 = time.perf_counter()
    out = {"profile": name, "M": op.M, "ctx_create_ms": (t1 - t0) * 1e3, "profile_upload_ms": (t2 - t1) * 1e3}
    rng = np.random.default_rng(5)
    for L in (100, 200):
        seqs = []
        for _ in range(200):
            d = np.full(L + 2, 255, np.uint8); d[1:L + 1] = rng.integers(0, 20, L); seqs.append(d)
        for nm, fn in (("msv", engine.p7_MSVFilter), ("vit", engine.p7_ViterbiFilter), ("fwd", engine.p7_ForwardParser)):
            fn(ctx, seqs[0], L, prof)
            t = time.perf_counter()
            for d in seqs:
                fn(ctx, d, L, prof)
            out["%s_L%d_us_per_call" % (nm, L)] = (time.perf_counter() - t) / len(seqs) * 1e6
    print(json.dumps(out))

main()
