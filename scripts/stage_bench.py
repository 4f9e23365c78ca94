#!/usr/bin/env python
"""Per-stage kernel throughput (GCUPS) on a resident synthetic database: Viterbi, Forward, Backward, bias, MSV-J.
Kernel times are CUDA-event brackets around each launch (p7gpu_ctx_set_timing).  Usage: stage_bench.py [profile] [nseq]"""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "synth_M400"
    nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % name))
    hdb = seqdb.synth_db(nseq, mean=350, seed=77)
    ctx = engine.Context(0); prof = engine.Profile(ctx, op); db = engine.SeqDB.upload(ctx, hdb)
    cells = float(op.M) * hdb.nres
    out = {"profile": name, "M": op.M, "nseq": nseq, "nres": hdb.nres}
    ctx.set_timing(True)
    for rep in range(2):
        ctx.reset_timing()
        engine.vitfilter_batch(ctx, prof, db)
        ms, n = ctx.kernel_time(engine.KERNEL_VIT); out["vit_gcups"] = cells / ms / 1e6; out["vit_ms"] = ms
        out["vit_redo"] = int(ctx.last_vit_redo)
        ctx.set_option("vit_exact_only", 1); ctx.reset_timing()
        engine.vitfilter_batch(ctx, prof, db)
        ms, n = ctx.kernel_time(engine.KERNEL_VIT); out["vit_exact32_gcups"] = cells / ms / 1e6
        ctx.set_option("vit_exact_only", 0)
        ctx.reset_timing()
        engine.biasfilter_batch(ctx, prof, db)
        ms, n = ctx.kernel_time(engine.KERNEL_BIAS); out["bias_ms"] = ms; out["bias_gres_s"] = hdb.nres / ms / 1e6
        ctx.reset_timing()
        t0 = time.perf_counter()
        r = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=False)
        out["pipeline_max_wall_ms"] = (time.perf_counter() - t0) * 1e3
        for k, nm in ((engine.KERNEL_SSV, "ssv"), (engine.KERNEL_FWD, "fwd"), (engine.KERNEL_BWD, "bwd")):
            ms, n = ctx.kernel_time(k); out[nm + "_gcups"] = cells / ms / 1e6 if ms else None; out[nm + "_ms"] = ms
        out["counters_max"] = [int(x) for x in r.counters]
    print(json.dumps(out))

main()
