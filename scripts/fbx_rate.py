#!/usr/bin/env python
"""Kernel rates of the Forward/Backward parsers — the fast kernels (fwd_mw/bwd_mw) and the exact-order kernels (fbx.cu) — on a
slice of iid targets, CUDA-event timed through p7gpu_ctx_set_timing:   python scripts/fbx_rate.py [--n 20000] [--models M400,M2000]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000); ap.add_argument("--mean-len", type=int, default=350)
    ap.add_argument("--models", default="synth_M100,synth_M400,synth_M800,synth_M2000")
    ap.add_argument("--groups", default="0"); ap.add_argument("--warps", type=int, default=0)
    args = ap.parse_args()
    from biobench2_b200 import engine, seqdb
    from biobench2_b200.oprofile import OProfile
    ctx = engine.Context(0)
    hdb = seqdb.synth_db(args.n, mean=args.mean_len, seed=5)
    db = engine.SeqDB.upload(ctx, hdb)
    out = {}
    for name in args.models.split(","):
        op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % name))
        prof = engine.Profile(ctx, op)
        rec = {}
        for mode, exact, grp in [("fast", False, 0)] + [("exact_g%s" % g, True, int(g)) for g in args.groups.split(",")]:
            if grp * op.M * 12 > 220 * 1024:
                continue
            ctx.set_option("fbx_groups", grp); ctx.set_option("fbx_warps", args.warps)
            engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=False, region_trigger=True, fb_exact=exact).close()
            ctx.set_timing(True); ctx.reset_timing()
            r = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=False, region_trigger=True, fb_exact=exact)
            kf = ctx.kernel_time(engine.KERNEL_FWDX if exact else engine.KERNEL_FWD)[0]
            kb = ctx.kernel_time(engine.KERNEL_BWDX if exact else engine.KERNEL_BWD)[0]
            ctx.set_timing(False)
            cells = op.M * float(hdb.nres)
            rec[mode] = {"fwd_gcups": cells / (kf * 1e-3) / 1e9, "bwd_gcups": cells / (kb * 1e-3) / 1e9, "fwd_ms": kf, "bwd_ms": kb}
            r.close()
        out[name] = rec
        prof.close()
    print(json.dumps({"n": args.n, "mean_len": args.mean_len, "rates": out}))


if __name__ == "__main__":
    main()
