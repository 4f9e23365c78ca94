#!/usr/bin/env python
"""A few targets through every kernel family, small enough for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile

ctx = engine.Context(0)
hdb = seqdb.concat([seqdb.synth_db(48, mean=120, seed=3, degenerate_frac=0.02), seqdb.synth_db(2, fixed_length=700, seed=4)])
for name in ("synth_M2000", "synth_M800", "synth_M400", "synth_M100"):
    op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % name))
    prof = engine.Profile(ctx, op); db = engine.SeqDB.upload(ctx, hdb)
    engine.msvfilter_batch(ctx, prof, db); engine.vitfilter_batch(ctx, prof, db)
    fsc, fst, xmx, xoff, tot = engine.forward_batch(ctx, prof, db, hdb.lengths)
    bsc, bst, bx, btot = engine.backward_batch(ctx, prof, db, xmx, xoff)
    engine.region_trigger_batch(ctx, prof, db, xmx, bx, xoff)
    r = engine.pipeline_run(ctx, prof, db, do_max=True, region_trigger=True); r.close()
    r = engine.pipeline_run_host5(ctx, prof, engine.pack5(hdb.residues), hdb); r.close()
    r = engine.pipeline_run(ctx, prof, db, F1=0.5, F2=0.3, F3=0.2); r.close()
    r = engine.pipeline_run(ctx, prof, db, F1=0.5, F2=0.3, F3=0.2, fb_exact=True, region_trigger=True, decode=True); r.close()   # exact-order kernels + decoding
    ctx.set_option("ssv_tmem", 2); engine.ssvfilter_batch(ctx, prof, db)                 # Tensor Memory table wherever it fits (M <= 1024)
    ctx.set_option("ssv_tmem", 0); engine.ssvfilter_batch(ctx, prof, db); ctx.set_option("ssv_tmem", 1)
    print(name, "ok", float(np.nanmax(fsc)), flush=True)
    db.close(); prof.close()
ctx.close()
