#!/usr/bin/env python
"""The fast Forward/Backward CTA kernels alone (W = 4, 2, 1) on a handful of short targets: small enough for racecheck.
    compute-sanitizer --tool racecheck python scripts/sanitize_fb.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile

ctx = engine.Context(0)
hdb = seqdb.concat([seqdb.synth_db(10, mean=60, seed=3, degenerate_frac=0.02), seqdb.synth_db(1, fixed_length=150, seed=4)])
for name in ("synth_M2000", "synth_M800", "synth_M400"):
    op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % name))
    prof = engine.Profile(ctx, op); db = engine.SeqDB.upload(ctx, hdb)
    fsc, fst, xmx, xoff, tot = engine.forward_batch(ctx, prof, db, hdb.lengths)
    bsc, bst, bx, btot = engine.backward_batch(ctx, prof, db, xmx, xoff)
    print(name, "ok", float(np.nanmax(fsc)), float(np.nanmax(np.abs(fsc - bsc))), flush=True)
    db.close(); prof.close()
ctx.close()
