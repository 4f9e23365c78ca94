import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile
ctx = engine.Context(0); ctx.set_timing(True)
hdb = seqdb.synth_db(2000000, mean=350, seed=77)
db = engine.SeqDB.upload(ctx, hdb)
for name in ("synth_M800", "synth_M400", "Pkinase", "synth_M100", "globins4"):
    op = OProfile.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'prof_%s.npz' % name))
    prof = engine.Profile(ctx, op)
    out = {"profile": name, "M": op.M}
    res = {}
    for tm in (0, 2):
        ctx.set_option("ssv_tmem", tm)
        for rep in range(2):
            ctx.reset_timing()
            sc, st = engine.ssvfilter_batch(ctx, prof, db)
            ms, n = ctx.kernel_time(engine.KERNEL_SSV)
        out["tmem%d_tcups" % min(tm, 1)] = round(op.M * hdb.nres / ms / 1e9, 2)
        res[tm] = (sc.copy(), st.copy())
    out["identical"] = bool(np.array_equal(res[0][0].view(np.uint32), res[2][0].view(np.uint32)) and np.array_equal(res[0][1], res[2][1]))
    print(json.dumps(out))
    prof.close()
