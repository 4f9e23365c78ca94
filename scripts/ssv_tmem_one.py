import sys
sys.path.insert(0, '/root/repo')
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile
ctx = engine.Context(0)
hdb = seqdb.synth_db(600000, mean=350, seed=77)
db = engine.SeqDB.upload(ctx, hdb)
op = OProfile.load('/root/repo/tests/golden/prof_synth_M400.npz')
prof = engine.Profile(ctx, op)
ctx.set_option("ssv_tmem", int(sys.argv[1]))
for rep in range(2):
    engine.ssvfilter_batch(ctx, prof, db)
