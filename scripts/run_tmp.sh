timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -k "max_mode" 2>&1 | tail -3
timeout 600 python - <<'PY' 2>&1 | tail -12
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from conftest import golden_set
from biobench2_b200 import engine, seqdb
ctx = engine.Context(0)
hdb = seqdb.synth_db(50000, mean=350, seed=5)
db = engine.SeqDB.upload(ctx, hdb)
for name in ("globins4", "fn3", "Pkinase", "tufa420_sprot", "synth_M100", "synth_M400", "synth_M800", "synth_M2000"):
    g = golden_set(name); prof = engine.Profile(ctx, g.oprofile)
    r = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=False, region_trigger=True)
    print(name, g.oprofile.M, "flagged fraction on random targets: %.4f" % float((r.survivors["region_flag"] == 1).mean()), flush=True)
    r.close(); prof.close()
PY
