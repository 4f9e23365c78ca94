mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longmodel.py tests/test_gpu_pipeline.py tests/test_gpu_region_trigger.py -x -q 2>&1 | tail -3
for p in synth_M400 Pkinase synth_M100 synth_M2000; do timeout 300 python scripts/stage_bench.py $p 100000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['profile'], {k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('fwd_gcups','bwd_gcups')})"; done
python __graft_entry__.py smoke 2>&1 | tail -1
