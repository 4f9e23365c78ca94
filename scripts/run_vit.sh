mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_longmodel.py -x -q 2>&1 | tail -2
for p in synth_M800 synth_M2000; do timeout 300 python scripts/stage_bench.py $p 100000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['profile'], {k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('vit_gcups','vit_exact32_gcups','fwd_gcups','bwd_gcups')})"; done 2>&1 | tee gpurun_out/stage_vit.log
