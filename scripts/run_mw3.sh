mkdir -p gpurun_out
for mm in 513 1; do for p in synth_M400 Pkinase globins4 synth_M100 synth_M33; do echo "MINM=$mm $p"; P7GPU_MW_MINM=$mm timeout 300 python scripts/stage_bench.py $p 100000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('fwd_gcups','bwd_gcups','fwd_ms','bwd_ms')})"; done; done 2>&1 | tee gpurun_out/stage_mw3.log
P7GPU_MW_MINM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longmodel.py -x -q -k "forward or backward or cta or wrappers" 2>&1 | tail -3
