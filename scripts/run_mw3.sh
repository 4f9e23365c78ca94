mkdir -p gpurun_out
for mm in 513 1; do for p in synth_M400 Pkinase synth_M100; do echo "MINM=$mm $p"; P7GPU_MW_MINM=$mm timeout 300 python scripts/stage_bench.py $p 60000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('fwd_gcups','bwd_gcups','fwd_ms','bwd_ms','ssv_gcups','vit_gcups')})"; done; done 2>&1 | tee gpurun_out/stage_mw3.log
