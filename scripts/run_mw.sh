mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_longmodel.py -x -q > gpurun_out/pytest_mw.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_mw.log
for cm in 0 8; do for p in synth_M2000 synth_M800; do echo "CMAX=$cm $p"; P7GPU_MW_CMAX=$cm timeout 300 python scripts/stage_bench.py $p 60000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('fwd_gcups','bwd_gcups','fwd_ms','bwd_ms','ssv_gcups','vit_gcups')})"; done; done 2>&1 | tee gpurun_out/stage_mw.log
