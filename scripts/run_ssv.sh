mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_ssv.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_ssv.log
for p in synth_M100 fn3 Pkinase synth_M400; do timeout 300 python scripts/stage_bench.py $p 200000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['profile'], d['M'], {k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('ssv_gcups','vit_gcups','fwd_gcups','bwd_gcups')})"; done 2>&1 | tee gpurun_out/stage_evenq.log
