mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_region_trigger.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pipe.log
P7GPU_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --e2e-steps 3 --no-stages > gpurun_out/bench_trace.log 2>&1; grep "p7gpu" gpurun_out/bench_trace.log | tail -9
tail -1 gpurun_out/bench_trace.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
