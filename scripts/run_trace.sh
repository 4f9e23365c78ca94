mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
P7GPU_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 3 > gpurun_out/bench_trace.log 2>&1; grep p7gpu gpurun_out/bench_trace.log | tail -9; tail -1 gpurun_out/bench_trace.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
