mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_s4.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_s4.log | cut -c1-600
