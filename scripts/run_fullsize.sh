mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15
