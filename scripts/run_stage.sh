mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -k "unsupported or many" > gpurun_out/pytest_one.log 2>&1; tail -30 gpurun_out/pytest_one.log
