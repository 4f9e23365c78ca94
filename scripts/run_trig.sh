mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_region_trigger.py tests/test_gpu_pipeline.py tests/test_gpu_longmodel.py -x -q > gpurun_out/pytest_trig.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_trig.log
