timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_region_trigger.py tests/test_dbfile.py -x -q 2>&1 | tail -3
