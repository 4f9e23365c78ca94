timeout 1200 python -m pytest tests/test_gpu_pipeline.py -x -q -k "device_stage and tufa" 2>&1 | grep -E "^E|assert" | head -20
