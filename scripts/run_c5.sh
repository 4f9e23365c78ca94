mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -k "many" 2>&1 | tail -2
timeout 900 python scripts/config5_bench.py --nseq 6250000 --passes 24 2>&1 | tail -1 | tee gpurun_out/config5_n1.json
