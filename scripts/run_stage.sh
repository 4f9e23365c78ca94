mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
for p in synth_M400 synth_M2000 Pkinase synth_M800; do timeout 600 python scripts/stage_bench.py $p 100000 2>&1 | tail -1; done | tee gpurun_out/stage_bench.log
