mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "viterbi or vit or seeded or wrappers" > gpurun_out/pytest_vitp.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_vitp.log
for p in synth_M400 synth_M2000 Pkinase synth_M100; do timeout 300 python scripts/stage_bench.py $p 100000 2>&1 | tail -1; done | tee gpurun_out/stage_vitp.log
