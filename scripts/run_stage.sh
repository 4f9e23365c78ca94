mkdir -p gpurun_out
for p in synth_M400 synth_M100 synth_M2000; do timeout 600 python scripts/stage_bench.py $p 100000 2>&1 | tail -2; done | tee gpurun_out/stage_bench.log
