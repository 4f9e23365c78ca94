mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vit_kernel|fwd_kernel|bwd_kernel" -c 3 -o gpurun_out/prof_stages_r1 -f python scripts/stage_bench.py synth_M400 30000 > gpurun_out/ncu_stage.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_stage.log
