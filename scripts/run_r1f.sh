mkdir -p gpurun_out
T=/tmp/p7prof; mkdir -p $T
# launch list of the default bench command (ncu serialises; shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list rc=$?"
# one full capture of the dominant kernel at the bench's own size (for roofline.traffic)
timeout 900 ncu --set full --clock-control none -k regex:ssv_kernel -s 1 -c 1 -o $T/prof_ssv_r1f -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-stages > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu full rc=$?"
python profiles/summarize.py kernel $T/prof_ssv_r1f.ncu-rep > gpurun_out/r1f_ssv_kernel_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_kernel|bwd_kernel" -c 3 -o $T/prof_stages_r1f -f python scripts/stage_bench.py synth_M400 30000 > gpurun_out/ncu_stage.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages_r1f.ncu-rep > gpurun_out/r1f_stage_kernels_M400_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_kernel|bwd_kernel" -c 3 -o $T/prof_stages_M2000_r1f -f python scripts/stage_bench.py synth_M2000 6000 > gpurun_out/ncu_stage2000.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages_M2000_r1f.ncu-rep > gpurun_out/r1f_stage_kernels_M2000_full.txt
ls -la gpurun_out
