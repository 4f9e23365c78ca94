# ncu evidence for the round: launch list of the default bench command + one full capture per kernel family.
# Summaries (text) are written to gpurun_out/; copy the ones to keep into profiles/.
R=${1:-r1h}
mkdir -p gpurun_out
T=/tmp/p7prof; mkdir -p $T
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list rc=$?"
python profiles/summarize.py launches gpurun_out/launches_$R.csv > gpurun_out/${R}_launches_bench.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_mw_kernel|bwd_mw_kernel|region_trigger_kernel|bias_kernel" -c 6 -o $T/prof_stages -f python scripts/stage_bench.py synth_M400 30000 > gpurun_out/ncu_stage.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages.ncu-rep > gpurun_out/${R}_stage_kernels_M400_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_mw_kernel|bwd_mw_kernel" -c 3 -o $T/prof_stages2000 -f python scripts/stage_bench.py synth_M2000 6000 > gpurun_out/ncu_stage2000.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages2000.ncu-rep > gpurun_out/${R}_stage_kernels_M2000_full.txt
ls -la gpurun_out | tail -8
