mkdir -p gpurun_out
T=/tmp/p7prof; mkdir -p $T
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fwd_mw_kernel|bwd_mw_kernel" -c 2 -o $T/prof_mw -f python scripts/stage_bench.py synth_M2000 6000 > gpurun_out/ncu_mw.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_mw.ncu-rep > gpurun_out/mw_kernels_M2000_full.txt
ncu -i $T/prof_mw.ncu-rep --page source --csv --kernel-name regex:fwd_mw_kernel > gpurun_out/mw_fwd_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
timeout 1200 python -m pytest tests/test_gpu_longmodel.py -x -q > gpurun_out/pytest_mw.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mw.log
