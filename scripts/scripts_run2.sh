mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -1 gpurun_out/bench_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssv_kernel -s 1 -c 1 -o gpurun_out/prof_ssv_r1 -f python bench.py --nseq 2000000 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
