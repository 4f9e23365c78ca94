mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_ssv.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ssv.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_ssv.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_ssv.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline_alu']['kernel_gcups']); p=d['pipeline']; print(p['ms'], p['kernel_ms'])"
T=/tmp/p7prof; mkdir -p $T
timeout 900 ncu --set full --clock-control none -k regex:ssv_kernel -s 1 -c 1 -o $T/prof_ssv -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-stages > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu full rc=$?"
python profiles/summarize.py kernel $T/prof_ssv.ncu-rep > gpurun_out/r1h_ssv_kernel_full.txt; grep -E "dram__bytes|time_duration|t_sector_hit|registers" gpurun_out/r1h_ssv_kernel_full.txt
