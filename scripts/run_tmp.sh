mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-stages --e2e-steps 3 > gpurun_out/bench_al.log 2>&1; tail -1 gpurun_out/bench_al.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline_alu']['kernel_gcups'], d['roofline_alu']['frac'], d['detail']['residues_per_gpu'])"
