# SSV/MSV parity tests and the resident MSV pass of the bench (value) alone
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "integer or filters or slot or ssv or msv or fullsize or homolog or long_target or idx" 2>&1 | tail -2
Q="--steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-stages --strong-total 0 --config5-passes 0 --no-app"
for i in 1 2; do timeout 300 python bench.py $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident ms', round(d['ms_per_step'],3), 'TCUPS', round(d['value']/1e3,3), 'roofline', round(d['roofline']['frac'],4))"; done
