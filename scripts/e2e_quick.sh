# the streamed MSV pass (bench.py e2e leg) alone: resident ms, e2e ms, SSV launch sum of one streamed step
Q="--steps 3 --warmup 2 --e2e-steps 4 --no-cpu-baseline --no-stages --strong-total 0 --config5-passes 0 --no-app"
for i in 1 2; do timeout 300 python bench.py $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['e2e']['kernel_ms_one_step']['ssv_sum_over_chunk_launches'], d['e2e']['counters'])"; done
