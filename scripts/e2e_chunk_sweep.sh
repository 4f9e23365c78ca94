Q="--steps 3 --warmup 2 --e2e-steps 4 --no-cpu-baseline --no-stages --strong-total 0 --config5-passes 0 --no-app"
run() { echo "== $1"; env $1 timeout 300 python bench.py $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'h2d', d['e2e']['h2d_bytes_per_step'])"; }
run "X=1"
run "P7GPU_STREAM_CHUNKMAX_MB=1024"
run "P7GPU_STREAM_CHUNKMAX_MB=1024 P7GPU_STREAM_GROWTH_PCT=100"
run "P7GPU_STREAM_CHUNK0_MB=4 P7GPU_STREAM_CHUNKMAX_MB=512 P7GPU_STREAM_GROWTH_PCT=60"
run "P7GPU_STREAM_CHUNKMAX_MB=64"
