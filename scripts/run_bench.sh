mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_full4.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full4.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['cpu_baseline']); print(json.dumps(d['pipeline'], indent=1))"
tail -5 gpurun_out/bench_full4.log | head -4 | cut -c1-300
