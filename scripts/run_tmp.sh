mkdir -p gpurun_out
P7GPU_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --e2e-steps 3 --no-stages > gpurun_out/bench_trace.log 2>&1; grep "p7gpu" gpurun_out/bench_trace.log | tail -9
tail -1 gpurun_out/bench_trace.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'])"
