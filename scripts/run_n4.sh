mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-stages --e2e-steps 3 > gpurun_out/bench_n4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n4.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
