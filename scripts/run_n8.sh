mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n8.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/config5_bench.py --gpus 8 --nseq 6250000 --passes 24 > gpurun_out/config5_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/config5_n8.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n4.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
