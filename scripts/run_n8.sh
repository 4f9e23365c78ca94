mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "n=$n rc=$?"; tail -1 gpurun_out/bench_n$n.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], d['detail']['msv_survivors_all_ranks'])"
done
