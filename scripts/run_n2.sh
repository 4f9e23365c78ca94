# two GPUs: the two-GPU integration tests, the sharded-search check over NCCL, the bench line at N=2
R=${1:-r5}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or sharded or multi" --timeout 500 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/nccl_merge_check.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-app --config5-passes 24 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/${R}_bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['strong']['value'])"
