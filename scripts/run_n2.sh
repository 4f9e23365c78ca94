mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_n2.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_ref_n2.log | cut -c1-600
