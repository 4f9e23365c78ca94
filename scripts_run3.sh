mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full2.log
python - <<'P'
import torch,time
x=torch.empty(1<<30,dtype=torch.uint8,pin_memory=True); d=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
for i in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x,non_blocking=True); torch.cuda.synchronize(); print("H2D pinned GB/s", 1.0737/(time.perf_counter()-t))
for i in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); x.copy_(d,non_blocking=True); torch.cuda.synchronize(); print("D2H pinned GB/s", 1.0737/(time.perf_counter()-t))
P
