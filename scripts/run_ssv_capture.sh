# ncu --set full of the dominant kernel at the bench's own size (source of profiles/ssv_traffic.json -> roofline.traffic)
R=${1:-r4}
mkdir -p gpurun_out; T=/tmp/p7prof; mkdir -p $T
timeout 900 ncu --set full --clock-control none -k regex:"ssv_tm_kernel|ssv_kernel" -s 1 -c 1 -o $T/prof_ssv -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-stages --strong-total 0 --config5-passes 0 --no-app > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu ssv rc=$?"
python profiles/summarize.py kernel $T/prof_ssv.ncu-rep > gpurun_out/${R}_ssv_kernel_full.txt; grep -E "^##|dram__bytes|time_duration|t_sector_hit" gpurun_out/${R}_ssv_kernel_full.txt
