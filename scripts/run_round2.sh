# Round-2 evidence run on one B200 (gpurun): CLI wall clock, full GPU test suite, the bench line, ncu launch list and full captures.
# Text summaries land in gpurun_out/ (scratch); the ones to keep are copied into profiles/ by hand.
R=${1:-r2}
mkdir -p gpurun_out
T=/tmp/p7prof; mkdir -p $T
for i in 1 2 3; do ( time integration/_build/hmmsearch_b200 --times -o /dev/null --cpu 16 tests/_refdata/tufa420.hmm tests/_refdata/uniprot_sprot.fasta ) 2>&1 | grep -E "^#|real"; done > gpurun_out/${R}_cli_times.txt
for i in 1 2 3; do ( time oracle/_ref/hmmsearch -o /dev/null --cpu 16 tests/_refdata/tufa420.hmm tests/_refdata/uniprot_sprot.fasta ) 2>&1 | grep real; done >> gpurun_out/${R}_cli_times.txt
cat gpurun_out/${R}_cli_times.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${R}_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_ref_n1.json 2>> gpurun_out/${R}_bench_n1.err; echo "ref bench rc=$?"
Q="--steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --strong-total 0 --config5-passes 0 --no-app"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$R.csv python bench.py $Q > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list rc=$?"
python profiles/summarize.py launches gpurun_out/launches_$R.csv > gpurun_out/${R}_launches_bench.txt
timeout 900 ncu --set full --clock-control none -k regex:"ssv_tm_kernel|ssv_kernel" -s 1 -c 1 -o $T/prof_ssv -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-stages --strong-total 0 --no-app > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu ssv rc=$?"
python profiles/summarize.py kernel $T/prof_ssv.ncu-rep > gpurun_out/${R}_ssv_kernel_full.txt; grep -E "dram__bytes|time_duration|t_sector_hit" gpurun_out/${R}_ssv_kernel_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_mw_kernel|bwd_mw_kernel|region_trigger_kernel|bias_kernel" -c 6 -o $T/prof_stages -f python scripts/stage_bench.py synth_M400 30000 > gpurun_out/ncu_stage.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages.ncu-rep > gpurun_out/${R}_stage_kernels_M400_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"vitp_kernel|fwd_mw_kernel|bwd_mw_kernel" -c 3 -o $T/prof_stages2000 -f python scripts/stage_bench.py synth_M2000 6000 > gpurun_out/ncu_stage2000.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_stages2000.ncu-rep > gpurun_out/${R}_stage_kernels_M2000_full.txt
timeout 900 ncu --set full --clock-control none -k regex:"fbx_fwd_kernel|fbx_bwd_kernel|domain_decoding_kernel" -s 2 -c 2 -o $T/prof_fbx -f python scripts/fbx_rate.py --n 20000 --models synth_M400 > gpurun_out/ncu_fbx.log 2>&1; echo rc=$?
python profiles/summarize.py kernel $T/prof_fbx.ncu-rep > gpurun_out/${R}_fbx_kernels_M400_full.txt
ls -la gpurun_out | tail -12
