# round-end state on one B200: the bench line, the reference arm, smoke, then the full GPU test suite (profiles/rN_*)
R=${1:-r5}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${R}_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_ref_n1.json 2>> gpurun_out/${R}_bench_n1.err; echo "ref bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
