mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dbfile.py -x -q 2>&1 | tail -2
timeout 900 python scripts/dbfile_bench.py 2000000 2>&1 | tail -1 | tee gpurun_out/dbfile_bench.json
