# full GPU test suite at HEAD, then the hmmscan rates again (one-block profile upload)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r8_pytest_gpu.log
export P7SCAN_DB=tests/_refdata/scan
OUT=gpurun_out/r8_hmmscan_rate.jsonl
: > $OUT
for nq in 1 10 25 100; do timeout 300 python scripts/hmmscan_rate.py 2000 $nq >> $OUT 2>> gpurun_out/r8_scan_rate.err; done
P7SCAN_B200_OPTS="--prefilter 0" timeout 300 python scripts/hmmscan_rate.py 2000 100 >> $OUT 2>> gpurun_out/r8_scan_rate.err
cat $OUT
