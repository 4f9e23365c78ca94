# block prefilter of hmmscan_b200: parity tests, then wall clock with and without it (2000-model pressed database built in the
# authoring container under tests/_refdata/scan, P7SCAN_DB)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "scan_msv_prefilter" -x > gpurun_out/r7_scan_tests.log 2>&1
timeout 900 python -m pytest tests/test_integration_hmmscan.py -q -m gpu -x >> gpurun_out/r7_scan_tests.log 2>&1
tail -5 gpurun_out/r7_scan_tests.log
fi
export P7SCAN_DB=tests/_refdata/scan
OUT=gpurun_out/${TAG:-r7}_hmmscan_rate.jsonl
: > $OUT
for nq in ${NQS:-1 10 50}; do
  timeout 300 python scripts/hmmscan_rate.py 2000 $nq >> $OUT 2>> gpurun_out/r7_scan_rate.err
  P7SCAN_B200_OPTS="--prefilter 0" timeout 300 python scripts/hmmscan_rate.py 2000 $nq >> $OUT 2>> gpurun_out/r7_scan_rate.err
done
cat $OUT
