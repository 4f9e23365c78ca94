mkdir -p gpurun_out
for p in synth_M400 synth_M800 synth_M2000; do timeout 300 python scripts/stage_bench.py $p 200000 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['profile'], round(d['ssv_gcups'],1))"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "integer or geometry or seeded" 2>&1 | tail -2
