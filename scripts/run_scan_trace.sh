# where a model that is uploaded spends its time in hmmscan_b200 (P7GPU_TRACE checkpoints of p7gpu_pipeline_run)
export P7SCAN_DB=tests/_refdata/scan
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from biobench2_b200 import seqdb
aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
qdb = seqdb.synth_db(10, mean=350, seed=5)
with open('/tmp/q10.fa', 'w') as f:
    for q in range(10):
        o, L = int(qdb.offsets[q]), int(qdb.lengths[q])
        f.write(">q%d\n%s\n" % (q, "".join(aa[qdb.residues[o:o + L]])))
PY
integration/_build/hmmscan_b200 --times -o /dev/null tests/_refdata/scan/scan2000.hmm /tmp/q10.fa > /dev/null 2>&1
P7GPU_TRACE=1 integration/_build/hmmscan_b200 --times -o /dev/null tests/_refdata/scan/scan2000.hmm /tmp/q10.fa 2> gpurun_out/r8_scan_trace.err
tail -60 gpurun_out/r8_scan_trace.err
