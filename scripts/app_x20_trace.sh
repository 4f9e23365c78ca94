# where the first query's "GPU pipeline" time goes on 20 copies of Swiss-Prot (P7GPU_TRACE checkpoints), five runs back to back
python - <<'PY'
body = open('tests/_refdata/uniprot_sprot.fasta','rb').read()
with open('/tmp/sprot_x20.fasta','wb') as f:
    for r in range(20):
        f.write(body.replace(b">", b">r%02d_" % r) if r else body)
import os; os.sync()
PY
H=tests/_refdata/tufa420.hmm; F=/tmp/sprot_x20.fasta
for i in 1 2 3; do sleep ${1:-0.5}; echo "-- run $i"; P7GPU_TRACE=1 integration/_build/hmmsearch_b200 --times -o /dev/null --cpu 16 $H $F 2>&1 | grep -E "p7gpu|# query|# 1 GPU" | tail -24; done
