# whole-application timing on 20 concatenated copies of Swiss-Prot (names made unique), with the engine's phase trace
set -e
python - <<'PY'
body = open('tests/_refdata/uniprot_sprot.fasta','rb').read()
with open('/tmp/sprot_x20.fasta','wb') as f:
    for r in range(20):
        f.write(body.replace(b">", b">r%02d_" % r) if r else body)
PY
H=tests/_refdata/tufa420.hmm; F=/tmp/sprot_x20.fasta
for v in "" "--pinfirst"; do
  for i in 1 2; do ( time integration/_build/hmmsearch_b200 $v --times -o /tmp/b200.out --tblout /tmp/b200.tbl --cpu 16 $H $F ) 2>&1 | grep -E "^#|real"; done
done
( time oracle/_ref/hmmsearch -o /tmp/ref.out --tblout /tmp/ref.tbl --cpu 16 $H $F ) 2>&1 | grep real
diff <(grep -v "^#" /tmp/ref.tbl) <(grep -v "^#" /tmp/b200.tbl) > /dev/null && echo "tblout identical"
P7GPU_TRACE=1 integration/_build/hmmsearch_b200 --times -o /dev/null --cpu 16 $H $F 2>&1 | grep -E "p7gpu|^#" | head -60
