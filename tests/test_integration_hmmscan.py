"""integration/impl_b200/hmmscan_b200: hmmscan (queries against a pressed profile database) on the B200 engine.

`hmmscan_b200` and the unmodified reference `oracle/_ref/hmmscan` scan the same query file against the same pressed database
(built at test time with the reference's own hmmpress from the committed tutorial models plus calibrated sampled models);
per query the pipeline statistics (models, nodes, the four filter counters, Z, domZ) must be equal and every non-comment
line of the main output, --tblout and --domtblout byte-identical.  The .h3f part of a model is read with the reference's
p7_oprofile_ReadMSV, the .h3p part (p7_oprofile_ReadRest) only for models some query got past the MSV filter of.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_integration_hmmsearch import payload, seeded_fasta, GOLDEN

SCAN = os.path.join(ROOT, "integration", "_build", "hmmscan_b200")
REFSCAN = os.path.join(ROOT, "oracle", "_ref", "hmmscan")
PRESS = os.path.join(ROOT, "oracle", "_ref", "hmmpress")

need_bins = pytest.mark.skipif(not (os.path.exists(SCAN) and os.path.exists(REFSCAN) and os.path.exists(PRESS)),
                               reason="integration/_build or oracle/_ref not built (needs the reference sources)")


def stats_blocks(path):
    """the per-query 'Internal pipeline statistics summary' numbers"""
    out = []
    with open(path) as f:
        for l in f:
            m = re.match(r"^(Query sequence\(s\)|Target model\(s\)|Passed MSV filter|Passed bias filter|Passed Vit filter|Passed Fwd filter|"
                         r"Initial search space \(Z\)|Domain search space  \(domZ\)):\s+(\d+)(?:\s+\((\d+) (?:residues|nodes))?", l)
            if m:
                out.append((m.group(1), int(m.group(2)), m.group(3)))
    return out


@need_bins
def test_scan_binary_links_the_engine():
    syms = subprocess.run(["nm", SCAN], capture_output=True, text=True).stdout
    assert re.search(r" T p7_MSVFilter$", syms, re.M) and re.search(r" T p7_MSVFilter_sse$", syms, re.M)
    assert re.search(r" T p7_oprofile_ReadMSV$", syms, re.M) and re.search(r" U p7gpu_pipeline_run$", syms, re.M)
    assert subprocess.run([SCAN, "-h"], capture_output=True, text=True).returncode == 0


def build_pressed_db(tmp_path, synth=((60, 11), (110, 12), (180, 13), (250, 14), (330, 15))):
    from oracle import ref
    db = tmp_path / "scan.hmm"
    with open(db, "wb") as f:
        for m in ("globins4", "fn3", "Pkinase"):
            f.write(open(os.path.join(GOLDEN, m + ".hmm"), "rb").read())
        for M, seed in synth:
            rp = ref.RefProfile.sample(M, seed, calibrate=True)
            one = tmp_path / ("s%d.hmm" % M)
            rp.write_hmm(str(one)); rp.close()
            f.write(open(one, "rb").read())
    p = subprocess.run([PRESS, str(db)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return db


def queries(tmp_path, n, seed):
    """n iid queries with homologs of the three tutorial models planted among them"""
    fa = tmp_path / "q.fa"
    parts = []
    for k, m in enumerate(("globins4", "fn3", "Pkinase")):
        part = tmp_path / ("q_%s.fa" % m)
        seeded_fasta(str(part), os.path.join(GOLDEN, m + ".hmm"), n // 3, seed + k, nhom=max(2, n // 40))
        txt = open(part, "rb").read().replace(b">seq", b">" + m.encode() + b"_seq").replace(b">hom", b">" + m.encode() + b"_hom")
        parts.append(txt)
    fa.write_bytes(b"".join(parts))
    return fa


@pytest.mark.gpu
@need_bins
@pytest.mark.parametrize("extra", [[], ["--max"], ["-E", "1e-3", "--domE", "1e-2", "--nobias"]])
def test_hmmscan_b200_identical_to_reference(tmp_path, extra):
    db = build_pressed_db(tmp_path)
    fa = queries(tmp_path, 150 if "--max" in extra else 420, seed=7)
    res = {}
    for name, exe in (("ref", REFSCAN), ("b200", SCAN)):
        o = tmp_path / (name + ".out"); t = tmp_path / (name + ".tbl"); d = tmp_path / (name + ".dom")
        p = subprocess.run([exe, "-o", str(o), "--tblout", str(t), "--domtblout", str(d)] + extra + [str(db), str(fa)],
                           capture_output=True, text=True, timeout=1800)
        assert p.returncode == 0, (name, p.stdout[-1500:], p.stderr[-1500:])
        res[name] = (str(o), str(t), str(d))
    sr, sb = stats_blocks(res["ref"][0]), stats_blocks(res["b200"][0])
    assert sr == sb and len(sr) >= 8 * 100, (len(sr), len(sb), [x for x in zip(sr, sb) if x[0] != x[1]][:5])
    assert sum(v for k, v, _ in sr if k == "Passed Fwd filter") >= 10
    for label, a, b in (("--tblout", res["ref"][1], res["b200"][1]), ("--domtblout", res["ref"][2], res["b200"][2]),
                        ("main output", res["ref"][0], res["b200"][0])):
        la, lb = payload(a), payload(b)
        if la != lb:
            diff = [(x, y) for x, y in zip(la, lb) if x != y]
            pytest.fail("hmmscan %s %s: %d of %d lines differ; first: %r" % (" ".join(extra), label, len(diff) + abs(len(la) - len(lb)), len(la), diff[:3]))


def run_both(tmp_path, db, fa, extra_b200, extra=()):
    res = {}
    err = ""
    for name, exe in (("ref", REFSCAN), ("b200", SCAN)):
        o = tmp_path / (name + ".out"); t = tmp_path / (name + ".tbl"); d = tmp_path / (name + ".dom")
        args = [exe, "-o", str(o), "--tblout", str(t), "--domtblout", str(d)] + list(extra) + (list(extra_b200) if name == "b200" else [])
        p = subprocess.run(args + [str(db), str(fa)], capture_output=True, text=True, timeout=1800)
        assert p.returncode == 0, (name, p.stdout[-1500:], p.stderr[-1500:])
        res[name] = (str(o), str(t), str(d))
        if name == "b200":
            err = p.stderr
    return res, err


def assert_identical(res, what):
    sr, sb = stats_blocks(res["ref"][0]), stats_blocks(res["b200"][0])
    assert sr == sb, (what, len(sr), len(sb), [x for x in zip(sr, sb) if x[0] != x[1]][:5])
    for label, a, b in (("--tblout", res["ref"][1], res["b200"][1]), ("--domtblout", res["ref"][2], res["b200"][2]),
                        ("main output", res["ref"][0], res["b200"][0])):
        la, lb = payload(a), payload(b)
        if la != lb:
            diff = [(x, y) for x, y in zip(la, lb) if x != y]
            pytest.fail("hmmscan %s %s: %d of %d lines differ; first: %r" % (what, label, len(diff) + abs(len(la) - len(lb)), len(la), diff[:3]))
    return sr


@pytest.mark.gpu
@need_bins
@pytest.mark.parametrize("nq,opts", [(6, []), (1, []), (45, ["--prefilter", "5"]), (45, ["--prefilter", "0"]), (420, ["--prefilter", "3"])])
def test_hmmscan_b200_block_prefilter_identical_to_reference(tmp_path, nq, opts):
    """Few queries against many models: the (model, query) block prefilter (p7gpu_scan_msv_prefilter) decides which models are
    uploaded at all.  Output must stay byte-identical to the reference's hmmscan, with block sizes that do and do not divide the
    number of models, and some models must actually end in the prefilter when the queries are few."""
    synth = tuple((M, 100 + M) for M in (40, 60, 85, 110, 140, 180, 215, 250, 290, 330, 400, 520))
    db = build_pressed_db(tmp_path, synth=synth)
    if nq >= 45:
        fa = queries(tmp_path, nq, seed=11)
    else:                                              # a planted homolog of one tutorial model + iid queries
        full = queries(tmp_path, 45, seed=13)
        recs = open(full, "rb").read().split(b">")[1:]
        hom = [r for r in recs if r.startswith(b"Pkinase_hom")][:1]
        iid = [r for r in recs if b"_seq" in r.split(b"\n", 1)[0]][:max(0, nq - 1)]
        fa = tmp_path / "few.fa"
        fa.write_bytes(b"".join(b">" + r for r in (iid + hom)))
    res, err = run_both(tmp_path, db, fa, ["--times"] + opts)
    sr = assert_identical(res, "nq=%d %s" % (nq, " ".join(opts)))
    m = re.search(r"(\d+) ended in the block prefilter \(blocks of (\d+)\)", err)
    assert m, err
    if nq <= 6:
        assert int(m.group(2)) == 512 and int(m.group(1)) >= 5, err         # most of the 15 models never reach the device on their own
        assert sum(v for k, v, _ in sr if k == "Passed Fwd filter") >= 1
    else:
        assert int(m.group(2)) == int(opts[1])
