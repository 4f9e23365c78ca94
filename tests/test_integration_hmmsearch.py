"""integration/impl_b200: the reference's own hmmsearch host code linked on top of libp7gpu.so.

The parity artefact of the whole path: `hmmsearch_b200` and the unmodified reference `oracle/_ref/hmmsearch` run on the same
files; the six pipeline counters must be equal and every non-comment line of the main output (hit tables, domain tables,
alignments), of --tblout and of --domtblout must be BYTE-IDENTICAL.  That holds because (i) the integer filters are
bit-exact, (ii) the Forward/Backward specials and scores that reach the host come from the exact-order kernels
(csrc/fbx.cu) and (iii) everything after the Backward parser is the reference's own p7_Pipeline() (replayed, unmodified).

`impl_b200_utest` compares the five drop-in symbols with the reference's SSE implementations in one process, bit for bit.

CPU part (-m "not gpu"): the binaries exist, carry both symbol sets, print help, and refuse to run without a GPU.
Both binaries are built by integration/Makefile and oracle/Makefile from the reference sources where they lie (authoring
container); they travel to the GPU box with the snapshot.  Tests skip when a binary or input file is absent.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from biobench2_b200 import engine

B200 = os.path.join(ROOT, "integration", "_build", "hmmsearch_b200")
UTEST = os.path.join(ROOT, "integration", "_build", "impl_b200_utest")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "hmmsearch")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFDATA = os.path.join(ROOT, "tests", "_refdata")

need_bins = pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REFBIN) and os.path.exists(UTEST)),
                               reason="integration/_build or oracle/_ref not built (needs the reference sources)")

# lines that legitimately differ: '#' comments (banner, option echo, CPU time, Mc/sec) — everything else must match
TIMING = re.compile(r"^# (CPU time|Mc/sec)")


def payload(path):
    with open(path, errors="replace") as f:
        return [l for l in f.read().split("\n") if not l.startswith("#")]


def counters_of(path):
    out = {}
    with open(path) as f:
        for l in f:
            m = re.match(r"^(Target sequences|Passed MSV filter|Passed bias filter|Passed Vit filter|Passed Fwd filter):\s+(\d+)\s+\((\d+ residues)?", l)
            if m:
                out[m.group(1)] = int(m.group(2))
                if m.group(3):
                    out["residues"] = int(m.group(3).split()[0])
            m = re.match(r"^(Initial search space \(Z\)|Domain search space  \(domZ\)):\s+(\d+)", l)
            if m:
                out[m.group(1)] = int(m.group(2))
    return out


def run_both(tmp_path, hmm, fasta, extra=(), b200_extra=(), tag="x"):
    res = {}
    for name, exe, more in (("ref", REFBIN, ()), ("b200", B200, tuple(b200_extra))):
        o = tmp_path / ("%s_%s.out" % (tag, name)); t = tmp_path / ("%s_%s.tbl" % (tag, name)); d = tmp_path / ("%s_%s.dom" % (tag, name))
        cmd = [exe, "-o", str(o), "--tblout", str(t), "--domtblout", str(d)] + list(extra) + list(more) + [str(hmm), str(fasta)]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
        assert p.returncode == 0, (cmd, p.stdout[-2000:], p.stderr[-2000:])
        res[name] = (str(o), str(t), str(d))
    return res


def assert_identical(res, what=""):
    ro, rt, rd = res["ref"]; bo, bt, bd = res["b200"]
    cr, cb = counters_of(ro), counters_of(bo)
    assert cr == cb and len(cr) >= 7, (what, cr, cb)
    for label, a, b in (("--tblout", rt, bt), ("--domtblout", rd, bd), ("main output", ro, bo)):
        la, lb = payload(a), payload(b)
        if la != lb:
            diff = [(x, y) for x, y in zip(la, lb) if x != y]
            pytest.fail("%s %s: %d of %d lines differ (ref %d lines, b200 %d lines); first: %r" %
                        (what, label, len(diff) + abs(len(la) - len(lb)), len(la), len(la), len(lb), diff[:3]))
    return cr


# ----------------------------------------------------------------------------------------------- CPU part
@need_bins
def test_binaries_carry_both_engines():
    """the five exact-signature symbols (impl_sse.h:309,311,325,328,344) are impl_b200's; the reference's own are kept as *_sse"""
    syms = subprocess.run(["nm", B200], capture_output=True, text=True).stdout
    for s in ("p7_SSVFilter", "p7_MSVFilter", "p7_ViterbiFilter", "p7_ForwardParser", "p7_BackwardParser"):
        assert re.search(r" T %s$" % s, syms, re.M), s
        assert re.search(r" T %s_sse$" % s, syms, re.M), s + "_sse"
    assert re.search(r" U p7gpu_pipeline_run_file_sharded", syms) and re.search(r" T p7_Pipeline$", syms, re.M)
    h = subprocess.run([B200, "-h"], capture_output=True, text=True)
    assert h.returncode == 0 and "--gpus" in h.stdout and "--tblout" in h.stdout and "--F3" in h.stdout


@need_bins
def test_no_cpu_fallback_in_the_binary(tmp_path):
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    fa = tmp_path / "t.fa"; fa.write_text(">a\nACDEFGHIKLMNPQRSTVWY\n")
    p = subprocess.run([B200, os.path.join(GOLDEN, "globins4.hmm"), str(fa)], capture_output=True, text=True)
    assert p.returncode != 0 and "CUDA device" in (p.stdout + p.stderr)
    p = subprocess.run([UTEST, "-N", "2", "-H", "0"], capture_output=True, text=True)
    assert p.returncode != 0                                   # the five symbols raise eslESYS without a device


# ----------------------------------------------------------------------------------------------- GPU part
def seeded_fasta(path, hmm_path, n, seed, nhom):
    """n iid background targets (mean length 350) with nhom homologs emitted from the profile planted among them"""
    from biobench2_b200 import seqdb
    from oracle import ref
    db = seqdb.synth_db(n, mean=350, seed=seed, degenerate_frac=0.002)
    syms = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-BJZOUX*~", np.uint8)
    rp = ref.RefProfile.open(hmm_path)
    rng = np.random.default_rng(seed)
    where = set(int(x) for x in rng.choice(n, size=nhom, replace=False))
    with open(path, "wb") as f:
        k = 0
        for s in range(n):
            if s in where:
                h = rp.emit(1000 + k, Lcfg=350); k += 1
                body = h if h is not None and len(h) else db.residues[db.offsets[s]: db.offsets[s] + db.lengths[s]]
                f.write(b">hom%06d planted homolog %d\n" % (s, k))
            else:
                body = db.residues[db.offsets[s]: db.offsets[s] + db.lengths[s]]
                f.write(b">seq%06d random background  target\n" % s)
            t = syms[body].tobytes()
            for a in range(0, len(t), 60):
                f.write(t[a:a + 60] + b"\n")
    rp.close()


@pytest.mark.gpu
@need_bins
@pytest.mark.parametrize("args", [["-M", "33"], ["-M", "145"], ["-M", "400", "-L", "350"], ["-M", "1100", "-N", "40", "-H", "6"],
                                  [os.path.join(GOLDEN, "globins4.hmm")], [os.path.join(GOLDEN, "fn3.hmm")],
                                  [os.path.join(GOLDEN, "Pkinase.hmm"), "-L", "500"]])
def test_five_symbols_bit_identical_to_sse_in_one_process(args):
    p = subprocess.run([UTEST] + args, capture_output=True, text=True, timeout=1800)
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert p.stdout.count("bit-identical") == 5


@pytest.mark.gpu
@need_bins
def test_config1_tufa420_vs_swissprot_full_size(tmp_path):
    """BASELINE.json configs[0], runBench.pl:115: hmmsearch tufa420.hmm uniprot_sprot.fasta — 153,325 targets"""
    hmm, fa = os.path.join(REFDATA, "tufa420.hmm"), os.path.join(REFDATA, "uniprot_sprot.fasta")
    if not (os.path.exists(hmm) and os.path.exists(fa)):
        pytest.skip("tests/_refdata not unpacked (HMMER/input/inputData.tgz)")
    res = run_both(tmp_path, hmm, fa, tag="sprot")
    c = assert_identical(res, "tufa420 x Swiss-Prot")
    assert c["Target sequences"] == 153325 and c["residues"] == 56403273
    assert (c["Passed MSV filter"], c["Passed bias filter"], c["Passed Vit filter"], c["Passed Fwd filter"]) == (4765, 4765, 1552, 743)
    assert c["Domain search space  (domZ)"] == 741


@pytest.mark.gpu
@need_bins
@pytest.mark.parametrize("model", ["Pkinase", "globins4", "fn3"])
def test_seeded_database_1e5_targets(tmp_path, model):
    hmm = os.path.join(GOLDEN, model + ".hmm")
    fa = tmp_path / "db.fa"
    seeded_fasta(str(fa), hmm, 100000, seed={"Pkinase": 3, "globins4": 4, "fn3": 5}[model], nhom=300)
    c = assert_identical(run_both(tmp_path, hmm, fa, tag=model), model)
    assert c["Target sequences"] == 100000 and c["Passed Fwd filter"] >= 60


@pytest.mark.gpu
@need_bins
def test_max_mode_and_option_variants(tmp_path):
    hmm = os.path.join(GOLDEN, "fn3.hmm")
    fa = tmp_path / "db.fa"
    seeded_fasta(str(fa), hmm, 3000, seed=9, nhom=40)
    assert_identical(run_both(tmp_path, hmm, fa, extra=["--max"], tag="max"), "--max")
    assert_identical(run_both(tmp_path, hmm, fa, extra=["--nobias", "--F1", "0.05", "--F2", "0.01", "--F3", "1e-3"], tag="f"), "--nobias --F*")
    assert_identical(run_both(tmp_path, hmm, fa, extra=["-E", "1e-5", "--domE", "1e-3", "--nonull2", "--noali", "-Z", "1e6"], tag="e"), "-E --nonull2 -Z")
    assert_identical(run_both(tmp_path, hmm, fa, extra=["-T", "20", "--domT", "10", "--incT", "25", "--acc", "--notextw"], tag="t"), "-T")


@pytest.mark.gpu
@need_bins
def test_multi_query_file_and_packed_database_reuse(tmp_path):
    hmms = tmp_path / "three.hmm"
    hmms.write_bytes(b"".join(open(os.path.join(GOLDEN, m + ".hmm"), "rb").read() for m in ("globins4", "fn3", "Pkinase")))
    fa = tmp_path / "db.fa"
    seeded_fasta(str(fa), os.path.join(GOLDEN, "fn3.hmm"), 20000, seed=12, nhom=60)
    res = run_both(tmp_path, hmms, fa, b200_extra=["--p7db", str(tmp_path / "db.p7db")], tag="multi")
    assert_identical(res, "three queries")
    # second run straight from the packed database written by the first
    res2 = run_both(tmp_path, hmms, fa, tag="again")
    o = tmp_path / "p7db.out"; t = tmp_path / "p7db.tbl"; d = tmp_path / "p7db.dom"
    p = subprocess.run([B200, "-o", str(o), "--tblout", str(t), "--domtblout", str(d), str(hmms), str(tmp_path / "db.p7db")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert payload(str(t)) == payload(res2["ref"][1]) and payload(str(d)) == payload(res2["ref"][2])


@pytest.mark.gpu
@need_bins
def test_two_gpus_same_output(tmp_path):
    if engine.device_count() < 2:
        pytest.skip("needs two GPUs")
    hmm = os.path.join(GOLDEN, "Pkinase.hmm")
    fa = tmp_path / "db.fa"
    seeded_fasta(str(fa), hmm, 50000, seed=21, nhom=200)
    assert_identical(run_both(tmp_path, hmm, fa, b200_extra=["--gpus", "2", "--cpu", "3"], tag="g2"), "--gpus 2")


@pytest.mark.gpu
@need_bins
def test_fast_specials_report_differing_lines(tmp_path):
    """--fbfast: Forward/Backward specials from the fast kernels (<= 1e-4 relative).  Counters and the set of reported targets
    must still be the reference's; the number of table lines whose text differs is printed as a measured statistic."""
    hmm = os.path.join(GOLDEN, "Pkinase.hmm")
    fa = tmp_path / "db.fa"
    seeded_fasta(str(fa), hmm, 50000, seed=22, nhom=400)
    res = run_both(tmp_path, hmm, fa, b200_extra=["--fbfast"], tag="fast")
    assert counters_of(res["ref"][0]) == counters_of(res["b200"][0])
    a, b = payload(res["ref"][1]), payload(res["b200"][1])
    assert [l.split()[0] for l in a if l] == [l.split()[0] for l in b if l]
    ndiff = sum(x != y for x, y in zip(a, b))
    da, dbb = payload(res["ref"][2]), payload(res["b200"][2])
    nddiff = sum(x != y for x, y in zip(da, dbb)) + abs(len(da) - len(dbb))
    print("--fbfast: %d of %d --tblout lines and %d of %d --domtblout lines differ in text from the reference" % (ndiff, len(a), nddiff, len(da)))


@pytest.mark.gpu
def test_two_ranks_nccl_gather_into_the_reference_hit_list():
    """one process per GPU (torch.distributed, NCCL): survivors + specials gathered to rank 0, finished by the reference's
    post-processing and p7_tophits_SortBySortkey/Threshold == the same search on one rank == the reference's own pipeline"""
    from oracle import ref
    if engine.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not ref.available():
        pytest.skip("oracle/_ref/libp7ref.so not built")
    p = subprocess.run(["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "scripts", "nccl_merge_check.py"), "Pkinase", "60000"],
                       capture_output=True, text=True, timeout=1500)
    print(p.stdout[-2000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert '"identical": true' in p.stdout
