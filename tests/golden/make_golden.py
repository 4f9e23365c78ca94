#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libp7ref.so).

Run in the authoring container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
The outputs are committed; nothing at test/bench time reads /root/reference.

  prof_<name>.npz : "oprofile dump" — the striped P7_OPROFILE arrays + scalars exactly as the reference's
                    p7_oprofile_Convert() produced them, plus the bias-filter HMM (p7_bg_SetFilter).
  vec_<name>.npz  : a small packed target set and, per target, the reference's own answers for every
                    stage of p7_Pipeline(): SSV/MSV/Viterbi (score bits + status), Forward/Backward
                    (scores, totscale, xmx specials for the first few), nullsc, filtersc, per-L scalars,
                    and the whole-pipeline counters + hit list.
"""
import os
import sys
import tarfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref                      # noqa: E402
from biobench2_b200 import seqdb            # noqa: E402
from biobench2_b200.oprofile import OProfile  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REFROOT = "/root/reference/HMMER"
NXMX = 6   # targets per set with full Forward/Backward specials


def dump_profile(p, name):
    t, pi, eo = p.filter_hmm()
    op = OProfile.from_view(p.view(), name=name, bg_filter=(t, pi, eo, p.omega()))
    op.save(os.path.join(OUT, "prof_%s.npz" % name))
    return op


def golden_vectors(p, name, db, with_pipeline=True):
    n = db.nseq
    out = dict(residues=db.residues, offsets=db.offsets, lengths=db.lengths)
    for stage in ("ssv", "msv", "vit", "fwd"):
        sc, st = p.stage_batch(stage, db.residues, db.offsets, db.lengths)
        out[stage + "_sc"] = sc
        out[stage + "_st"] = st
    out["nullsc"] = np.array([p.nullsc(int(L)) for L in db.lengths], np.float32)
    out["filtersc"] = np.array([p.filtersc(db.dsq(s), int(db.lengths[s])) for s in range(n)], np.float32)
    ls = [p.length_scalars(int(L)) for L in db.lengths]
    out["tjb_b"] = np.array([a[0] for a in ls], np.int32)
    out["xw_move"] = np.array([a[1] for a in ls], np.int32)
    out["pmove"] = np.array([a[2] for a in ls], np.float32)
    # Forward/Backward specials for the first few non-empty targets
    xsel = [s for s in range(n) if db.lengths[s] > 0][:NXMX]
    out["xmx_sel"] = np.array(xsel, np.int64)
    fx, bx, meta = [], [], []
    for s in xsel:
        r = p.fwdbck(db.dsq(s), int(db.lengths[s]))
        fx.append(r["fwd_xmx"].ravel()); bx.append(r["bck_xmx"].ravel())
        meta.append([r["fwdsc"], r["bcksc"], r["fwd_totscale"], r["bck_totscale"], r["bck_own_scales"], r["status"]])
    out["fwd_xmx"] = np.concatenate(fx) if fx else np.zeros(0, np.float32)
    out["bck_xmx"] = np.concatenate(bx) if bx else np.zeros(0, np.float32)
    out["fb_meta"] = np.array(meta, np.float64)
    if with_pipeline:
        counters, hits = p.pipeline(db.residues, db.offsets, db.lengths)
        out["pli_counters"] = counters
        out["hit_seqidx"] = np.array([h.seqidx for h in hits], np.int64)
        out["hit_score"] = np.array([h.score for h in hits], np.float32)
        out["hit_lnP"] = np.array([h.lnP for h in hits], np.float64)
        out["hit_ndom"] = np.array([h.ndom for h in hits], np.int32)
        out["hit_flags"] = np.array([h.flags for h in hits], np.int32)
        c2, _ = p.pipeline(db.residues, db.offsets, db.lengths, do_max=True)
        out["pli_counters_max"] = c2
    np.savez_compressed(os.path.join(OUT, "vec_%s.npz" % name), **out)
    print("  %-28s n=%d  msv status %s  pli %s hits %d" % (
        name, n, dict(zip(*np.unique(out["msv_st"], return_counts=True))),
        list(out.get("pli_counters", [])), len(out.get("hit_seqidx", []))))


def seqs_of(db):
    return [db.residues[int(o):int(o) + int(L)] for o, L in zip(db.offsets, db.lengths)]


def mixed_db(p, n_random, n_homolog, seed, mean=350, extra=()):
    """iid background targets (1% degenerate residues) + homologs emitted from the profile + edge cases."""
    seqs = seqs_of(seqdb.synth_db(n_random, mean=mean, seed=seed, degenerate_frac=0.01))
    for j in range(n_homolog):
        x = p.emit(seed * 1000 + j, Lcfg=mean)
        if x is not None and len(x) > 0:
            seqs.append(x)
    seqs.extend(extra)
    rng = np.random.default_rng(seed + 1)
    order = rng.permutation(len(seqs))
    return seqdb.pack([seqs[i] for i in order])


def edge_cases(rng):
    f = seqdb.AMINO_FREQ / seqdb.AMINO_FREQ.sum()

    def iid(L):
        return rng.choice(20, size=L, p=f).astype(np.uint8)
    out = [iid(1), iid(2), iid(3), iid(15), iid(16), iid(17), iid(31), iid(33), iid(5000)]
    out.append(np.full(200, 26, np.uint8))          # poly-X
    out.append(np.full(120, 9, np.uint8))           # poly-L (low complexity)
    x = iid(300); x[50] = 27; x[100] = 28; x[150] = 20; out.append(x)   # '*', '~', '-' inside
    out.append(np.zeros(0, np.uint8))               # empty target (p7_pipeline.c:682)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(123)

    # ---- synthetic exact-M profiles (SURVEY.md §8d recipe i) ----
    for M, nrand, nhom in ((400, 260, 24), (100, 120, 10), (800, 60, 8), (2000, 24, 4), (33, 60, 6), (1, 30, 0)):
        name = "synth_M%d" % M
        p = ref.RefProfile.sample(M, 42, True)
        dump_profile(p, name)
        db = mixed_db(p, nrand, nhom, seed=M, extra=edge_cases(rng) if M in (400, 33) else ())
        golden_vectors(p, name, db)
        p.close()

    # ---- real models shipped with the reference ----
    for name, path in (("globins4", REFROOT + "/tutorial/globins4.hmm"), ("fn3", REFROOT + "/tutorial/fn3.hmm"),
                       ("Pkinase", REFROOT + "/tutorial/Pkinase.hmm")):
        p = ref.RefProfile.open(path)
        dump_profile(p, name)
        p.write_hmm(os.path.join(OUT, "%s.hmm" % name))     # re-serialised by p7_hmmfile_WriteASCII: lets GPU-box tests rebuild the profile
        if name == "globins4":
            seqs = seqs_of(seqdb.read_fasta(REFROOT + "/tutorial/globins45.fa")) + seqs_of(seqdb.synth_db(80, seed=5))
            db = seqdb.pack(seqs)
        else:
            db = mixed_db(p, 100, 12, seed=len(name))
        golden_vectors(p, name, db)
        p.close()

    # ---- the shipped biobench2 workload: tufa420.hmm (HMMER2 format) vs Swiss-Prot (runBench.pl:115) ----
    tgz = REFROOT + "/input/inputData.tgz"
    tmp = "/tmp/biobench2_input"
    if not os.path.exists(tmp):
        os.makedirs(tmp, exist_ok=True)
        with tarfile.open(tgz) as tf:
            tf.extractall(tmp)
    hmm = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs if f == "tufa420.hmm"][0]
    fa = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs if f == "uniprot_sprot.fasta"][0]
    p = ref.RefProfile.open(hmm)
    dump_profile(p, "tufa420")
    full = seqdb.read_fasta(fa, max_seqs=20000)
    # every 100th of the first 20000 Swiss-Prot targets + the top hits the reference reports among them
    _, hits = p.pipeline(full.residues, full.offsets, full.lengths)
    sub_idx = sorted(set(range(0, full.nseq, 100)) | set(int(h.seqidx) for h in hits[:40]))
    allseqs = seqs_of(full)
    db = seqdb.pack([allseqs[s] for s in sub_idx])
    golden_vectors(p, "tufa420_sprot", db)
    p.close()


if __name__ == "__main__":
    main()
