"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol include/p7gpu.h declares,
fails loudly without a device, and the product never routes through the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from biobench2_b200 import engine, seqdb


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "p7gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p7gpu_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(engine.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(engine.EXPORTS) == declared


def test_no_cpu_fallback_without_device():
    if engine.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.P7GpuError) as e:
        engine.Context(0)
    assert e.value.code == engine.ESYS
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """Nothing under biobench2_b200/ may import, link, dlopen or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "biobench2_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|#include\s*[\"<][^\n]*oracle|liboracle|libp7ref|oracle/_ref|p7o_|p7ref_")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_host_scalar_helpers_match_oracle():
    from oracle import port
    lib = engine.lib()
    for L in (1, 2, 17, 350, 5000, 43786, 100000):
        assert np.float32(lib.p7gpu_nullsc(L)).view(np.uint32) == np.float32(port.nullsc(L)).view(np.uint32)
    assert lib.p7gpu_msv_pvalue(-13.17, -6.86, -11.87, 0.698) == port.gumbel_surv(
        float(np.float32((np.float32(-13.17) - np.float32(-6.86)) / 0.69314718055994529)), float(np.float32(-11.87)), float(np.float32(0.698)))


def test_packed_db_layout_and_sharding():
    db = seqdb.synth_db(500, mean=120, seed=1)
    assert db.residues[0] == 255
    for s in (0, 1, 250, 499):
        d = db.dsq(s)
        assert d[0] == 255 and d[-1] == 255 and len(d) == db.lengths[s] + 2 and (d[1:-1] < 20).all()
    cover = []
    for r in range(4):
        lo, hi = db.shard(r, 4)
        cover.extend(range(lo, hi))
        res = int(db.lengths[lo:hi].astype(np.int64).sum())
        assert abs(res - db.nres / 4) < 0.1 * db.nres
    assert cover == list(range(db.nseq))
    sub = db.subset([3, 5])
    assert np.array_equal(sub.dsq(1), db.dsq(5))


def test_digitize_matches_easel_alphabet():
    assert list(seqdb.digitize("ACDEFGHIKLMNPQRSTVWY-BJZOUX*~")) == list(range(29))
    with pytest.raises(ValueError):
        seqdb.digitize("AC1")


def test_pack5_transfer_format_round_trip_on_the_host():
    """p7gpu_pack5 is host code: eight residue codes (0..28; the 255 sentinel as 31) in five bytes, little-endian bit order,
    padded with sentinels to a multiple of 64 residues.  Decoding it with numpy must give the input back."""
    import numpy as np
    from biobench2_b200 import engine, seqdb
    hdb = seqdb.synth_db(300, mean=90, seed=8, degenerate_frac=0.05)
    res = hdb.residues
    p5 = engine.pack5(res, nthreads=3)
    assert p5.shape[0] == (res.shape[0] + 63) // 64 * 40 + 40
    ngroups = p5.shape[0] // 5
    b = p5.reshape(ngroups, 5).astype(np.uint64)
    v = b[:, 0] | (b[:, 1] << np.uint64(8)) | (b[:, 2] << np.uint64(16)) | (b[:, 3] << np.uint64(24)) | (b[:, 4] << np.uint64(32))
    codes = np.stack([(v >> np.uint64(5 * r)) & np.uint64(31) for r in range(8)], axis=1).reshape(-1).astype(np.uint8)
    codes[codes == 31] = 255
    assert np.array_equal(codes[:res.shape[0]], res)
    assert (codes[res.shape[0]:] == 255).all()
    assert np.array_equal(engine.pack5(res, nthreads=1), p5)


def test_stage_thresholds_reproduce_the_p_value_tests():
    """The device-side stage decisions compare seq_score with a float threshold (p7gpu_stage_threshold, host code).  For
    calibrated parameter ranges and the pipeline's F values, `x >= T` must equal `not (P(x) > F)` with P evaluated by the
    oracle's restatement of esl_gumbel_surv / esl_exp_surv (pinned to the reference), at T, around T and far from it."""
    import numpy as np
    from biobench2_b200 import engine
    from oracle import port
    rng = np.random.default_rng(11)
    surv = {"gumbel": port.gumbel_surv, "exp": port.exp_surv}
    n = 0
    for kind in ("gumbel", "exp"):
        for F in (0.02, 1e-3, 1e-5, 0.3, 1e-9):
            for _ in range(12):
                mu = np.float32(rng.uniform(-25.0, 5.0)); lam = np.float32(rng.uniform(0.55, 0.85))
                T = engine.stage_threshold(kind, F, mu, lam)
                assert T is not None
                near = [np.nextafter(T, np.float32(-np.inf), dtype=np.float32), T, np.nextafter(T, np.float32(np.inf), dtype=np.float32)]
                xs = near + [np.float32(T + d) for d in rng.uniform(-3, 3, 40)] + [np.float32(T * (1 + e)) for e in rng.uniform(-1e-5, 1e-5, 40)]
                for x in xs:
                    passes = not (surv[kind](float(np.float32(x)), float(mu), float(lam)) > F)
                    assert (np.float32(x) >= T) == passes, (kind, F, mu, lam, x, T)
                    n += 1
    assert n > 9000
    assert engine.stage_threshold("gumbel", 0.02, -10.0, 0.7) is not None


def _scan_walk(tab, P, npass, seq, thresh):
    """ssv_scan_kernel's walk over the product's scan table (ssv_msv.cu), in numpy: per pass 32 lanes x 4 words x 2 halves follow
    their diagonals from word 256*pass + lane (+32q), one column per row, wrapping at P; running maximum over all passes;
    leave once a cell reaches the overflow threshold (checked every 64 rows)."""
    xE = 0
    for p in range(npass):
        v = np.zeros((32, 4, 2), np.int64)
        c = p * 256 + np.arange(32)
        for i, x in enumerate(seq):
            w = tab[x][c[:, None] + 32 * np.arange(4)[None, :]]
            g = np.stack([(w & 0xffff).astype(np.uint16).view(np.int16), (w >> 16).astype(np.uint16).view(np.int16)], -1).astype(np.int64)
            v = np.maximum(v + g, 0)
            assert v.max() < 32768
            xE = max(xE, int(v.max()))
            c = c + 1
            c[c >= P] -= P
            if (i & 63) == 63 and xE >= thresh:
                return xE
    return xE


def _ssv_epilogue(maxu, tjb, v):
    """the decision tree of p7_SSVFilter after get_xE (ssvfilter.c:884-925) -> (status, score bits or None, xJ)"""
    xE = min(128 + maxu, 255)
    if xE >= 255 - v.bias_b:
        return (engine.ENORESULT if v.base_b - tjb - v.tbm_b < 128 else engine.ERANGE), None, 0
    xE16 = ((xE + v.base_b - tjb - v.tbm_b) & 0xffff) - 128 & 0xffff
    if xE16 >= 255 - v.bias_b:
        return engine.ERANGE, None, 0
    xJ = (xE16 - v.tec_b) & 0xffff
    if xJ > v.base_b:
        return engine.ENORESULT, None, 0
    sc = np.float32(np.float32(np.float32(xJ - tjb) - np.float32(v.base_b)) / np.float32(v.scale_b)) - np.float32(3.0)
    return 0, np.float32(sc), xJ


@pytest.mark.parametrize("name", ["synth_M1", "synth_M33", "synth_M100", "globins4", "fn3", "Pkinase", "synth_M400", "synth_M800"])
def test_scan_table_walk_equals_oracle_ssv(name):
    """Host logic of the hmmscan block prefilter, no device: the table p7gpu_scan_table_build lays out for a model, walked the way
    ssv_scan_kernel walks it (pass by pass over groups of 256 diagonals), must give the CPU oracle's p7_SSVFilter status and
    score bits on seeded targets and on the reference-generated golden targets (homologs, overflow, degenerate residues)."""
    from conftest import golden_set, bits, GOLDEN
    from biobench2_b200.oprofile import OProfile
    from oracle import port
    op = OProfile.load(os.path.join(GOLDEN, "prof_%s.npz" % name))
    v = op.view()
    o = port.OracleProfile(v, op)
    g = golden_set(name)
    pick = np.argsort(-g.vec["msv_sc"] if "msv_sc" in g.vec else np.arange(g.db.nseq))[:6]
    hdb = seqdb.concat([g.db.subset(pick), seqdb.synth_db(10 if op.M > 300 else 30, mean=120, seed=op.M, degenerate_frac=0.03),
                        seqdb.pack([np.zeros(1, np.uint8), np.arange(70, dtype=np.uint8) % 29])])
    osc, ost = o.stage_batch("ssv", hdb)
    nchecked = 0
    for s in range(hdb.nseq):
        L = int(hdb.lengths[s])
        seq = hdb.residues[int(hdb.offsets[s]): int(hdb.offsets[s]) + L]
        tab, P, npass, tjb, xjmin = engine.scan_table(op, L)
        assert P == 256 * npass and P >= op.M + 1 and P - 256 < op.M + 1 and tab.shape == (29, P + 96)
        assert 0 <= xjmin <= 256
        if tjb + v.tbm_b + v.tec_b + v.bias_b >= 127:
            assert ost[s] == engine.ENORESULT
            continue
        st, sc, xJ = _ssv_epilogue(_scan_walk(tab, P, npass, seq, 127 - v.bias_b), tjb, v)
        assert st == ost[s], (name, s, L, st, ost[s])
        if st == 0:
            assert bits(sc) == bits(osc[s]), (name, s, L, sc, osc[s])
            nchecked += 1
    assert nchecked >= 5
    o.close()
