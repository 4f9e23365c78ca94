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
