"""The CPU oracle (oracle/p7_oracle.c) against the committed reference-generated golden vectors.

This is what pins the oracle: tests/golden/vec_*.npz were produced by the unmodified reference HMMER
(tests/golden/make_golden.py).  Integer stages must be bit-exact (score bits + status); the float
parsers within 1e-5 relative (k-ordered vs striped summation order).
"""
import numpy as np
import pytest

from conftest import bits
from oracle import port


@pytest.fixture
def oprof(gset):
    o = port.OracleProfile(gset.oprofile.view(), gset.oprofile)
    yield o
    o.close()


def test_length_scalars(gset, oprof):
    v = gset.vec
    for s in range(gset.db.nseq):
        L = int(gset.db.lengths[s])
        assert oprof.tjb_b(L) == v["tjb_b"][s]
        assert oprof.xw_move(L) == v["xw_move"][s]
        assert bits(oprof.xf_ncj(L)[0]) == bits(v["pmove"][s])
        assert bits(port.nullsc(L)) == bits(v["nullsc"][s])


@pytest.mark.parametrize("stage", ["ssv", "msv", "vit"])
def test_integer_filters_bit_exact(gset, oprof, stage):
    sc, st = oprof.stage_batch(stage, gset.db)
    ref_sc, ref_st = gset.vec[stage + "_sc"], gset.vec[stage + "_st"]
    assert np.array_equal(st, ref_st)
    keep = ref_st != 19 if stage == "ssv" else np.ones_like(ref_st, bool)   # eslENORESULT leaves the score undefined
    assert np.array_equal(bits(sc[keep]), bits(ref_sc[keep]))


def test_viterbi_always_dd_is_equivalent(gset, oprof):
    """evaluating D->D on every row gives the same scores as the lazy-F replica."""
    n = min(gset.db.nseq, 40)
    for s in range(n):
        L = int(gset.db.lengths[s])
        if L == 0:
            continue
        a = oprof.vit(gset.db.dsq(s), L)
        b = oprof.vit_alwaysdd(gset.db.dsq(s), L)
        assert a[0] == b[0] and bits(a[1]) == bits(b[1])


def test_forward_scores(gset, oprof):
    sc, st = oprof.stage_batch("fwd", gset.db)
    ref_sc, ref_st = gset.vec["fwd_sc"], gset.vec["fwd_st"]
    ok = (ref_st == 0) & (gset.db.lengths > 0)
    assert np.array_equal(st[ok], ref_st[ok])
    np.testing.assert_allclose(sc[ok], ref_sc[ok], rtol=1e-5, atol=1e-5)


def test_forward_backward_specials(gset, oprof):
    v = gset.vec
    pos = 0
    for j, s in enumerate(v["xmx_sel"]):
        L = int(gset.db.lengths[s])
        n = (L + 1) * 6
        fx = v["fwd_xmx"][pos:pos + n].reshape(L + 1, 6); bx = v["bck_xmx"][pos:pos + n].reshape(L + 1, 6)
        pos += n
        fwdsc, bcksc, ftot, btot, own, status = v["fb_meta"][j]
        st, sc, xmx, tot = oprof.fwd(gset.db.dsq(s), L)
        assert st == 0
        np.testing.assert_allclose(sc, fwdsc, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(tot, ftot, rtol=1e-6, atol=1e-6)
        assert np.array_equal(xmx[:, 5] > 1.0, fx[:, 5] > 1.0)   # identical rescaling decisions
        np.testing.assert_allclose(xmx, fx, rtol=2e-5, atol=1e-30)
        # backward from the REFERENCE's forward specials (same scale factors)
        st, sc, bxmx, tot, own2 = oprof.bwd(gset.db.dsq(s), L, fx)
        assert st == 0 and own2 == int(own)
        np.testing.assert_allclose(sc, bcksc, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(bxmx, bx, rtol=2e-5, atol=1e-30)
        # Forward == Backward (fwdback.c:895-948 utest property)
        np.testing.assert_allclose(fwdsc, bcksc, rtol=1e-4, atol=1e-4)


def test_bias_filter_score(gset):
    t, pi, eo, _ = gset.oprofile.bg_filter
    v = gset.vec
    for s in range(gset.db.nseq):
        L = int(gset.db.lengths[s])
        if L == 0:
            continue
        got = port.filtersc((t, pi, eo), gset.db.dsq(s), L)
        assert bits(got) == bits(v["filtersc"][s]), (s, L, got, v["filtersc"][s])


def test_pipeline_counters(gset, oprof):
    t, pi, eo, _ = gset.oprofile.bg_filter
    r = oprof.pipeline(gset.db, (t, pi, eo))
    assert list(r["counters"]) == list(gset.vec["pli_counters"])
    r = oprof.pipeline(gset.db, (t, pi, eo), F1=1.0, F2=1.0, F3=1.0, do_biasfilter=False)
    assert list(r["counters"]) == list(gset.vec["pli_counters_max"])


def test_survival_functions():
    from oracle import ref
    if not ref.available():
        pytest.skip("reference build not present")
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, mu, lam = rng.normal(0, 10), rng.normal(-8, 3), abs(rng.normal(0.69, 0.1))
        assert port.gumbel_surv(x, mu, lam) == ref.lib().p7ref_gumbel_surv(x, mu, lam)
        assert port.exp_surv(x, mu, lam) == ref.lib().p7ref_exp_surv(x, mu, lam)
