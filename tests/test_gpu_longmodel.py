"""Forward/Backward for long models: the CTA-per-target kernels (csrc/fwd_mw.cu, bwd_mw.cu) against the reference's
golden vectors and against the one-warp kernels, in every CTA geometry family (W <= 4, <= 8, <= 12 warps).

M > 512 uses these kernels by default (synth_M800, synth_M2000 in test_gpu_parity.py); here the geometry is forced
through the tuning environment (P7GPU_MW_MINM / P7GPU_MW_CMAX, read at profile upload) so that shorter models and the
warp-boundary cases (model end inside a warp, padding warps) are covered as well.  Tolerance: north_star's 1e-4 relative.
"""
import os
import numpy as np
import pytest

from conftest import golden_set, GOLDEN_SETS
from biobench2_b200 import engine

pytestmark = pytest.mark.gpu
FB_RTOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


class forced_geometry:
    def __init__(self, minm, cmax):
        self.env = {"P7GPU_MW_MINM": str(minm), "P7GPU_MW_CMAX": str(cmax)}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        os.environ.update(self.env)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def check_against_golden(ctx, prof, db, g):
    v = g.vec
    sc, st, _, _, _ = engine.forward_batch(ctx, prof, db, g.db.lengths, want_xmx=False)
    ok = (v["fwd_st"] == 0) & (g.db.lengths > 0)
    assert np.array_equal(st[ok], v["fwd_st"][ok])
    np.testing.assert_allclose(sc[ok], v["fwd_sc"][ok], rtol=FB_RTOL, atol=1e-4)
    sel = v["xmx_sel"]
    lens = g.db.lengths[sel]
    sc, st, xmx, xoff, tot = engine.forward_batch(ctx, prof, db, lens, idx=sel)
    meta = v["fb_meta"]
    np.testing.assert_allclose(sc, meta[:, 0], rtol=FB_RTOL, atol=1e-4)
    f = xmx.reshape(-1, 6); rf = v["fwd_xmx"].reshape(-1, 6)
    assert np.array_equal(f[:, 5] > 1.0, rf[:, 5] > 1.0)          # identical rescaling decisions
    np.testing.assert_allclose(f, rf, rtol=FB_RTOL, atol=1e-30)
    bsc, bst, bx, btot = engine.backward_batch(ctx, prof, db, v["fwd_xmx"], xoff, idx=sel)
    assert (bst == 0).all()
    np.testing.assert_allclose(bsc, meta[:, 1], rtol=FB_RTOL, atol=1e-4)
    np.testing.assert_allclose(bx.reshape(-1, 6), v["bck_xmx"].reshape(-1, 6), rtol=FB_RTOL, atol=1e-30)
    bsc2, _, _, _ = engine.backward_batch(ctx, prof, db, xmx, xoff, idx=sel)
    np.testing.assert_allclose(bsc2, sc, rtol=FB_RTOL, atol=1e-4)    # Forward == Backward (fwdback.c:895-948)
    return sc, xmx, bx


@pytest.mark.parametrize("cmax", [0, 13, 8, 6])
@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_cta_per_target_kernels_vs_golden(ctx, name, cmax):
    g = golden_set(name)
    with forced_geometry(1, cmax):
        prof = engine.Profile(ctx, g.oprofile)
    db = engine.SeqDB.upload(ctx, g.db)
    try:
        check_against_golden(ctx, prof, db, g)
    finally:
        db.close(); prof.close()


@pytest.mark.parametrize("name", ["synth_M800", "synth_M400", "Pkinase"])        # the one-warp cross-check kernels cover M <= 1024
def test_cta_per_target_kernels_vs_one_warp_kernels(ctx, name):
    g = golden_set(name)
    with forced_geometry(1, 0):
        prof = engine.Profile(ctx, g.oprofile)
    db = engine.SeqDB.upload(ctx, g.db)
    try:
        a = check_against_golden(ctx, prof, db, g)
        ctx.set_option("fb_one_warp", 1)
        b = check_against_golden(ctx, prof, db, g)
        for x, y in zip(a, b):
            np.testing.assert_allclose(x, y, rtol=2e-5, atol=1e-30)
    finally:
        ctx.set_option("fb_one_warp", 0)
        db.close(); prof.close()
