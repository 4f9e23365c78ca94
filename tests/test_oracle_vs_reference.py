"""The CPU oracle against the live reference build (oracle/_ref/libp7ref.so) on fresh seeded inputs.

Skipped when the reference build is absent (it exists in the authoring container and travels to the GPU box).
"""
import numpy as np
import pytest

from conftest import bits, golden_set
from biobench2_b200 import seqdb
from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libp7ref.so not built")


@pytest.mark.parametrize("M,seed", [(7, 1), (64, 2), (145, 3), (257, 4)])
def test_sampled_models_bit_exact(M, seed):
    """Same shape as the reference's own utests (msvfilter.c:708-726): sampled model, iid targets, M=1.. etc."""
    p = ref.RefProfile.sample(M, seed, calibrate=False)
    o = port.OracleProfile(p.view(), p)
    db = seqdb.synth_db(60, mean=200, seed=seed, degenerate_frac=0.02)
    for stage in ("ssv", "msv", "vit"):
        a_sc, a_st = p.stage_batch(stage, db.residues, db.offsets, db.lengths)
        b_sc, b_st = o.stage_batch(stage, db)
        assert np.array_equal(a_st, b_st), stage
        keep = a_st != 19
        assert np.array_equal(bits(a_sc[keep]), bits(b_sc[keep])), stage
    a_sc, a_st = p.stage_batch("fwd", db.residues, db.offsets, db.lengths)
    b_sc, b_st = o.stage_batch("fwd", db)
    np.testing.assert_allclose(b_sc, a_sc, rtol=1e-5, atol=1e-5)
    o.close(); p.close()


def test_long_target_forces_msv_j_path():
    """tjb+tbm+tec+bias >= 127 (ssvfilter.c:884) makes SSV answer eslENORESULT for every target."""
    g = golden_set("synth_M2000")
    p_view = g.oprofile.view()
    o = port.OracleProfile(p_view, g.oprofile)
    rp = None
    # the reference profile is rebuilt from the same recipe the golden generator used
    rp = ref.RefProfile.sample(2000, 42, True)
    db = seqdb.synth_db(1, fixed_length=200000, seed=9)
    a_sc, a_st = rp.stage_batch("ssv", db.residues, db.offsets, db.lengths)
    b_sc, b_st = o.stage_batch("ssv", db)
    assert np.array_equal(a_st, b_st) and (a_st == 19).all()
    a_sc, a_st = rp.stage_batch("msv", db.residues, db.offsets, db.lengths)
    b_sc, b_st = o.stage_batch("msv", db)
    assert np.array_equal(a_st, b_st) and np.array_equal(bits(a_sc), bits(b_sc))
    o.close(); rp.close()
