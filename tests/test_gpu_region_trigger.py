"""GPU: the on-device region trigger (csrc/decode.cu) — mocc maxima bit-exact against the oracle on the same
specials, the flag equal to the reference's rt1 trigger, and the --max pipeline returning specials only where it fires
while the final hit list stays the reference's."""
import os
import numpy as np
import pytest

from conftest import bits, golden_set, GOLDEN, GOLDEN_SETS
from biobench2_b200 import engine, seqdb
from oracle import port, ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_trigger_kernel_bit_exact_vs_oracle_on_reference_specials(ctx, name):
    g = golden_set(name)
    v = g.vec
    sel = v["xmx_sel"]
    lens = g.db.lengths[sel]
    xoff = np.concatenate([[0], np.cumsum((lens.astype(np.int64) + 1) * 6)])[:-1]
    prof = engine.Profile(ctx, g.oprofile); db = engine.SeqDB.upload(ctx, g.db)
    o = port.OracleProfile(g.oprofile.view(), g.oprofile)
    mocc, flag = engine.region_trigger_batch(ctx, prof, db, v["fwd_xmx"], v["bck_xmx"], xoff, idx=sel)
    for j, s in enumerate(sel):
        L = int(lens[j])
        if L == 0:
            continue
        fx = v["fwd_xmx"][xoff[j]: xoff[j] + (L + 1) * 6]; bx = v["bck_xmx"][xoff[j]: xoff[j] + (L + 1) * 6]
        d = o.domain_decoding(L, fx, bx, 0)
        assert bits(mocc[j]) == bits(d["mocc_max"]), (name, j, mocc[j], d["mocc_max"])
        assert flag[j] == d["triggered"]
    # own-scales targets are left to the host
    _, flag2 = engine.region_trigger_batch(ctx, prof, db, v["fwd_xmx"], v["bck_xmx"], xoff, idx=sel, own_scales=np.ones(len(sel), np.int32))
    assert (flag2 == 1).all()
    o.close(); db.close(); prof.close()


@pytest.mark.parametrize("name", ["synth_M400", "Pkinase", "synth_M800"])
def test_max_pipeline_returns_specials_only_where_the_trigger_fires(ctx, name):
    g = golden_set(name)
    prof = engine.Profile(ctx, g.oprofile)
    hdb = seqdb.synth_db(3000, mean=300, seed=5, degenerate_frac=0.01)
    hdb = seqdb.concat([g.db, hdb])
    db = engine.SeqDB.upload(ctx, hdb)
    full = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=True)
    trig = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=True, region_trigger=True)
    assert list(full.counters) == list(trig.counters)
    assert np.array_equal(full.survivors["seqidx"], trig.survivors["seqidx"])
    assert (full.survivors["region_flag"] == -1).all() and (full.survivors["xmx_offset"] >= 0).all()
    fl = trig.survivors["region_flag"]
    assert set(np.unique(fl)) <= {0, 1}
    assert ((trig.survivors["xmx_offset"] >= 0) == (fl == 1)).all()
    assert 0 < (fl == 1).sum() < len(fl) or name != "synth_M400"
    o = port.OracleProfile(g.oprofile.view(), g.oprofile)
    for a, b in zip(full.survivors, trig.survivors):
        n = (int(a["L"]) + 1) * 6
        fx = full.fwd_xmx[a["xmx_offset"]: a["xmx_offset"] + n]; bx = full.bck_xmx[a["xmx_offset"]: a["xmx_offset"] + n]
        d = o.domain_decoding(int(a["L"]), fx, bx, int(a["bck_own_scales"]))
        # fast specials (<= 1e-4 relative): the pipeline flags conservatively, rt1 - 1e-4 (decode.cu); exact-order specials: rt1 itself
        want = 1 if (d["triggered"] or a["bck_own_scales"] or d["mocc_max"] >= np.float32(0.25) - np.float32(1.0e-4)) else 0
        assert b["region_flag"] == want
        assert bits(b["mocc_max"]) == bits(d["mocc_max"])
        if b["region_flag"] == 1:                 # the returned specials are the same bytes
            assert np.array_equal(bits(trig.fwd_xmx[b["xmx_offset"]: b["xmx_offset"] + n]), bits(fx))
            assert np.array_equal(bits(trig.bck_xmx[b["xmx_offset"]: b["xmx_offset"] + n]), bits(bx))
    o.close(); db.close(); prof.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libp7ref.so not built")
@pytest.mark.parametrize("name", ["globins4", "fn3"])
def test_max_mode_hit_list_vs_reference_with_device_trigger(ctx, name):
    """hmmsearch --max: GPU Forward/Backward on every target, region trigger on the device, the reference's unchanged
    domain definition / hit creation on the flagged targets only -> the reference's own --max hit list."""
    g = golden_set(name)
    rp = ref.RefProfile.open(os.path.join(GOLDEN, "%s.hmm" % name))
    prof = engine.Profile(ctx, g.oprofile)
    r = engine.pipeline_run_host(ctx, prof, g.db, do_max=True, region_trigger=True)
    counters, ref_hits = rp.pipeline(g.db.residues, g.db.offsets, g.db.lengths, do_max=True)
    assert list(r.counters) == list(counters)
    sv = r.survivors[r.survivors["region_flag"] == 1]
    assert len(sv) < len(r.survivors)
    hits = rp.finish_from_specials(g.db.residues, g.db.offsets, g.db.lengths, sv["seqidx"], sv["fwdsc"], r.fwd_xmx, r.bck_xmx,
                                   sv["xmx_offset"], sv["fwd_totscale"], sv["bck_totscale"],
                                   int(r.counters[0]), int(r.counters[1]), do_max=True)
    # the reference's pre-filter at hit creation uses the running Z (p7_pipeline.c:577,804) while the finishing harness knows
    # the final Z: compare what p7_tophits_Threshold (final Z, p7_tophits.c:865-945) REPORTS — the --tblout body
    REPORTED = 1 << 1                                  # p7_IS_REPORTED, hmmer.h:693
    hits = [h for h in hits if h.flags & REPORTED]; ref_hits = [h for h in ref_hits if h.flags & REPORTED]
    assert len(ref_hits) >= 4
    assert [h.seqidx for h in hits] == [h.seqidx for h in ref_hits]
    assert [h.flags for h in hits] == [h.flags for h in ref_hits]
    assert [h.ndom for h in hits] == [h.ndom for h in ref_hits]
    assert [(h.best_dom_ienv, h.best_dom_jenv) for h in hits] == [(h.best_dom_ienv, h.best_dom_jenv) for h in ref_hits]
    np.testing.assert_allclose([h.score for h in hits], [h.score for h in ref_hits], atol=1e-3)
    rp.close(); prof.close()
