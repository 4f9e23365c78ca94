"""The batched p7_Pipeline() (p7gpu_pipeline_run / _run_host, p7gpu_msvfilter_stream) against the reference.

Parity surface: the six pipeline counters (they only match if every pass/fail decision matches),
the survivor set with its scores, and — after the reference's own unchanged host post-processing of our
Forward/Backward specials — the final hit list.  Needs a B200.
"""
import os

import numpy as np
import pytest

from conftest import bits, golden_set, GOLDEN_SETS, GOLDEN
from biobench2_b200 import engine, seqdb
from oracle import port, ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


def test_bias_filter_score_bit_exact_vs_golden(ctx):
    for name in GOLDEN_SETS:
        g = golden_set(name)
        prof = engine.Profile(ctx, g.oprofile)
        db = engine.SeqDB.upload(ctx, g.db)
        got = engine.biasfilter_batch(ctx, prof, db)
        keep = g.db.lengths > 0
        bad = np.nonzero(bits(got[keep]) != bits(g.vec["filtersc"][keep]))[0]
        assert bad.size == 0, (name, bad[:5], got[keep][bad[:5]], g.vec["filtersc"][keep][bad[:5]])
        db.close(); prof.close()


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_pipeline_counters_and_survivors_vs_golden(ctx, name):
    g = golden_set(name)
    prof = engine.Profile(ctx, g.oprofile)
    db = engine.SeqDB.upload(ctx, g.db)
    r = engine.pipeline_run(ctx, prof, db)
    assert list(r.counters) == list(g.vec["pli_counters"]), (list(r.counters), list(g.vec["pli_counters"]))
    r2 = engine.pipeline_run_host(ctx, prof, g.db)
    assert list(r2.counters) == list(g.vec["pli_counters"])
    assert [s["seqidx"] for s in r2.survivors] == [s["seqidx"] for s in r.survivors]
    rm = engine.pipeline_run(ctx, prof, db, do_max=True, want_specials=False)
    assert list(rm.counters) == list(g.vec["pli_counters_max"])
    # survivors against the CPU oracle's chain (scores: integer stages bit-exact, Forward within 1e-4)
    o = port.OracleProfile(g.oprofile.view(), g.oprofile)
    t, pi, eo, _ = g.oprofile.bg_filter
    oc = o.pipeline(g.db, (t, pi, eo))
    want = np.nonzero(oc["stage"] == 4)[0]
    assert [s["seqidx"] for s in r.survivors] == list(want)
    for s in r.survivors:
        i = s["seqidx"]
        assert bits(s["usc"]) == bits(oc["usc"][i]) and bits(s["filtersc"]) == bits(oc["filtersc"][i])
        if not np.isnan(oc["vfsc"][i]):
            assert bits(s["vfsc"]) == bits(oc["vfsc"][i])
        else:
            assert np.isnan(s["vfsc"]) and s["vit_status"] == -1      # Viterbi skipped (p7_pipeline.c:716)
        np.testing.assert_allclose(s["fwdsc"], oc["fwdsc"][i], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(s["bcksc"], s["fwdsc"], rtol=1e-4, atol=1e-4)
    o.close(); db.close(); prof.close()


def test_msv_only_mode_and_pass_list(ctx):
    """stop_after_msv (BASELINE config 2): survivors = targets with P(MSV) <= F1, decided exactly as p7_pipeline.c:691-693."""
    g = golden_set("synth_M400")
    hdb = seqdb.synth_db(30000, mean=300, seed=21, degenerate_frac=0.002)
    prof = engine.Profile(ctx, g.oprofile)
    db = engine.SeqDB.upload(ctx, hdb)
    o = port.OracleProfile(g.oprofile.view(), g.oprofile)
    osc, ost = o.stage_batch("msv", hdb)
    for F1 in (0.02, 0.3, 1e-4):
        want = []
        for s in range(hdb.nseq):
            L = int(hdb.lengths[s])
            seq_score = float(np.float32((np.float32(osc[s]) - port.nullsc(L)) / 0.69314718055994529))
            P = port.gumbel_surv(seq_score, float(np.float32(g.oprofile.evparam[0])), float(np.float32(g.oprofile.evparam[1])))
            if L > 0 and not (P > F1):
                want.append(s)
        r = engine.pipeline_run(ctx, prof, db, F1=F1, stop_after_msv=True)
        assert list(r.msv_pass["seqidx"]) == want, F1
        assert r.counters[2] == len(want) and r.counters[0] == hdb.nseq and r.counters[1] == hdb.nres
        assert np.array_equal(bits(r.msv_pass["usc"]), bits(osc[want]))
    o.close(); db.close(); prof.close()


def test_streamed_host_database_equals_resident(ctx):
    """p7gpu_msvfilter_stream: several chunks, copies overlapped with compute, same scores/statuses/pass list."""
    g = golden_set("synth_M100")
    hdb = seqdb.synth_db(420000, mean=350, seed=5)          # ~147 MB of residues -> 4 chunks
    prof = engine.Profile(ctx, g.oprofile)
    db = engine.SeqDB.upload(ctx, hdb)
    sc0, st0 = engine.msvfilter_batch(ctx, prof, db)
    sc1, st1, pl, kept = engine.msvfilter_stream(ctx, prof, hdb, F1=0.02, keep_db=True)
    assert np.array_equal(bits(sc0), bits(sc1)) and np.array_equal(st0, st1)
    r = engine.pipeline_run(ctx, prof, db, stop_after_msv=True)
    assert np.array_equal(pl, r.msv_pass["seqidx"].astype(np.int32))
    assert abs(len(pl) / hdb.nseq - 0.02) < 0.02               # F1 = 0.02 on a random database
    # the database returned by the stream is a normal resident database
    idx = np.arange(0, hdb.nseq, 5000, dtype=np.int64)
    a = engine.vitfilter_batch(ctx, prof, db, idx); b = engine.vitfilter_batch(ctx, prof, kept, idx)
    assert np.array_equal(bits(a[0]), bits(b[0]))
    kept.close(); db.close(); prof.close()


def test_stream_rejects_malformed_database(ctx):
    g = golden_set("synth_M100")
    hdb = seqdb.synth_db(100, mean=100, seed=1)
    bad = seqdb.PackedDB(hdb.residues, hdb.offsets.copy(), hdb.lengths.copy())
    bad.lengths[50] = 10 ** 6
    prof = engine.Profile(ctx, g.oprofile)
    with pytest.raises(engine.P7GpuError) as e:
        engine.msvfilter_stream(ctx, prof, bad)
    assert e.value.code == engine.EINVAL
    prof.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libp7ref.so not built")
@pytest.mark.parametrize("name,source", [("globins4", "file"), ("fn3", "file"), ("Pkinase", "file")])
def test_final_hit_list_vs_reference(ctx, name, source):
    """GPU survivors + specials -> the reference's own domain definition / hit creation (p7_pipeline.c:737-898) ->
    the same hit list the reference's whole pipeline reports (seqidx order, ndom, flags; scores to 1e-3 bits)."""
    g = golden_set(name)
    rp = ref.RefProfile.open(os.path.join(GOLDEN, "%s.hmm" % name))
    prof = engine.Profile(ctx, g.oprofile)
    r = engine.pipeline_run_host(ctx, prof, g.db)
    counters, ref_hits = rp.pipeline(g.db.residues, g.db.offsets, g.db.lengths)
    assert list(r.counters) == list(counters)
    sv = r.survivors
    hits = rp.finish_from_specials(g.db.residues, g.db.offsets, g.db.lengths,
                                   [s["seqidx"] for s in sv], [s["fwdsc"] for s in sv], r.fwd_xmx, r.bck_xmx,
                                   [s["xmx_offset"] for s in sv], [s["fwd_totscale"] for s in sv], [s["bck_totscale"] for s in sv],
                                   int(r.counters[0]), int(r.counters[1]))
    assert len(ref_hits) >= 4
    assert [h.seqidx for h in hits] == [h.seqidx for h in ref_hits]
    assert [h.ndom for h in hits] == [h.ndom for h in ref_hits]
    assert [h.flags for h in hits] == [h.flags for h in ref_hits]
    assert [(h.best_dom_ienv, h.best_dom_jenv) for h in hits] == [(h.best_dom_ienv, h.best_dom_jenv) for h in ref_hits]
    np.testing.assert_allclose([h.score for h in hits], [h.score for h in ref_hits], atol=1e-3)
    np.testing.assert_allclose([h.lnP for h in hits], [h.lnP for h in ref_hits], rtol=1e-4, atol=1e-3)
    exact = sum(1 for a, b in zip(hits, ref_hits) if np.float32(a.score).view(np.uint32) == np.float32(b.score).view(np.uint32))
    print("%s: %d hits, %d with bit-identical bit scores" % (name, len(hits), exact))
    rp.close(); prof.close()


def test_partition_independence_of_sharded_search(ctx):
    """Any sharding gives the same counters and survivors (SURVEY.md §8b 'partition-independence'): run the shards of a
    3-way split one after the other on this GPU and merge them like rank 0 would."""
    from biobench2_b200 import search
    g = golden_set("globins4")
    prof = engine.Profile(ctx, g.oprofile)
    whole = engine.pipeline_run_host(ctx, prof, g.db, want_specials=False)
    ctr = np.zeros(6, np.uint64); recs = []
    for r in range(3):
        shard, first = search.shard_db(g.db, r, 3)
        res = engine.pipeline_run_host(ctx, prof, shard, want_specials=False)
        ctr += res.counters
        recs.append(search.survivors_to_records(res, first))
    rec = np.concatenate(recs)
    assert list(ctr) == list(whole.counters)
    assert list(rec["seqidx"]) == [s["seqidx"] for s in whole.survivors]
    assert np.array_equal(bits(rec["fwdsc"]), bits(np.array([s["fwdsc"] for s in whole.survivors], np.float32)))
    c1, r1 = search.sharded_pipeline(ctx, prof, g.db)                       # world size 1: no process group needed
    assert list(c1) == list(whole.counters) and np.array_equal(r1["seqidx"], rec["seqidx"])
    prof.close()


def test_many_profiles_one_resident_database(ctx):
    """hmmscan-style batch (BASELINE config 5): the database is uploaded once, each profile is a pipeline call on it
    (the outer loop of hmmsearch.c:471 with the per-query FASTA rewind of :480-487 gone); counters match the oracle's chain."""
    hdb = seqdb.synth_db(4000, mean=250, seed=31, degenerate_frac=0.003)
    db = engine.SeqDB.upload(ctx, hdb)
    for name in ("synth_M33", "synth_M100", "fn3", "globins4", "Pkinase", "synth_M400", "synth_M800"):
        g = golden_set(name)
        prof = engine.Profile(ctx, g.oprofile)
        r = engine.pipeline_run(ctx, prof, db, want_specials=False)
        o = port.OracleProfile(g.oprofile.view(), g.oprofile)
        t, pi, eo, _ = g.oprofile.bg_filter
        oc = o.pipeline(hdb, (t, pi, eo))
        assert list(r.counters) == list(oc["counters"]), name
        assert [s["seqidx"] for s in r.survivors] == list(np.nonzero(oc["stage"] == 4)[0]), name
        o.close(); prof.close()
    db.close()


def test_pipeline_run_many_equals_one_call_per_profile(ctx):
    """p7gpu_pipeline_run_many (hmmscan-style batch, BASELINE config 5) against one p7gpu_pipeline_run per profile, on the
    committed scan profiles (tests/golden/scan, M = 100..800)."""
    import glob
    from biobench2_b200.oprofile import OProfile
    paths = sorted(glob.glob(os.path.join(GOLDEN, "scan", "prof_scan_M*.npz")))
    assert len(paths) >= 8
    hdb = seqdb.synth_db(6000, mean=300, seed=77, degenerate_frac=0.003)
    db = engine.SeqDB.upload(ctx, hdb)
    profs = [engine.Profile(ctx, OProfile.load(p)) for p in paths]
    many = engine.pipeline_run_many(ctx, profs, db, want_specials=False)
    npast = 0
    for prof, r in zip(profs, many):
        one = engine.pipeline_run(ctx, prof, db, want_specials=False)
        assert list(one.counters) == list(r.counters)
        assert np.array_equal(one.survivors["seqidx"], r.survivors["seqidx"])
        assert np.array_equal(bits(one.survivors["fwdsc"]), bits(r.survivors["fwdsc"]))
        npast += int(r.counters[2])
    assert npast > 0
    for p in profs:
        p.close()
    db.close()


@pytest.mark.parametrize("nq,F1", [(1, 0.02), (7, 0.02), (300, 0.02), (300, 0.2), (40, 1e-4)])
def test_scan_msv_prefilter_flags_every_model_with_an_msv_survivor(ctx, nq, F1):
    """p7gpu_scan_msv_prefilter (one launch over (model, query) pairs, gain table picked per item) against the per-model path
    hmmscan_b200 ran before it: for every model, the count must equal the number of queries whose SSV score passes F1 or whose
    SSV status is eslERANGE/eslENORESULT in p7gpu_ssvfilter_batch, and a model the MSV stage of the pipeline finds a survivor
    for is never reported as 0.  Models M = 1 .. 2000 (1 to 8 passes of 256 diagonals), queries with homologs and degenerates."""
    import glob
    from biobench2_b200.oprofile import OProfile
    paths = sorted(glob.glob(os.path.join(GOLDEN, "scan", "prof_scan_M*.npz"))) + \
        [os.path.join(GOLDEN, "prof_%s.npz" % n) for n in ("synth_M1", "synth_M33", "synth_M2000", "globins4", "Pkinase", "fn3")]
    ops = [OProfile.load(p) for p in paths]
    parts = [seqdb.synth_db(max(1, nq - 12), mean=280, seed=100 + nq, degenerate_frac=0.01)]
    if nq >= 40:
        parts += [golden_set(n).db.subset(np.arange(4)) for n in ("globins4", "Pkinase", "fn3")]     # planted homologs (high scores, overflow)
    hdb = seqdb.concat(parts) if len(parts) > 1 else parts[0]
    db = engine.SeqDB.upload(ctx, hdb)
    got = engine.scan_msv_prefilter(ctx, ops, db, hdb.lengths, F1=F1)
    nflag = 0
    for o, g in zip(ops, got):
        prof = engine.Profile(ctx, o)
        sc, st = engine.ssvfilter_batch(ctx, prof, db)
        L_ = engine.lib()
        nullsc = np.array([L_.p7gpu_nullsc(int(L)) for L in hdb.lengths], np.float32)
        mu, lam = prof.evparam[0], prof.evparam[1]
        ok = st == 0
        passes = np.array([ok[k] and not (L_.p7gpu_msv_pvalue(float(sc[k]), float(nullsc[k]), mu, lam) > F1) for k in range(hdb.nseq)])
        expect = int(np.sum((hdb.lengths > 0) & (passes | ~ok)))
        assert int(g) == expect, (o.M, int(g), expect)
        r = engine.pipeline_run(ctx, prof, db, F1=F1, stop_after_msv=True)
        assert int(r.counters[2]) <= int(g)
        if int(r.counters[2]) > 0:
            assert g > 0
        nflag += int(g > 0)
        prof.close()
    if nq >= 40 and F1 >= 0.02:
        assert nflag >= 3
    db.close()


@pytest.mark.parametrize("name", ["globins4", "Pkinase", "tufa420_sprot"])
def test_five_bit_transfer_format_gives_identical_results(ctx, name):
    """p7gpu_pipeline_run_host5: residues cross PCIe as eight 5-bit codes in five bytes and are expanded on the device;
    counters, survivors and specials must be those of the byte path, bit for bit."""
    g = golden_set(name)
    prof = engine.Profile(ctx, g.oprofile)
    hdb = seqdb.concat([g.db, seqdb.synth_db(5000, mean=200, seed=3, degenerate_frac=0.02)])
    p5 = engine.pack5(hdb.residues)
    assert p5.shape[0] == (hdb.residues.shape[0] + 63) // 64 * 40 + 40
    a = engine.pipeline_run_host5(ctx, prof, p5, hdb)
    b = engine.pipeline_run_host(ctx, prof, hdb)
    assert list(a.counters) == list(b.counters)
    assert np.array_equal(a.survivors["seqidx"], b.survivors["seqidx"]) and len(b.survivors) > 0
    for k in ("usc", "vfsc", "fwdsc", "bcksc", "filtersc"):
        assert np.array_equal(bits(a.survivors[k]), bits(b.survivors[k])), k
    assert np.array_equal(bits(a.fwd_xmx), bits(b.fwd_xmx)) and np.array_equal(bits(a.bck_xmx), bits(b.bck_xmx))
    c = engine.pipeline_run_host5(ctx, prof, p5, hdb, stop_after_msv=True)
    d = engine.pipeline_run_host(ctx, prof, hdb, stop_after_msv=True)
    assert np.array_equal(c.msv_pass["seqidx"], d.msv_pass["seqidx"]) and np.array_equal(bits(c.msv_pass["usc"]), bits(d.msv_pass["usc"]))
    with pytest.raises(engine.P7GpuError):
        engine.pack5(np.array([255, 3, 29, 255], np.uint8))
    prof.close()


def test_streamed_search_is_independent_of_the_chunk_plan(ctx, monkeypatch):
    """The streamed search (bytes and 5-bit form) cut into many small chunks — every chunk boundary, the expansion kernels on
    their own stream, the copies running ahead — against one chunk and against the resident database."""
    g = golden_set("synth_M400")
    prof = engine.Profile(ctx, g.oprofile)
    hdb = seqdb.concat([g.db, seqdb.synth_db(60000, mean=300, seed=21, degenerate_frac=0.01)])     # ~18 MB of residues
    p5 = engine.pack5(hdb.residues)
    db = engine.SeqDB.upload(ctx, hdb)
    ref = engine.pipeline_run(ctx, prof, db, stop_after_msv=True)
    want = (list(ref.counters), ref.msv_pass["seqidx"].copy(), bits(ref.msv_pass["usc"]).copy())
    for chunk0, chunkmax, growth in ((1, 1, 0), (1, 3, 100), (2, 64, 33), (64, 64, 0)):
        monkeypatch.setenv("P7GPU_STREAM_CHUNK0_MB", str(chunk0)); monkeypatch.setenv("P7GPU_STREAM_CHUNKMAX_MB", str(chunkmax))
        monkeypatch.setenv("P7GPU_STREAM_GROWTH_PCT", str(growth))
        for r in (engine.pipeline_run_host5(ctx, prof, p5, hdb, stop_after_msv=True), engine.pipeline_run_host(ctx, prof, hdb, stop_after_msv=True)):
            assert list(r.counters) == want[0], (chunk0, chunkmax, growth)
            assert np.array_equal(r.msv_pass["seqidx"], want[1]) and np.array_equal(bits(r.msv_pass["usc"]), want[2])
            r.close()
    monkeypatch.setenv("P7GPU_STREAM_CHUNK0_MB", "1"); monkeypatch.setenv("P7GPU_STREAM_CHUNKMAX_MB", "1")
    full = engine.pipeline_run_host5(ctx, prof, p5, hdb)                      # 1 MB chunks through the whole pipeline
    whole = engine.pipeline_run(ctx, prof, db)
    assert list(full.counters) == list(whole.counters) and np.array_equal(full.survivors["seqidx"], whole.survivors["seqidx"])
    full.close(); whole.close(); ref.close(); db.close(); prof.close()


@pytest.mark.parametrize("name", ["synth_M400", "globins4", "Pkinase", "tufa420_sprot", "synth_M100"])
def test_device_stage_decisions_equal_host_decisions(ctx, name):
    """Default path: the bias and Viterbi P-value tests are float-threshold compares on the device (thresholds found and
    verified with the host libm).  P7GPU_HOST_DECISIONS=1 evaluates every P-value on the host in the reference's own
    expressions.  Both must give the same counters, the same survivors and bit-identical scores."""
    g = golden_set(name)
    prof = engine.Profile(ctx, g.oprofile)
    hdb = seqdb.concat([g.db, seqdb.synth_db(60000, mean=300, seed=13, degenerate_frac=0.002)])
    db = engine.SeqDB.upload(ctx, hdb)
    dev = engine.pipeline_run(ctx, prof, db, F3=1e-2)            # a looser F3 so that some survivors carry Viterbi scores
    os.environ["P7GPU_HOST_DECISIONS"] = "1"
    try:
        host = engine.pipeline_run(ctx, prof, db, F3=1e-2)
    finally:
        del os.environ["P7GPU_HOST_DECISIONS"]
    assert list(dev.counters) == list(host.counters)
    assert dev.stage_residues == host.stage_residues
    assert np.array_equal(dev.survivors["seqidx"], host.survivors["seqidx"])
    for k in ("usc", "vfsc", "fwdsc", "bcksc", "filtersc", "nullsc"):
        assert np.array_equal(bits(dev.survivors[k]), bits(host.survivors[k])), k
    assert np.array_equal(dev.survivors["vit_status"], host.survivors["vit_status"])
    assert int(dev.counters[3]) <= int(dev.counters[2]) and int(dev.counters[4]) < int(dev.counters[3])   # the filters removed something
    db.close(); prof.close()


def test_unsupported_model_length_is_rejected(ctx):
    """M beyond the slot geometry (64*65-1 = 4159 nodes) must fail loudly, not silently mis-score."""
    import ctypes
    from biobench2_b200.oprofile import OProfileView
    g = golden_set("synth_M33")
    v = g.oprofile.view()
    big = OProfileView()
    ctypes.memmove(ctypes.byref(big), ctypes.byref(v), ctypes.sizeof(OProfileView))
    big.M = 5000

    class Fake:
        bg_filter = None

        def view(self):
            return big
    with pytest.raises(engine.P7GpuError) as e:
        engine.Profile(ctx, Fake())
    assert e.value.code == engine.EINVAL


@pytest.mark.parametrize("name", ["globins4", "synth_M400", "tufa420_sprot"])
def test_bias_filter_score_is_exact_by_construction(gpu_ctx, name):
    """bias.cu certifies every (float) log((double) x) it takes; targets it cannot certify are recomputed with the host libm.
    With the certainty window widened from 2^-48 to 2^-16 most targets take the host path: filtersc must stay the reference's
    bit for bit, the pipeline counters must not move, and the host path must actually have run."""
    from conftest import golden_set, bits
    g = golden_set(name)
    prof = engine.Profile(gpu_ctx, g.oprofile); db = engine.SeqDB.upload(gpu_ctx, g.db)
    idx = np.nonzero(g.db.lengths > 0)[0].astype(np.int64)
    try:
        before = gpu_ctx.bias_recomputed
        f0 = engine.biasfilter_batch(gpu_ctx, prof, db, idx)
        assert np.array_equal(bits(f0), bits(g.vec["filtersc"][idx]))
        normal = gpu_ctx.bias_recomputed - before
        gpu_ctx.set_option("bias_window_log2", 16)
        f1 = engine.biasfilter_batch(gpu_ctx, prof, db, idx)
        forced = gpu_ctx.bias_recomputed - before - normal
        assert np.array_equal(bits(f1), bits(g.vec["filtersc"][idx]))
        assert forced > normal and forced >= len(idx) // 20, (normal, forced, len(idx))
        r1 = engine.pipeline_run(gpu_ctx, prof, db, want_specials=False)
        gpu_ctx.set_option("bias_window_log2", 0)
        r0 = engine.pipeline_run(gpu_ctx, prof, db, want_specials=False)
        assert list(r1.counters) == list(r0.counters) == list(g.vec["pli_counters"])
    finally:
        gpu_ctx.set_option("bias_window_log2", 0)
        db.close(); prof.close()


@pytest.mark.parametrize("name", ["globins4", "Pkinase", "synth_M400", "tufa420_sprot"])
def test_bias_warp_kernel_equals_thread_kernel(gpu_ctx, name):
    """bias.cu has a throughput form (one thread per target) and a latency form (one warp per target, the logarithms of 32 rows
    in parallel, added in row order): both must give the reference's filtersc bit for bit — on the golden targets and on 3000
    seeded ones of every length class (L = 1, L < 32, L = 32k, long), with the certainty window at its default and widened."""
    from conftest import golden_set, bits
    g = golden_set(name)
    prof = engine.Profile(gpu_ctx, g.oprofile)
    extra = seqdb.pack([np.full(L, (7 * L) % 20, np.uint8) for L in (1, 2, 31, 32, 33, 64, 65, 1000)])
    hdb = seqdb.concat([g.db, seqdb.synth_db(3000, mean=220, seed=9, degenerate_frac=0.02), extra])
    db = engine.SeqDB.upload(gpu_ctx, hdb)
    idx = np.nonzero(hdb.lengths > 0)[0].astype(np.int64)
    gi = np.nonzero(g.db.lengths > 0)[0].astype(np.int64)
    try:
        out = {}
        for mode in (0, 2):
            gpu_ctx.set_option("bias_warp", mode)
            out[mode] = engine.biasfilter_batch(gpu_ctx, prof, db, idx)
            assert np.array_equal(bits(engine.biasfilter_batch(gpu_ctx, prof, db, gi)), bits(g.vec["filtersc"][gi])), mode
        assert np.array_equal(bits(out[0]), bits(out[2]))
        gpu_ctx.set_option("bias_window_log2", 16)
        for mode in (0, 2):                                   # the host recomputation behind either form
            gpu_ctx.set_option("bias_warp", mode)
            before = gpu_ctx.bias_recomputed
            f = engine.biasfilter_batch(gpu_ctx, prof, db, idx)
            assert np.array_equal(bits(f), bits(out[0])) and gpu_ctx.bias_recomputed - before >= len(idx) // 20, mode
    finally:
        gpu_ctx.set_option("bias_window_log2", 0)
        gpu_ctx.set_option("bias_warp", 1)
        db.close(); prof.close()


@pytest.mark.gpu
def test_result_may_outlive_its_context():
    """A pipeline result hands its large host buffers back to its context's cache when it is destroyed; a host language may
    collect the context first.  p7gpu_result_destroy must notice (live-context registry) instead of touching freed memory."""
    g = golden_set("globins4")
    c = engine.Context(0)
    prof = engine.Profile(c, g.oprofile)
    db = engine.SeqDB.upload(c, g.db)
    r1 = engine.pipeline_run(c, prof, db, stop_after_msv=True)                     # keeps an MSV pass list drawn from the cache
    r2 = engine.pipeline_run(c, prof, db, want_specials=True, region_trigger=False)  # keeps specials drawn from the cache
    n1, n2 = len(r1.msv_pass), len(r2.survivors)
    assert n1 > 0 and n2 > 0
    db.close(); prof.close(); c.close()
    r1.close(); r2.close()                                                         # after the context: must not crash
    c2 = engine.Context(0)                                                         # and the library is still usable
    c2.close()


@pytest.mark.gpu
def test_prepared_length_table_gives_identical_results(ctx):
    """p7gpu_profile_prepare_lengths (scan mode): per-length constants for the target lengths only.  Same survivors, scores and
    counters as the dense table; a later database with other lengths falls back to the dense table by itself."""
    g = golden_set("Pkinase")
    db = engine.SeqDB.upload(ctx, g.db)
    dense = engine.Profile(ctx, g.oprofile)
    r0 = engine.pipeline_run(ctx, dense, db, want_specials=False)
    sparse = engine.Profile(ctx, g.oprofile)
    sparse.prepare_lengths(np.unique(g.db.lengths), F1=0.02)
    r1 = engine.pipeline_run(ctx, sparse, db, want_specials=False)
    def same(a, b):                                                  # field by field: Viterbi scores of eslERANGE targets are NaN
        return a.dtype == b.dtype and a.shape == b.shape and all(
            np.array_equal(a[f], b[f], equal_nan=a.dtype[f].kind == "f") for f in a.dtype.names)
    assert list(r0.counters) == list(r1.counters)
    assert same(r0.survivors, r1.survivors) and len(r0.survivors) > 0
    m0 = engine.pipeline_run(ctx, dense, db, stop_after_msv=True); m1 = engine.pipeline_run(ctx, sparse, db, stop_after_msv=True)
    assert same(m0.msv_pass, m1.msv_pass)
    other = seqdb.synth_db(300, mean=int(g.db.lengths.max()) + 50, seed=5)          # lengths the table was not prepared for
    odb = engine.SeqDB.upload(ctx, other)
    a = engine.pipeline_run(ctx, dense, odb, stop_after_msv=True); b = engine.pipeline_run(ctx, sparse, odb, stop_after_msv=True)
    assert same(a.msv_pass, b.msv_pass)
    for r in (r0, r1, m0, m1, a, b):
        r.close()
    odb.close(); db.close(); dense.close(); sparse.close()
