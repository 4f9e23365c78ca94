"""The packed on-disk database (csrc/dbfile.cu): the parallel FASTA digitiser against the reference's own
reader (esl_sqfile_OpenDigital + esl_sqio_Read through oracle/_ref) — residues, lengths, names, descriptions and error
behaviour — and the mapped file as a search input.  Host code only: runs without a GPU (the library loads; no kernel runs)."""
import os
import numpy as np
import pytest

from biobench2_b200 import engine, seqdb
from oracle import ref

SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
TRICKY = (
    "\n  \n>sp|P1|ONE first protein  with  spaces \nACDEFGHIKLMNPQRSTVWY\nacdefghiklmnpqrstvwy\n"
    ">two\tdesc after tab\r\nBJZOUX*~-_.\r\nAC DE\tFG\r\n\r\n"
    ">empty_record\n>after_empty no residues before me\nMKV\n"
    ">gt_inside desc with > sign\nMMM>inline_header starts mid-line\nKKK\n"
    ">  spaced_name   trailing  \nWWW")                        # no newline at the end of the file


def ref_or_skip():
    if not ref.available():
        pytest.skip("oracle/_ref/libp7ref.so not built")


def roundtrip(tmp_path, text, threads=0, name="t.fa", align=1):
    fa = tmp_path / name; fa.write_text(text, newline="")
    out = tmp_path / (name + ".a%d.p7db" % align)
    n, r = engine.DbFile.from_fasta(fa, out, threads, align)
    return str(fa), engine.DbFile(out), n, r


def assert_same_as_reference(fa, d):
    st, rdb, err = ref.read_fasta(fa)
    assert st == 0, err
    p = d.packed()
    assert d.nseq == len(rdb["lengths"]) and np.array_equal(p.lengths, rdb["lengths"])
    assert np.array_equal(p.offsets, rdb["offsets"]) and np.array_equal(p.residues, rdb["residues"])
    assert [(d.name(s), d.desc(s)) for s in range(d.nseq)] == rdb["names"]
    assert d.nres == int(rdb["lengths"].sum()) and d.max_length == int(rdb["lengths"].max(initial=0))


def test_tricky_fasta_same_as_reference_reader(tmp_path):
    ref_or_skip()
    fa, d, n, r = roundtrip(tmp_path, TRICKY)
    assert n == 7
    assert_same_as_reference(fa, d)
    assert (d.name(0), d.desc(0)) == ("sp|P1|ONE", "first protein  with  spaces ")
    assert d.packed().lengths[2] == 0                          # header directly followed by the next header
    assert d.name(5) == "inline_header"                        # '>' in the middle of a sequence line starts a record
    d.close()


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_large_random_fasta_any_thread_count(tmp_path, threads):
    ref_or_skip()
    rng = np.random.default_rng(5)
    recs = []
    for i in range(3000):
        L = int(rng.integers(0, 700))
        s = "".join(rng.choice(list(SYMS[:20] + "XBZ"), L))
        w = int(rng.integers(20, 90))
        body = "\n".join(s[k:k + w] for k in range(0, L, w))
        recs.append(">seq%d some description %d > with a sign\n%s\n" % (i, i, body))
    fa, d, n, r = roundtrip(tmp_path, "".join(recs), threads)
    assert n == 3000
    assert_same_as_reference(fa, d)
    d.close()


@pytest.mark.parametrize("text,needle", [
    (">a\nACD1EF\n", "illegal character 1"),
    (">a\nAC\xe9DE\n", "non-ASCII"),
    ("ACDE\n>a\nAC\n", "expected FASTA to start with >"),
    (">\nACDE\n", "no FASTA name found"),
    (">a\nACDE\n>b desc only", "Premature EOF"),
])
def test_errors_like_the_reference(tmp_path, text, needle):
    ref_or_skip()
    fa = tmp_path / "bad.fa"; fa.write_bytes(text.encode("latin-1"))
    with pytest.raises(engine.P7GpuError) as e:
        engine.DbFile.from_fasta(fa, tmp_path / "bad.p7db")
    assert needle in str(e.value)
    st, _, err = ref.read_fasta(str(fa))
    assert st != 0 and needle in err                            # the reference refuses the same file with the same words


def test_aligned_layout_holds_the_same_targets(tmp_path):
    """align=16 (the default): every target starts on a 16-byte boundary, gaps are sentinels, content unchanged."""
    ref_or_skip()
    fa, d1, n1, r1 = roundtrip(tmp_path, TRICKY, align=1)
    _, d16, n16, r16 = roundtrip(tmp_path, TRICKY, align=16)
    p1, p16 = d1.packed(), d16.packed()
    assert n1 == n16 and r1 == r16 and np.array_equal(p1.lengths, p16.lengths)
    assert (p16.offsets % 16 == 0).all() and p16.residues.shape[0] % 16 == 0
    covered = np.zeros(p16.residues.shape[0], bool)
    for s in range(n1):
        a = p1.residues[p1.offsets[s]: p1.offsets[s] + p1.lengths[s]]
        b = p16.residues[p16.offsets[s]: p16.offsets[s] + p16.lengths[s]]
        assert np.array_equal(a, b) and p16.residues[p16.offsets[s] - 1] == 255 and p16.residues[p16.offsets[s] + p16.lengths[s]] == 255
        covered[p16.offsets[s]: p16.offsets[s] + p16.lengths[s]] = True
        assert (d1.name(s), d1.desc(s)) == (d16.name(s), d16.desc(s))
    assert (p16.residues[~covered] == 255).all()
    d1.close(); d16.close()


def test_empty_file_and_bad_magic(tmp_path):
    fa, d, n, r = roundtrip(tmp_path, "")
    assert n == 0 and d.nseq == 0 and d.nbytes == 1
    d.close()
    junk = tmp_path / "junk.p7db"; junk.write_bytes(b"x" * 4096)
    with pytest.raises(engine.P7GpuError):
        engine.DbFile(junk)


@pytest.mark.gpu
def test_search_from_mapped_file_equals_search_from_memory(tmp_path):
    from conftest import golden_set
    g = golden_set("globins4")
    lines = []
    for s in range(g.db.nseq):
        o, L = int(g.db.offsets[s]), int(g.db.lengths[s])
        lines.append(">t%d\n%s\n" % (s, "".join(SYMS[c] for c in g.db.residues[o:o + L])))
    fa, d, n, r = roundtrip(tmp_path, "".join(lines), align=16)
    assert n == g.db.nseq and np.array_equal(d.packed().lengths, g.db.lengths)
    ctx = engine.Context(0); prof = engine.Profile(ctx, g.oprofile)
    d.pin()
    a = engine.pipeline_run_file(ctx, prof, d)
    b = engine.pipeline_run_host(ctx, prof, g.db)
    assert list(a.counters) == list(b.counters) == list(g.vec["pli_counters"])
    assert np.array_equal(a.survivors["seqidx"], b.survivors["seqidx"])
    assert np.array_equal(a.fwd_xmx.view(np.uint32), b.fwd_xmx.view(np.uint32))
    a.close(); b.close(); prof.close(); d.close(); ctx.close()


def test_random_well_formed_fasta_property(tmp_path):
    """Property test (hypothesis): any well-formed FASTA text — mixed case, every alphabet symbol, '_' and '.' gaps, blank
    lines, CRLF, ragged line widths, spaces and tabs inside sequence lines, '>' inside descriptions — is digitised exactly
    as the reference's reader digitises it, at any thread count."""
    ref_or_skip()
    from hypothesis import given, settings, strategies as st, HealthCheck

    sym = st.sampled_from(list(SYMS + SYMS.lower()[:20] + "_."))
    name = st.text(alphabet=st.characters(min_codepoint=33, max_codepoint=126), min_size=1, max_size=12)
    desc = st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), min_size=0, max_size=30)
    eol = st.sampled_from(["\n", "\r\n"])
    filler = st.sampled_from(["", " ", "\t", "  "])

    @st.composite
    def record(draw):
        seq = "".join(draw(st.lists(sym, min_size=0, max_size=120)))
        width = draw(st.integers(5, 70))
        lines = [seq[k:k + width] for k in range(0, len(seq), width)]
        body = "".join(draw(filler) + ln + draw(filler) + draw(eol) + ("\n" if draw(st.integers(0, 9)) == 0 else "") for ln in lines)
        return ">" + draw(st.sampled_from(["", " "])) + draw(name) + draw(st.sampled_from([" ", "\t", "  "])) + draw(desc) + draw(eol) + body

    counter = [0]

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(st.lists(record(), min_size=1, max_size=12), st.integers(1, 5))
    def check(recs, threads):
        text = "".join(recs)
        if not text.endswith("\n") or text.rstrip("\r\n").rsplit("\n", 1)[-1].startswith(">"):
            text += "A\n"                                    # a header at the very end of the file is an error for both readers
        counter[0] += 1
        fa = tmp_path / ("h%d.fa" % counter[0]); fa.write_bytes(text.encode("ascii"))
        out = tmp_path / ("h%d.p7db" % counter[0])
        n, r = engine.DbFile.from_fasta(fa, out, threads, 1)
        d = engine.DbFile(out)
        try:
            assert_same_as_reference(str(fa), d)
        finally:
            d.close()

    check()
