"""Packed digital target database (the GPU-side replacement for ESL_SQ_BLOCK hand-off).

Layout (include/p7gpu.h §4): `residues` holds each target's dsq[1..L] back to back with ONE sentinel
byte 255 before the first target and after every target, so `residues[offsets[s]-1 : offsets[s]+L+1]`
is a valid sentinel-framed ESL_DSQ (HMMER/easel/easel.h:308-313).  Digital codes follow the easel amino
alphabet "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~" (HMMER/easel/esl_alphabet.c:276): 0-19 canonical, 20 gap,
21-25 BJZOU, 26 X, 27 *, 28 ~.
"""
import numpy as np

AMINO_SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
SENTINEL = 255

# p7_AminoFrequencies (HMMER/src/hmmer.c:163): Swiss-Prot 50.8 background, the iid generator the
# reference's own benchmark drivers use (esl_rsq_xfIID, HMMER/easel/esl_randomseq.c:764-772).
AMINO_FREQ = np.array([
    0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092, 0.0594422,
    0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364, 0.0540687, 0.0673417,
    0.0114135, 0.0304133], dtype=np.float64)

_DIGIT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(AMINO_SYMS):
    _DIGIT[ord(_c)] = _i
    _DIGIT[ord(_c.lower())] = _i
_DIGIT[ord("_")] = 20
_DIGIT[ord(".")] = 20


def digitize(text):
    """Text -> residue codes (no sentinels).  Unknown symbols raise, like esl_abc_Digitize (eslEINVAL)."""
    a = _DIGIT[np.frombuffer(text.encode("ascii"), dtype=np.uint8)]
    if (a == 255).any():
        raise ValueError("invalid residue symbol in sequence")
    return a


class PackedDB:
    """Host-side packed database: residues (u8), offsets (i64), lengths (i32)."""

    def __init__(self, residues, offsets, lengths, names=None):
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        self.names = names

    @property
    def nseq(self):
        return int(self.lengths.shape[0])

    @property
    def nres(self):
        return int(self.lengths.astype(np.int64).sum())

    def dsq(self, s):
        """Sentinel-framed view of target s (L+2 bytes)."""
        o, L = int(self.offsets[s]), int(self.lengths[s])
        return self.residues[o - 1:o + L + 1]

    def subset(self, idx):
        return pack([self.residues[int(self.offsets[s]):int(self.offsets[s]) + int(self.lengths[s])] for s in idx])

    def shard(self, rank, world):
        """Contiguous range of targets balanced by residue count; returns (lo, hi)."""
        cum = np.concatenate([[0], np.cumsum(self.lengths.astype(np.int64) + 1)])
        total = cum[-1]
        lo = int(np.searchsorted(cum, total * rank // world, side="left"))
        hi = int(np.searchsorted(cum, total * (rank + 1) // world, side="left")) if rank + 1 < world else self.nseq
        return min(lo, self.nseq), min(hi, self.nseq)


def pack(seqs, names=None, align=1):
    """List of uint8 residue arrays -> PackedDB.  align=16 starts every target on a 16-byte boundary (include/p7gpu.h §4)."""
    lengths = np.array([len(s) for s in seqs], dtype=np.int32)
    offsets = np.zeros(len(seqs), dtype=np.int64)
    span = (lengths.astype(np.int64) + 1 + align - 1) // align * align
    first = align if align > 1 else 1
    total = int(span.sum()) + first
    residues = np.full(total, SENTINEL, dtype=np.uint8)
    pos = first
    for j, s in enumerate(seqs):
        offsets[j] = pos
        residues[pos:pos + len(s)] = s
        pos += int(span[j])
    return PackedDB(residues, offsets, lengths, names)


def concat(dbs):
    """Several PackedDBs one after the other as one PackedDB."""
    return pack([d.residues[int(d.offsets[s]):int(d.offsets[s]) + int(d.lengths[s])] for d in dbs for s in range(d.nseq)])


def pack_from_lengths(lengths, flat_residues):
    """Vectorised pack: `flat_residues` are all residues back to back (sum(lengths) bytes)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    n = lengths.shape[0]
    l64 = lengths.astype(np.int64)
    starts_flat = np.concatenate([[0], np.cumsum(l64)[:-1]]) if n else np.zeros(0, np.int64)
    offsets = starts_flat + np.arange(1, n + 1, dtype=np.int64)
    total = int(l64.sum()) + n + 1
    residues = np.full(total, SENTINEL, dtype=np.uint8)
    # position of flat residue r of sequence s = r + s + 1
    seq_of = np.repeat(np.arange(n, dtype=np.int64), l64)
    residues[np.arange(flat_residues.shape[0], dtype=np.int64) + seq_of + 1] = flat_residues
    return PackedDB(residues, offsets, lengths)


def synth_lengths(n, mean=350, rng=None, lmin=25, lmax=5000):
    """Seeded Swiss-Prot-like log-normal length distribution clipped to [lmin, lmax]."""
    rng = rng or np.random.default_rng(42)
    sigma = 0.6
    mu = np.log(mean) - 0.5 * sigma * sigma
    L = np.rint(rng.lognormal(mu, sigma, size=n)).astype(np.int64)
    return np.clip(L, lmin, lmax).astype(np.int32)


def synth_db(n, mean=350, seed=42, fixed_length=None, freq=None, degenerate_frac=0.0):
    """iid background-frequency protein database (numpy, CPU; for tests and small runs)."""
    rng = np.random.default_rng(seed)
    lengths = (np.full(n, fixed_length, dtype=np.int32) if fixed_length is not None
               else synth_lengths(n, mean, rng))
    f = AMINO_FREQ if freq is None else np.asarray(freq, dtype=np.float64)
    cdf = np.cumsum(f / f.sum())
    tot = int(lengths.astype(np.int64).sum())
    flat = np.searchsorted(cdf, rng.random(tot), side="right").astype(np.uint8)
    np.minimum(flat, 19, out=flat)
    if degenerate_frac > 0:
        m = rng.random(tot) < degenerate_frac
        flat[m] = rng.choice(np.array([21, 22, 23, 24, 25, 26], dtype=np.uint8), size=int(m.sum()))
    return pack_from_lengths(lengths, flat)


def read_fasta(path, max_seqs=None):
    """Minimal FASTA reader -> PackedDB (names kept).  Host-side convenience for tests/fixtures."""
    names, seqs, cur = [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if names:
                    seqs.append(digitize("".join(cur)))
                    if max_seqs and len(seqs) >= max_seqs:
                        names = names[:len(seqs)]
                        cur = None
                        break
                names.append(line[1:].split()[0])
                cur = []
            else:
                cur.append(line.strip())
    if cur is not None and names:
        seqs.append(digitize("".join(cur)))
    db = pack(seqs, names)
    return db
