"""ctypes binding of oracle/_ref/libp7ref.so (reference HMMER 3.1b2 + oracle/ref_harness.c).

TEST INFRASTRUCTURE — see oracle/__init__.py.  The shared object is built in the authoring container
by `make -C oracle ref` and travels to the GPU box inside the repo snapshot; /root/reference itself is
never read at run time.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libp7ref.so")
HMMSEARCH = os.path.join(_HERE, "_ref", "hmmsearch")


def available():
    return os.path.exists(LIB_PATH)


class OProfileView(C.Structure):
    """Mirror of p7gpu_oprofile_view (include/p7gpu.h)."""
    _fields_ = [
        ("M", C.c_int32), ("Kp", C.c_int32), ("K", C.c_int32), ("mode", C.c_int32), ("nj", C.c_float),
        ("rbv", C.c_void_p), ("sbv", C.c_void_p),
        ("tbm_b", C.c_uint8), ("tec_b", C.c_uint8), ("tjb_b", C.c_uint8), ("base_b", C.c_uint8), ("bias_b", C.c_uint8),
        ("scale_b", C.c_float),
        ("rwv", C.c_void_p), ("twv", C.c_void_p),
        ("xw", (C.c_int16 * 2) * 4), ("scale_w", C.c_float), ("base_w", C.c_int16), ("ddbound_w", C.c_int16),
        ("rfv", C.c_void_p), ("tfv", C.c_void_p),
        ("xf", (C.c_float * 2) * 4),
        ("evparam", C.c_float * 6), ("compo", C.c_float * 20),
    ]


class RefHit(C.Structure):
    _fields_ = [("seqidx", C.c_int64), ("score", C.c_float), ("pre_score", C.c_float), ("sum_score", C.c_float),
                ("lnP", C.c_double), ("ndom", C.c_int32), ("nreported", C.c_int32), ("nincluded", C.c_int32),
                ("flags", C.c_int32), ("best_dom_score", C.c_float), ("best_dom_ienv", C.c_int32),
                ("best_dom_jenv", C.c_int32)]


_lib = None


def read_fasta(path, cap_bytes=1 << 26, cap_seqs=1 << 20, want_names=True):
    """The reference's own digital FASTA reader -> (status, PackedDB-like dict, error text)."""
    res = np.zeros(cap_bytes, np.uint8); off = np.zeros(cap_seqs, np.int64); ln = np.zeros(cap_seqs, np.int32)
    names = C.create_string_buffer(1 << 22); err = C.create_string_buffer(512)
    n = C.c_int64(0); nb = C.c_int64(0)
    L = lib()
    L.p7ref_read_fasta.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    st = L.p7ref_read_fasta(path.encode(), _p(res), cap_bytes, _p(off), _p(ln), cap_seqs, C.byref(n), C.byref(nb),
                            names if want_names else None, 1 << 22, err, 512)
    parts = names.raw.split(b"\0")
    nm = [(parts[2 * i].decode(), parts[2 * i + 1].decode()) for i in range(n.value)] if (st == 0 and want_names) else []
    return st, dict(residues=res[:nb.value].copy(), offsets=off[:n.value].copy(), lengths=ln[:n.value].copy(), names=nm), err.value.decode("latin-1")


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libp7ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.p7ref_open_hmm.restype = C.c_void_p
        L.p7ref_open_hmm.argtypes = [C.c_char_p]
        L.p7ref_sample_hmm.restype = C.c_void_p
        L.p7ref_sample_hmm.argtypes = [C.c_int, C.c_int, C.c_int]
        L.p7ref_write_hmm.argtypes = [C.c_void_p, C.c_char_p]
        L.p7ref_close.argtypes = [C.c_void_p]
        L.p7ref_M.argtypes = [C.c_void_p]
        L.p7ref_name.restype = C.c_char_p
        L.p7ref_name.argtypes = [C.c_void_p]
        L.p7ref_get_view.argtypes = [C.c_void_p, C.POINTER(OProfileView)]
        L.p7ref_emit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.p7ref_digitize.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.p7ref_background.argtypes = [C.c_void_p, C.c_void_p]
        L.p7ref_get_filter_hmm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.p7ref_omega.restype = C.c_float
        L.p7ref_omega.argtypes = [C.c_void_p]
        for nm in ("p7ref_ssv", "p7ref_msv", "p7ref_vit"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.p7ref_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_float)]
        L.p7ref_domain_decoding.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.p7ref_fwdbck.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.p7ref_nullsc.restype = C.c_float
        L.p7ref_nullsc.argtypes = [C.c_void_p, C.c_int]
        L.p7ref_filtersc.restype = C.c_float
        L.p7ref_filtersc.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.p7ref_length_scalars.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_float), C.POINTER(C.c_float)]
        for nm in ("p7ref_gumbel_surv", "p7ref_exp_surv"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_double] * 3
        L.p7ref_stage_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p]
        L.p7ref_stage_batch_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_void_p]
        L.p7ref_pipeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_char_p]
        L.p7ref_finish_from_specials.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_char_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


STAGES = {"ssv": 0, "msv": 1, "vit": 2, "fwd": 3}


class RefProfile:
    """A reference P7_HMM + P7_BG + P7_PROFILE + P7_OPROFILE bundle."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference harness returned NULL")
        self.h = C.c_void_p(handle)

    @classmethod
    def open(cls, path):
        return cls(lib().p7ref_open_hmm(path.encode()))

    @classmethod
    def sample(cls, M, seed=42, calibrate=True):
        return cls(lib().p7ref_sample_hmm(M, seed, 1 if calibrate else 0))

    def close(self):
        if self.h:
            lib().p7ref_close(self.h)
            self.h = None

    @property
    def M(self):
        return lib().p7ref_M(self.h)

    @property
    def name(self):
        return lib().p7ref_name(self.h).decode()

    def write_hmm(self, path):
        return lib().p7ref_write_hmm(self.h, path.encode())

    def view(self):
        v = OProfileView()
        lib().p7ref_get_view(self.h, C.byref(v))
        return v

    def digitize(self, text):
        dsq = np.zeros(len(text) + 2, dtype=np.uint8)
        st = lib().p7ref_digitize(self.h, text.encode(), _p(dsq))
        if st != 0:
            raise ValueError("esl_abc_Digitize status %d" % st)
        return dsq

    def emit(self, seed, Lcfg=350, maxL=20000):
        """Sample a homolog from the profile; returns residue codes (no sentinels) or None."""
        buf = np.zeros(maxL + 2, np.uint8)
        n = lib().p7ref_emit(self.h, seed, Lcfg, _p(buf), maxL)
        return None if n < 0 else buf[1:n + 1].copy()

    def background(self):
        f = np.zeros(20, dtype=np.float32)
        lib().p7ref_background(self.h, _p(f))
        return f

    def filter_hmm(self):
        t = np.zeros(6, np.float32); pi = np.zeros(3, np.float32); eo = np.zeros(29 * 2, np.float32)
        lib().p7ref_get_filter_hmm(self.h, _p(t), _p(pi), _p(eo))
        return t.reshape(2, 3), pi, eo.reshape(29, 2)

    def omega(self):
        return lib().p7ref_omega(self.h)

    def _one(self, fn, dsq, L):
        sc = C.c_float(0)
        st = fn(self.h, _p(dsq), L, C.byref(sc))
        return st, np.float32(sc.value)

    def ssv(self, dsq, L): return self._one(lib().p7ref_ssv, dsq, L)
    def msv(self, dsq, L): return self._one(lib().p7ref_msv, dsq, L)
    def vit(self, dsq, L): return self._one(lib().p7ref_vit, dsq, L)

    def fwd(self, dsq, L, want_xmx=True):
        sc = C.c_float(0); tot = C.c_float(0)
        xmx = np.zeros((L + 1, 6), np.float32) if want_xmx else None
        st = lib().p7ref_fwd(self.h, _p(dsq), L, C.byref(sc), _p(xmx), C.byref(tot))
        return st, np.float32(sc.value), xmx, np.float32(tot.value)

    def fwdbck(self, dsq, L):
        fsc = C.c_float(0); bsc = C.c_float(0); ft = C.c_float(0); bt = C.c_float(0); own = C.c_int(0)
        fx = np.zeros((L + 1, 6), np.float32); bx = np.zeros((L + 1, 6), np.float32)
        st = lib().p7ref_fwdbck(self.h, _p(dsq), L, C.byref(fsc), C.byref(bsc), _p(fx), _p(bx),
                                C.byref(ft), C.byref(bt), C.byref(own))
        return dict(status=st, fwdsc=np.float32(fsc.value), bcksc=np.float32(bsc.value), fwd_xmx=fx, bck_xmx=bx,
                    fwd_totscale=np.float32(ft.value), bck_totscale=np.float32(bt.value), bck_own_scales=own.value)

    def domain_decoding(self, L, fwd_xmx, bck_xmx, own_scales=0):
        """The reference's p7_DomainDecoding on the given specials + the region walk of p7_domaindef_ByPosteriorHeuristics
        -> dict(status, btot, etot, mocc, nregions, triggered)."""
        fx = np.ascontiguousarray(fwd_xmx, np.float32); bx = np.ascontiguousarray(bck_xmx, np.float32)
        bt = np.zeros(L + 1, np.float32); et = np.zeros(L + 1, np.float32); mo = np.zeros(L + 1, np.float32)
        nr = C.c_int(0); tr = C.c_int(0)
        st = lib().p7ref_domain_decoding(self.h, L, _p(fx), _p(bx), int(own_scales), _p(bt), _p(et), _p(mo), C.byref(nr), C.byref(tr))
        return dict(status=st, btot=bt, etot=et, mocc=mo, nregions=nr.value, triggered=tr.value)

    def nullsc(self, L):
        return np.float32(lib().p7ref_nullsc(self.h, L))

    def filtersc(self, dsq, L):
        return np.float32(lib().p7ref_filtersc(self.h, _p(dsq), L))

    def length_scalars(self, L):
        a = C.c_int(0); b = C.c_int(0); c = C.c_float(0); d = C.c_float(0)
        lib().p7ref_length_scalars(self.h, L, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, np.float32(c.value), np.float32(d.value)

    def stage_batch(self, stage, residues, offsets, lengths, idx=None, threads=1):
        n = len(idx) if idx is not None else len(lengths)
        sc = np.zeros(n, np.float32); st = np.zeros(n, np.int32)
        if threads > 1 and idx is None:
            lib().p7ref_stage_batch_mt(self.h, STAGES[stage], threads, _p(residues), _p(offsets), _p(lengths), n, _p(sc), _p(st))
        else:
            if idx is not None:
                idx = np.ascontiguousarray(idx, np.int64)
            lib().p7ref_stage_batch(self.h, STAGES[stage], _p(residues), _p(offsets), _p(lengths), _p(idx), n, _p(sc), _p(st))
        return sc, st

    def pipeline(self, residues, offsets, lengths, F1=0.02, F2=1e-3, F3=1e-5, do_max=False, nobias=False,
                 maxhits=100000, tblout=None):
        n = len(lengths)
        counters = np.zeros(6, np.uint64)
        hits = (RefHit * maxhits)()
        nh = C.c_int64(0)
        lib().p7ref_pipeline(self.h, _p(residues), _p(offsets), _p(lengths), n, F1, F2, F3, int(do_max), int(nobias),
                             _p(counters), C.cast(hits, C.c_void_p), maxhits, C.byref(nh),
                             tblout.encode() if tblout else None)
        return counters, [hits[i] for i in range(nh.value)]

    def finish_from_specials(self, residues, offsets, lengths, seqidx, fwdsc, fwd_xmx, bck_xmx, xmx_offsets,
                             fwd_tot, bck_tot, nseqs_total, nres_total, do_max=False, nobias=False,
                             maxhits=100000, tblout=None):
        hits = (RefHit * maxhits)()
        nh = C.c_int64(0)
        seqidx = np.ascontiguousarray(seqidx, np.int64); fwdsc = np.ascontiguousarray(fwdsc, np.float32)
        xmx_offsets = np.ascontiguousarray(xmx_offsets, np.int64)
        fwd_tot = np.ascontiguousarray(fwd_tot, np.float32); bck_tot = np.ascontiguousarray(bck_tot, np.float32)
        lib().p7ref_finish_from_specials(self.h, _p(residues), _p(offsets), _p(lengths), len(seqidx), _p(seqidx),
                                         _p(fwdsc), _p(fwd_xmx), _p(bck_xmx), _p(xmx_offsets), _p(fwd_tot), _p(bck_tot),
                                         nseqs_total, nres_total, int(do_max), int(nobias),
                                         C.cast(hits, C.c_void_p), maxhits, C.byref(nh),
                                         tblout.encode() if tblout else None)
        return [hits[i] for i in range(nh.value)]
