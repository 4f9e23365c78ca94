"""ctypes binding of oracle/liboracle.so — our plain-C restatement (oracle/p7_oracle.c).

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, f) for f in ("p7_oracle.c", "p7_fbx_oracle.c", "p7_oracle.h")]
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
            build()
        L = C.CDLL(LIB_PATH)
        L.p7o_profile_from_view.restype = C.c_void_p
        L.p7o_profile_from_view.argtypes = [C.c_void_p]
        L.p7o_profile_free.argtypes = [C.c_void_p]
        L.p7o_tjb_b.restype = C.c_uint8
        L.p7o_tjb_b.argtypes = [C.c_void_p, C.c_int]
        L.p7o_xw_move.restype = C.c_int16
        L.p7o_xw_move.argtypes = [C.c_void_p, C.c_int]
        L.p7o_xf_ncj.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.p7o_nullsc.restype = C.c_float
        L.p7o_nullsc.argtypes = [C.c_int]
        for nm in ("p7o_ssvfilter", "p7o_msvfilter", "p7o_vitfilter", "p7o_vitfilter_alwaysdd"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.p7o_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.p7o_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                   C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.p7o_forward_striped.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.p7o_backward_striped.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                           C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.p7o_filtersc.restype = C.c_float
        L.p7o_domain_decoding.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.p7o_filtersc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for nm in ("p7o_gumbel_surv", "p7o_exp_surv"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [C.c_double] * 3
        L.p7o_stage_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p]
        L.p7o_pipeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


STAGES = {"ssv": 0, "msv": 1, "vit": 2, "fwd": 3}


class OracleProfile:
    """De-striped profile built from a p7gpu_oprofile_view (any ctypes Structure with that layout)."""

    def __init__(self, view, keepalive=None):
        self._keep = (view, keepalive)
        self.h = C.c_void_p(lib().p7o_profile_from_view(C.byref(view)))
        self.M = view.M
        self.evparam = [float(x) for x in view.evparam]

    def close(self):
        if self.h:
            lib().p7o_profile_free(self.h)
            self.h = None

    def _one(self, fn, dsq, L):
        sc = C.c_float(0)
        st = fn(self.h, _p(dsq), L, C.byref(sc))
        return st, np.float32(sc.value)

    def ssv(self, dsq, L): return self._one(lib().p7o_ssvfilter, dsq, L)
    def msv(self, dsq, L): return self._one(lib().p7o_msvfilter, dsq, L)
    def vit(self, dsq, L): return self._one(lib().p7o_vitfilter, dsq, L)
    def vit_alwaysdd(self, dsq, L): return self._one(lib().p7o_vitfilter_alwaysdd, dsq, L)

    def fwd(self, dsq, L, want_xmx=True):
        sc = C.c_float(0); tot = C.c_float(0)
        xmx = np.zeros((L + 1, 6), np.float32) if want_xmx else None
        st = lib().p7o_forward(self.h, _p(dsq), L, _p(xmx), C.byref(tot), C.byref(sc))
        return st, np.float32(sc.value), xmx, np.float32(tot.value)

    def bwd(self, dsq, L, fwd_xmx):
        sc = C.c_float(0); tot = C.c_float(0); own = C.c_int(0)
        xmx = np.zeros((L + 1, 6), np.float32)
        fwd_xmx = np.ascontiguousarray(fwd_xmx, np.float32)
        st = lib().p7o_backward(self.h, _p(dsq), L, _p(fwd_xmx), _p(xmx), C.byref(tot), C.byref(own), C.byref(sc))
        return st, np.float32(sc.value), xmx, np.float32(tot.value), own.value

    # the two parsers in the reference's own (striped) order of float operations: bit-exact (oracle/p7_fbx_oracle.c)
    def fwd_striped(self, dsq, L):
        sc = C.c_float(0); tot = C.c_float(0)
        xmx = np.zeros((L + 1, 6), np.float32)
        st = lib().p7o_forward_striped(C.byref(self._keep[0]), _p(dsq), L, _p(xmx), C.byref(tot), C.byref(sc))
        return st, np.float32(sc.value), xmx, np.float32(tot.value)

    def bwd_striped(self, dsq, L, fwd_xmx):
        sc = C.c_float(0); tot = C.c_float(0); own = C.c_int(0)
        xmx = np.zeros((L + 1, 6), np.float32)
        fwd_xmx = np.ascontiguousarray(fwd_xmx, np.float32)
        st = lib().p7o_backward_striped(C.byref(self._keep[0]), _p(dsq), L, _p(fwd_xmx), _p(xmx), C.byref(tot), C.byref(own), C.byref(sc))
        return st, np.float32(sc.value), xmx, np.float32(tot.value), own.value

    def domain_decoding(self, L, fwd_xmx, bck_xmx, own_scales=0):
        """-> dict(status, btot, etot, mocc [(L+1) each], mocc_max, triggered)"""
        fx = np.ascontiguousarray(fwd_xmx, np.float32); bx = np.ascontiguousarray(bck_xmx, np.float32)
        bt = np.zeros(L + 1, np.float32); et = np.zeros(L + 1, np.float32); mo = np.zeros(L + 1, np.float32)
        mx = C.c_float(0); tr = C.c_int(0)
        st = lib().p7o_domain_decoding(self.h, L, _p(fx), _p(bx), int(own_scales), _p(bt), _p(et), _p(mo), C.byref(mx), C.byref(tr))
        return dict(status=st, btot=bt, etot=et, mocc=mo, mocc_max=np.float32(mx.value), triggered=tr.value)

    def tjb_b(self, L): return lib().p7o_tjb_b(self.h, L)
    def xw_move(self, L): return lib().p7o_xw_move(self.h, L)

    def xf_ncj(self, L):
        a = C.c_float(0); b = C.c_float(0)
        lib().p7o_xf_ncj(self.h, L, C.byref(a), C.byref(b))
        return np.float32(a.value), np.float32(b.value)

    def stage_batch(self, stage, db, idx=None):
        n = len(idx) if idx is not None else db.nseq
        sc = np.zeros(n, np.float32); st = np.zeros(n, np.int32)
        if idx is not None:
            idx = np.ascontiguousarray(idx, np.int64)
        lib().p7o_stage_batch(self.h, STAGES[stage], _p(db.residues), _p(db.offsets), _p(db.lengths), _p(idx), n, _p(sc), _p(st))
        return sc, st

    def pipeline(self, db, filter_hmm, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True):
        t, pi, eo = [np.ascontiguousarray(a, np.float32) for a in filter_hmm]
        n = db.nseq
        counters = np.zeros(6, np.uint64); stage = np.zeros(n, np.uint8)
        usc = np.zeros(n, np.float32); vf = np.zeros(n, np.float32); fw = np.zeros(n, np.float32); fs = np.zeros(n, np.float32)
        lib().p7o_pipeline(self.h, _p(t), _p(pi), _p(eo), _p(db.residues), _p(db.offsets), _p(db.lengths), n,
                           F1, F2, F3, int(do_biasfilter), _p(counters), _p(stage), _p(usc), _p(vf), _p(fw), _p(fs))
        return dict(counters=counters, stage=stage, usc=usc, vfsc=vf, fwdsc=fw, filtersc=fs)


def nullsc(L):
    return np.float32(lib().p7o_nullsc(L))


def filtersc(filter_hmm, dsq, L):
    t, pi, eo = [np.ascontiguousarray(a, np.float32) for a in filter_hmm]
    return np.float32(lib().p7o_filtersc(_p(t), _p(pi), _p(eo), _p(dsq), L))


def gumbel_surv(x, mu, lam): return lib().p7o_gumbel_surv(x, mu, lam)
def exp_surv(x, mu, lam): return lib().p7o_exp_surv(x, mu, lam)
