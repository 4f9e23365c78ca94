"""P7_OPROFILE hand-over: the striped arrays exactly as the reference's impl_sse keeps them.

In an integrated build the reference's host code (p7_hmmfile_Read -> p7_ProfileConfig ->
p7_oprofile_Convert, HMMER/src/hmmsearch.c:443-503) produces the P7_OPROFILE and a 20-line stub fills a
p7gpu_oprofile_view from it.  In this repository the same bytes travel as an .npz
"oprofile dump" (tests/golden/*.npz, written by tests/golden/make_golden.py from the real reference),
which this module loads back into a view.  Nothing here computes scores.
"""
import ctypes as C
import numpy as np


class OProfileView(C.Structure):
    """ctypes mirror of p7gpu_oprofile_view (include/p7gpu.h)."""
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


def nq(M, w):
    """p7O_NQB/NQW/NQF (HMMER/src/impl_sse/impl_sse.h:25-27)."""
    return max(2, (M - 1) // w + 1)


_ARRAYS = (("rbv", np.uint8), ("sbv", np.uint8), ("rwv", np.int16), ("twv", np.int16), ("rfv", np.float32), ("tfv", np.float32))
_SCALARS = ("M", "Kp", "K", "mode", "nj", "tbm_b", "tec_b", "tjb_b", "base_b", "bias_b", "scale_b", "scale_w", "base_w", "ddbound_w")


class OProfile:
    """Host copy of a P7_OPROFILE's score tables + scalars; `.view()` gives the C-ABI struct."""

    def __init__(self, fields, name="", bg_filter=None):
        self.f = fields
        self.name = name
        self.bg_filter = bg_filter  # (t[2,3], pi[3], eo[29,2], omega) of the bias-filter HMM, or None
        self._view = None

    @property
    def M(self):
        return int(self.f["M"])

    @property
    def evparam(self):
        return [float(x) for x in self.f["evparam"]]

    @classmethod
    def from_view(cls, v, name="", bg_filter=None):
        """Copy the arrays out of a live view (pointers into the reference's memory)."""
        M, Kp = v.M, v.Kp
        sizes = {"rbv": Kp * nq(M, 16) * 16, "sbv": Kp * (nq(M, 16) + 17) * 16, "rwv": Kp * nq(M, 8) * 8,
                 "twv": 8 * nq(M, 8) * 8, "rfv": Kp * nq(M, 4) * 4, "tfv": 8 * nq(M, 4) * 4}
        f = {}
        for nm, dt in _ARRAYS:
            ptr = getattr(v, nm)
            n = sizes[nm]
            ct = {np.uint8: C.c_uint8, np.int16: C.c_int16, np.float32: C.c_float}[dt]
            f[nm] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy().astype(dt)
        for nm in _SCALARS:
            f[nm] = np.array(getattr(v, nm))
        f["xw"] = np.array([[v.xw[s][t] for t in range(2)] for s in range(4)], dtype=np.int16)
        f["xf"] = np.array([[v.xf[s][t] for t in range(2)] for s in range(4)], dtype=np.float32)
        f["evparam"] = np.array(list(v.evparam), dtype=np.float32)
        f["compo"] = np.array(list(v.compo), dtype=np.float32)
        return cls(f, name, bg_filter)

    def save(self, path):
        d = dict(self.f)
        d["name"] = np.array(self.name)
        if self.bg_filter is not None:
            t, pi, eo, omega = self.bg_filter
            d["bgf_t"] = np.asarray(t, np.float32); d["bgf_pi"] = np.asarray(pi, np.float32)
            d["bgf_eo"] = np.asarray(eo, np.float32); d["bgf_omega"] = np.array(omega, np.float32)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        f = {k: z[k] for k in z.files if k not in ("name", "bgf_t", "bgf_pi", "bgf_eo", "bgf_omega")}
        bg = None
        if "bgf_t" in z.files:
            bg = (z["bgf_t"], z["bgf_pi"], z["bgf_eo"], float(z["bgf_omega"]))
        return cls(f, str(z["name"]) if "name" in z.files else "", bg)

    def view(self):
        if self._view is None:
            f = self.f
            v = OProfileView()
            for nm, dt in _ARRAYS:
                f[nm] = np.ascontiguousarray(f[nm], dtype=dt)
                setattr(v, nm, f[nm].ctypes.data)
            v.M = int(f["M"]); v.Kp = int(f["Kp"]); v.K = int(f["K"]); v.mode = int(f["mode"]); v.nj = float(f["nj"])
            for nm in ("tbm_b", "tec_b", "tjb_b", "base_b", "bias_b", "base_w", "ddbound_w"):
                setattr(v, nm, int(f[nm]))
            v.scale_b = float(f["scale_b"]); v.scale_w = float(f["scale_w"])
            for s in range(4):
                for t in range(2):
                    v.xw[s][t] = int(f["xw"][s][t]); v.xf[s][t] = float(f["xf"][s][t])
            for i in range(6):
                v.evparam[i] = float(f["evparam"][i])
            for i in range(20):
                v.compo[i] = float(f["compo"][i])
            self._view = v
        return self._view
