"""ctypes binding of libp7gpu.so (include/p7gpu.h): the only way this package computes anything.

There is deliberately no CPU path: if the shared library is missing, or no sm_100 device is present,
every compute call raises.  (The CPU oracle lives in oracle/ and is test infrastructure.)
"""
import ctypes as C
import os
import numpy as np

from .oprofile import OProfileView

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libp7gpu.so")

OK, EMEM, EINVAL, ESYS, ERANGE, ENORESULT = 0, 5, 11, 12, 16, 19
KERNEL_SSV, KERNEL_MSVJ, KERNEL_VIT, KERNEL_FWD, KERNEL_BWD, KERNEL_BIAS, KERNEL_TRIG, KERNEL_FWDX, KERNEL_BWDX = range(9)

EXPORTS = [
    "p7gpu_ctx_create", "p7gpu_ctx_destroy", "p7gpu_last_error", "p7gpu_device_count", "p7gpu_ctx_set_stream",
    "p7gpu_ctx_launch_count", "p7gpu_ctx_synchronize", "p7gpu_ctx_set_option", "p7gpu_ctx_last_vit_redo", "p7gpu_ctx_bias_recomputed",
    "p7gpu_ctx_set_timing", "p7gpu_ctx_reset_timing", "p7gpu_ctx_kernel_time", "p7gpu_measure_peaks",
    "p7gpu_profile_upload", "p7gpu_profile_destroy", "p7gpu_profile_M", "p7gpu_profile_set_biasfilter", "p7gpu_profile_prepare_lengths",
    "p7gpu_seqdb_upload", "p7gpu_seqdb_adopt_device", "p7gpu_seqdb_destroy", "p7gpu_seqdb_nseq", "p7gpu_seqdb_nres",
    "p7gpu_msvfilter_batch", "p7gpu_ssvfilter_batch", "p7gpu_vitfilter_batch", "p7gpu_forward_batch",
    "p7gpu_backward_batch", "p7gpu_region_trigger_batch", "p7gpu_msvfilter_device", "p7gpu_biasfilter_batch", "p7gpu_msvfilter_stream",
    "p7gpu_SSVFilter", "p7gpu_MSVFilter", "p7gpu_ViterbiFilter", "p7gpu_ForwardParser", "p7gpu_BackwardParser", "p7gpu_BackwardParser2",
    "p7gpu_pipeline_run", "p7gpu_pipeline_run_many", "p7gpu_scan_msv_prefilter", "p7gpu_scan_table_build", "p7gpu_pipeline_run_host", "p7gpu_pipeline_run_host5", "p7gpu_pack5", "p7gpu_pack5_bytes", "p7gpu_stage_threshold", "p7gpu_result_nmsvpass", "p7gpu_result_msvpass", "p7gpu_result_stage_residues", "p7gpu_result_counters", "p7gpu_result_nsurvivors", "p7gpu_result_survivors",
    "p7gpu_result_fwd_xmx", "p7gpu_result_bck_xmx", "p7gpu_result_decoding", "p7gpu_result_destroy",
    "p7gpu_nullsc", "p7gpu_msv_pvalue",
    "p7gpu_dbfile_from_fasta", "p7gpu_dbfile_open", "p7gpu_dbfile_close", "p7gpu_dbfile_nseq", "p7gpu_dbfile_nres",
    "p7gpu_dbfile_nbytes", "p7gpu_dbfile_maxlength", "p7gpu_dbfile_residues", "p7gpu_dbfile_offsets", "p7gpu_dbfile_lengths",
    "p7gpu_dbfile_name", "p7gpu_dbfile_desc", "p7gpu_dbfile_pin", "p7gpu_pipeline_run_file",
    "p7gpu_shard_plan", "p7gpu_pipeline_run_sharded", "p7gpu_pipeline_run_file_sharded", "p7gpu_counters_merge",
]


class P7GpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libp7gpu status %d: %s" % (code, msg))
        self.code = code


class PipelineOpts(C.Structure):
    _fields_ = [("F1", C.c_double), ("F2", C.c_double), ("F3", C.c_double),
                ("do_biasfilter", C.c_int32), ("do_max", C.c_int32), ("want_specials", C.c_int32),
                ("stop_after_msv", C.c_int32), ("region_trigger", C.c_int32), ("fb_exact", C.c_int32),
                ("decode", C.c_int32), ("stage_records", C.c_int32)]


class MsvSurvivor(C.Structure):
    _fields_ = [("seqidx", C.c_int64), ("L", C.c_int32), ("msv_status", C.c_int32), ("usc", C.c_float), ("stage", C.c_int32)]


class PipelineCounters(C.Structure):
    _fields_ = [("nseqs", C.c_uint64), ("nres", C.c_uint64), ("n_past_msv", C.c_uint64), ("n_past_bias", C.c_uint64),
                ("n_past_vit", C.c_uint64), ("n_past_fwd", C.c_uint64)]


class Survivor(C.Structure):
    _fields_ = [("seqidx", C.c_int64), ("L", C.c_int32), ("msv_status", C.c_int32), ("vit_status", C.c_int32),
                ("usc", C.c_float), ("vfsc", C.c_float), ("fwdsc", C.c_float), ("bcksc", C.c_float),
                ("nullsc", C.c_float), ("filtersc", C.c_float), ("fwd_totscale", C.c_float), ("bck_totscale", C.c_float),
                ("bck_own_scales", C.c_int32), ("xmx_offset", C.c_int64), ("region_flag", C.c_int32), ("mocc_max", C.c_float),
                ("dec_offset", C.c_int64)]


SURVIVOR_NP = np.dtype([("seqidx", "<i8"), ("L", "<i4"), ("msv_status", "<i4"), ("vit_status", "<i4"),
                        ("usc", "<f4"), ("vfsc", "<f4"), ("fwdsc", "<f4"), ("bcksc", "<f4"), ("nullsc", "<f4"),
                        ("filtersc", "<f4"), ("fwd_totscale", "<f4"), ("bck_totscale", "<f4"), ("bck_own_scales", "<i4"),
                        ("xmx_offset", "<i8"), ("region_flag", "<i4"), ("mocc_max", "<f4"), ("dec_offset", "<i8")])
assert SURVIVOR_NP.itemsize == C.sizeof(Survivor) == 80


_lib = None


def lib():
    """Load libp7gpu.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libp7gpu.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C biobench2_b200/csrc`; this package has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.p7gpu_last_error.restype = C.c_char_p
        L.p7gpu_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.p7gpu_ctx_destroy.argtypes = [vp]
        L.p7gpu_ctx_set_stream.argtypes = [vp, vp]
        L.p7gpu_ctx_launch_count.restype = i64
        L.p7gpu_ctx_launch_count.argtypes = [vp]
        L.p7gpu_ctx_synchronize.argtypes = [vp]
        L.p7gpu_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_int]
        L.p7gpu_ctx_last_vit_redo.restype = i64
        L.p7gpu_ctx_last_vit_redo.argtypes = [vp]
        L.p7gpu_ctx_bias_recomputed.restype = i64
        L.p7gpu_ctx_bias_recomputed.argtypes = [vp]
        L.p7gpu_ctx_set_timing.argtypes = [vp, C.c_int]
        L.p7gpu_ctx_reset_timing.argtypes = [vp]
        L.p7gpu_ctx_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64)]
        L.p7gpu_measure_peaks.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.p7gpu_profile_upload.argtypes = [vp, C.POINTER(OProfileView), C.POINTER(vp)]
        L.p7gpu_profile_destroy.argtypes = [vp]
        L.p7gpu_profile_M.argtypes = [vp]
        L.p7gpu_seqdb_upload.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(vp)]
        L.p7gpu_seqdb_adopt_device.argtypes = [vp, vp, i64, vp, vp, i64, i32, C.POINTER(vp)]
        L.p7gpu_seqdb_destroy.argtypes = [vp]
        L.p7gpu_seqdb_nseq.restype = i64
        L.p7gpu_seqdb_nseq.argtypes = [vp]
        L.p7gpu_seqdb_nres.restype = i64
        L.p7gpu_seqdb_nres.argtypes = [vp]
        for nm in ("p7gpu_msvfilter_batch", "p7gpu_ssvfilter_batch", "p7gpu_vitfilter_batch"):
            getattr(L, nm).argtypes = [vp, vp, vp, vp, i64, vp, vp]
        L.p7gpu_forward_batch.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
        L.p7gpu_backward_batch.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        L.p7gpu_msvfilter_device.argtypes = [vp, vp, vp, C.c_double, vp, vp, vp, vp]
        L.p7gpu_profile_set_biasfilter.argtypes = [vp, vp, vp, vp]
        L.p7gpu_profile_prepare_lengths.argtypes = [vp, vp, C.c_int64, C.c_double]
        L.p7gpu_biasfilter_batch.argtypes = [vp, vp, vp, vp, i64, vp]
        L.p7gpu_region_trigger_batch.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        L.p7gpu_msvfilter_stream.argtypes = [vp, vp, vp, i64, vp, vp, i64, C.c_double, vp, vp, vp, C.POINTER(i64), C.POINTER(vp)]
        L.p7gpu_pipeline_run_host.argtypes = [vp, vp, vp, i64, vp, vp, i64, C.POINTER(PipelineOpts), C.POINTER(vp)]
        L.p7gpu_result_stage_residues.restype = C.POINTER(C.c_uint64)
        L.p7gpu_result_stage_residues.argtypes = [vp]
        L.p7gpu_result_nmsvpass.restype = i64
        L.p7gpu_result_nmsvpass.argtypes = [vp]
        L.p7gpu_result_msvpass.restype = C.POINTER(MsvSurvivor)
        L.p7gpu_result_msvpass.argtypes = [vp]
        for nm in ("p7gpu_SSVFilter", "p7gpu_MSVFilter", "p7gpu_ViterbiFilter"):
            getattr(L, nm).argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_float)]
        L.p7gpu_ForwardParser.argtypes = [vp, vp, C.c_int, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.p7gpu_BackwardParser.argtypes = [vp, vp, C.c_int, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.p7gpu_BackwardParser2.argtypes = [vp, vp, C.c_int, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.p7gpu_pipeline_run.argtypes = [vp, vp, vp, C.POINTER(PipelineOpts), C.POINTER(vp)]
        L.p7gpu_shard_plan.argtypes = [vp, vp, i64, C.c_int, vp]
        L.p7gpu_pipeline_run_sharded.argtypes = [vp, vp, C.c_int, vp, vp, i64, vp, vp, i64, C.POINTER(PipelineOpts), vp, vp]
        L.p7gpu_pipeline_run_file_sharded.argtypes = [vp, vp, C.c_int, vp, C.POINTER(PipelineOpts), vp, vp]
        L.p7gpu_counters_merge.argtypes = [C.POINTER(PipelineCounters), C.POINTER(PipelineCounters)]
        L.p7gpu_counters_merge.restype = None
        L.p7gpu_dbfile_from_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]
        L.p7gpu_dbfile_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.p7gpu_dbfile_close.argtypes = [vp]
        for nm in ("nseq", "nres", "nbytes"):
            getattr(L, "p7gpu_dbfile_" + nm).argtypes = [vp]; getattr(L, "p7gpu_dbfile_" + nm).restype = i64
        L.p7gpu_dbfile_maxlength.argtypes = [vp]; L.p7gpu_dbfile_maxlength.restype = C.c_int32
        for nm in ("residues", "offsets", "lengths"):
            getattr(L, "p7gpu_dbfile_" + nm).argtypes = [vp]; getattr(L, "p7gpu_dbfile_" + nm).restype = vp
        for nm in ("name", "desc"):
            getattr(L, "p7gpu_dbfile_" + nm).argtypes = [vp, i64]; getattr(L, "p7gpu_dbfile_" + nm).restype = C.c_char_p
        L.p7gpu_dbfile_pin.argtypes = [vp]
        L.p7gpu_pipeline_run_file.argtypes = [vp, vp, vp, C.POINTER(PipelineOpts), C.POINTER(vp)]
        L.p7gpu_stage_threshold.argtypes = [C.c_int, C.c_double, C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.p7gpu_pack5_bytes.argtypes = [i64]; L.p7gpu_pack5_bytes.restype = i64
        L.p7gpu_pack5.argtypes = [vp, i64, vp, C.c_int]
        L.p7gpu_pipeline_run_host5.argtypes = [vp, vp, vp, i64, vp, vp, i64, C.POINTER(PipelineOpts), C.POINTER(vp)]
        L.p7gpu_pipeline_run_many.argtypes = [vp, vp, i64, vp, C.POINTER(PipelineOpts), vp]
        L.p7gpu_scan_msv_prefilter.argtypes = [vp, vp, i64, vp, vp, C.c_double, vp]
        L.p7gpu_scan_table_build.argtypes = [vp, vp, vp, C.c_int32, C.c_double, vp]
        L.p7gpu_result_counters.restype = C.POINTER(PipelineCounters)
        L.p7gpu_result_counters.argtypes = [vp]
        L.p7gpu_result_nsurvivors.restype = i64
        L.p7gpu_result_nsurvivors.argtypes = [vp]
        L.p7gpu_result_survivors.restype = C.POINTER(Survivor)
        L.p7gpu_result_survivors.argtypes = [vp]
        L.p7gpu_result_fwd_xmx.restype = C.POINTER(C.c_float)
        L.p7gpu_result_fwd_xmx.argtypes = [vp]
        L.p7gpu_result_bck_xmx.restype = C.POINTER(C.c_float)
        L.p7gpu_result_bck_xmx.argtypes = [vp]
        L.p7gpu_result_decoding.restype = C.POINTER(C.c_float)
        L.p7gpu_result_decoding.argtypes = [vp]
        L.p7gpu_result_destroy.argtypes = [vp]
        L.p7gpu_nullsc.restype = C.c_float
        L.p7gpu_nullsc.argtypes = [C.c_int]
        L.p7gpu_msv_pvalue.restype = C.c_double
        L.p7gpu_msv_pvalue.argtypes = [C.c_float] * 4
        _lib = L
    return _lib


def _check(st):
    if st != OK:
        raise P7GpuError(st, lib().p7gpu_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().p7gpu_device_count()


class Context:
    """One per worker / GPU (cf. WORKER_INFO, HMMER/src/hmmsearch.c:36-44)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().p7gpu_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib().p7gpu_ctx_destroy(self.h)
            self.h = None

    def set_stream(self, cuda_stream_ptr):
        _check(lib().p7gpu_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _check(lib().p7gpu_ctx_synchronize(self.h))

    @property
    def launches(self):
        return lib().p7gpu_ctx_launch_count(self.h)

    def set_option(self, name, value):
        _check(lib().p7gpu_ctx_set_option(self.h, name.encode(), int(value)))

    @property
    def last_vit_redo(self):
        """targets of the last Viterbi launch that left the packed kernel's window and were redone exactly"""
        return lib().p7gpu_ctx_last_vit_redo(self.h)

    @property
    def bias_recomputed(self):
        """bias-filter scores recomputed on the host so far (targets the bias kernel could not certify, bias.cu)"""
        return lib().p7gpu_ctx_bias_recomputed(self.h)

    def set_timing(self, on=True):
        _check(lib().p7gpu_ctx_set_timing(self.h, int(on)))

    def reset_timing(self):
        _check(lib().p7gpu_ctx_reset_timing(self.h))

    def kernel_time(self, kind):
        """(total ms, launches) of one kernel kind since reset_timing(); kinds: KERNEL_SSV .. KERNEL_BWD."""
        ms = C.c_double(0); n = C.c_int64(0)
        _check(lib().p7gpu_ctx_kernel_time(self.h, kind, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def measure_peaks(self):
        """Measured issue peaks: (packed int16x2 add-max Gop/s, non-fused fp32 GFLOP/s, LDS.32 GB/s)."""
        a = C.c_double(0); b = C.c_double(0); c = C.c_double(0)
        _check(lib().p7gpu_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


class Profile:
    """GPU-resident repacked profile (p7gpu_profile)."""

    def __init__(self, ctx, oprofile):
        """`oprofile`: biobench2_b200.oprofile.OProfile, or anything with .view() -> p7gpu_oprofile_view."""
        self.ctx = ctx
        self._src = oprofile
        v = oprofile.view()
        if not isinstance(v, OProfileView):      # a structurally identical ctypes struct from elsewhere
            v = C.cast(C.pointer(v), C.POINTER(OProfileView)).contents
        h = C.c_void_p()
        _check(lib().p7gpu_profile_upload(ctx.h, C.byref(v), C.byref(h)))
        self.h = h
        self.M = v.M
        self.evparam = [float(x) for x in v.evparam]
        bg = getattr(oprofile, "bg_filter", None)
        if bg is not None:
            self.set_biasfilter(*bg[:3])

    def set_biasfilter(self, t, pi, eo):
        """bg->fhmm tables after p7_bg_SetFilter (HMMER/src/p7_bg.c:430-453): t[2][3], pi[3], eo[Kp][2]."""
        t = np.ascontiguousarray(t, np.float32); pi = np.ascontiguousarray(pi, np.float32); eo = np.ascontiguousarray(eo, np.float32)
        assert t.size == 6 and pi.size == 3 and eo.size == 58
        _check(lib().p7gpu_profile_set_biasfilter(self.h, _p(t), _p(pi), _p(eo)))

    def prepare_lengths(self, lengths, F1=0.02):
        """Per-length constants for exactly these target lengths (scan mode: few fixed targets, many profiles)."""
        a = np.ascontiguousarray(lengths, np.int32)
        _check(lib().p7gpu_profile_prepare_lengths(self.h, _p(a), a.shape[0], float(F1)))

    def close(self):
        if self.h:
            lib().p7gpu_profile_destroy(self.h)
            self.h = None


class SeqDB:
    """GPU-resident packed target database (p7gpu_seqdb)."""

    def __init__(self, ctx, handle, nseq, keep=None):
        self.ctx, self.h, self.nseq, self._keep = ctx, handle, nseq, keep

    @classmethod
    def upload(cls, ctx, packed):
        h = C.c_void_p()
        _check(lib().p7gpu_seqdb_upload(ctx.h, _p(packed.residues), packed.residues.shape[0], _p(packed.offsets),
                                        _p(packed.lengths), packed.nseq, C.byref(h)))
        return cls(ctx, h, packed.nseq)

    @classmethod
    def adopt_torch(cls, ctx, residues_t, offsets_t, lengths_t, max_length):
        """Wrap torch CUDA tensors (uint8 / int64 / int32) without copying; residues need >= 32 slack bytes."""
        h = C.c_void_p()
        nbytes = int(residues_t.numel()) - 32
        _check(lib().p7gpu_seqdb_adopt_device(ctx.h, C.c_void_p(residues_t.data_ptr()), nbytes,
                                              C.c_void_p(offsets_t.data_ptr()), C.c_void_p(lengths_t.data_ptr()),
                                              int(lengths_t.numel()), int(max_length), C.byref(h)))
        return cls(ctx, h, int(lengths_t.numel()), keep=(residues_t, offsets_t, lengths_t))

    @property
    def nres(self):
        return lib().p7gpu_seqdb_nres(self.h)

    def close(self):
        if self.h:
            lib().p7gpu_seqdb_destroy(self.h)
            self.h = None


def _stage(fn, ctx, prof, db, idx):
    n = len(idx) if idx is not None else db.nseq
    idx_a = np.ascontiguousarray(idx, np.int64) if idx is not None else None
    sc = np.zeros(n, np.float32); st = np.zeros(n, np.int32)
    _check(fn(ctx.h, prof.h, db.h, _p(idx_a), n, _p(sc), _p(st)))
    return sc, st


def ssvfilter_batch(ctx, prof, db, idx=None): return _stage(lib().p7gpu_ssvfilter_batch, ctx, prof, db, idx)
def msvfilter_batch(ctx, prof, db, idx=None): return _stage(lib().p7gpu_msvfilter_batch, ctx, prof, db, idx)
def vitfilter_batch(ctx, prof, db, idx=None): return _stage(lib().p7gpu_vitfilter_batch, ctx, prof, db, idx)


def msvfilter_batch_into(ctx, prof, db, sc, st):
    """p7gpu_msvfilter_batch over the whole database into caller-owned (e.g. pinned) host arrays."""
    assert sc.dtype == np.float32 and st.dtype == np.int32 and sc.shape[0] >= db.nseq and st.shape[0] >= db.nseq
    _check(lib().p7gpu_msvfilter_batch(ctx.h, prof.h, db.h, None, db.nseq, _p(sc), _p(st)))


def xmx_offsets_for(lengths):
    l = np.asarray(lengths, np.int64)
    return np.concatenate([[0], np.cumsum((l + 1) * 6)[:-1]]).astype(np.int64), int(((l + 1) * 6).sum())


def forward_batch(ctx, prof, db, lengths, idx=None, want_xmx=True):
    """lengths: host array of the selected targets' lengths (sizes the specials)."""
    n = len(idx) if idx is not None else db.nseq
    idx_a = np.ascontiguousarray(idx, np.int64) if idx is not None else None
    sc = np.zeros(n, np.float32); st = np.zeros(n, np.int32); tot = np.zeros(n, np.float32)
    xoff, ext = xmx_offsets_for(lengths)
    xmx = np.zeros(max(ext, 1), np.float32) if want_xmx else None
    _check(lib().p7gpu_forward_batch(ctx.h, prof.h, db.h, _p(idx_a), n, _p(sc), _p(st), _p(xmx),
                                     _p(xoff) if want_xmx else None, _p(tot)))
    return sc, st, xmx, xoff, tot


def backward_batch(ctx, prof, db, fwd_xmx, xoff, idx=None):
    n = len(idx) if idx is not None else db.nseq
    idx_a = np.ascontiguousarray(idx, np.int64) if idx is not None else None
    sc = np.zeros(n, np.float32); st = np.zeros(n, np.int32); tot = np.zeros(n, np.float32)
    bx = np.zeros_like(fwd_xmx)
    _check(lib().p7gpu_backward_batch(ctx.h, prof.h, db.h, _p(idx_a), n, _p(fwd_xmx), _p(xoff), _p(sc), _p(st), _p(bx), _p(tot)))
    return sc, st, bx, tot


def region_trigger_batch(ctx, prof, db, fwd_xmx, bck_xmx, xoff, idx=None, own_scales=None):
    """max_i mocc[i] and the region-trigger flag of p7_domaindef_ByPosteriorHeuristics per target (p7gpu_region_trigger_batch)."""
    n = len(idx) if idx is not None else db.nseq
    idx_a = np.ascontiguousarray(idx, np.int64) if idx is not None else None
    own = np.ascontiguousarray(own_scales, np.int32) if own_scales is not None else None
    mocc = np.zeros(n, np.float32); flag = np.zeros(n, np.int32)
    _check(lib().p7gpu_region_trigger_batch(ctx.h, prof.h, db.h, _p(idx_a), n, _p(fwd_xmx), _p(bck_xmx),
                                            _p(np.ascontiguousarray(xoff, np.int64)), _p(own), _p(mocc), _p(flag)))
    return mocc, flag


def msvfilter_device(ctx, prof, db, d_sc_ptr, d_status_ptr, F1=0.02, d_pass_list_ptr=None, d_pass_count_ptr=None):
    """MSV stage with results left in HBM (device pointers: nseq floats / int32; optional pass list int32[nseq] + uint32 count)."""
    _check(lib().p7gpu_msvfilter_device(ctx.h, prof.h, db.h, F1, C.c_void_p(d_sc_ptr), C.c_void_p(d_status_ptr),
                                        C.c_void_p(d_pass_list_ptr) if d_pass_list_ptr else None,
                                        C.c_void_p(d_pass_count_ptr) if d_pass_count_ptr else None))


def biasfilter_batch(ctx, prof, db, idx=None):
    n = len(idx) if idx is not None else db.nseq
    idx_a = np.ascontiguousarray(idx, np.int64) if idx is not None else None
    out = np.zeros(n, np.float32)
    _check(lib().p7gpu_biasfilter_batch(ctx.h, prof.h, db.h, _p(idx_a), n, _p(out)))
    return out


def msvfilter_stream(ctx, prof, packed, F1=0.02, sc=None, st=None, want_pass=True, keep_db=False):
    """Streamed MSV pass over a HOST PackedDB (p7gpu_msvfilter_stream).  Returns (sc, status, pass_list, SeqDB|None)."""
    n = packed.nseq
    sc = np.zeros(n, np.float32) if sc is None else sc
    st = np.zeros(n, np.int32) if st is None else st
    pl = np.zeros(max(n, 1), np.int32) if want_pass else None
    npass = C.c_int64(0); h = C.c_void_p()
    _check(lib().p7gpu_msvfilter_stream(ctx.h, prof.h, _p(packed.residues), packed.residues.shape[0], _p(packed.offsets),
                                        _p(packed.lengths), n, F1, _p(sc), _p(st), _p(pl), C.byref(npass),
                                        C.byref(h) if keep_db else None))
    db = SeqDB(ctx, h, n) if keep_db and h else None
    return sc, st, (pl[:npass.value] if want_pass else None), db


class _ResultHandle:
    """Owns one p7gpu_pipeline_result; the arrays viewing its buffers keep it alive."""

    def __init__(self, h):
        self.h = h

    def destroy(self):
        if self.h:
            lib().p7gpu_result_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class _Owned(np.ndarray):
    """ndarray view of library memory that keeps its owner alive (also through derived views, via .base)."""
    _owner = None


def _view(ptr, nbytes, dtype, owner):
    raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
    a = raw.view(dtype).view(_Owned)
    a._owner = owner
    return a


def decoding_block(result, sv):
    """(btot, etot, mocc, fwd_scale, fwd_xC_L) of one survivor record of an opts.decode run"""
    L = int(sv["L"]); o = int(sv["dec_offset"]); d = result.decoding
    return d[o: o + L + 1], d[o + L + 1: o + 2 * (L + 1)], d[o + 2 * (L + 1): o + 3 * (L + 1)], d[o + 3 * (L + 1): o + 4 * (L + 1)], d[o + 4 * (L + 1)]


class PipelineResult:
    """A p7gpu_pipeline_result.  The record arrays are views of the library's own buffers (no copy).  They stay valid while
    they (or this object) are referenced, or until close(); all of it must end before the Context is closed."""

    def __init__(self, h):
        L = lib()
        self._own = own = _ResultHandle(h)
        c = L.p7gpu_result_counters(h).contents
        self.counters = np.array([c.nseqs, c.nres, c.n_past_msv, c.n_past_bias, c.n_past_vit, c.n_past_fwd], np.uint64)
        sr = L.p7gpu_result_stage_residues(h)
        self.stage_residues = {k: int(sr[i]) for i, k in enumerate(("msv", "bias", "vit", "fwd", "bwd", "fwdx"))}
        n = L.p7gpu_result_nsurvivors(h)      # one structured array (a --max run returns millions of records)
        self.survivors = _view(L.p7gpu_result_survivors(h), n * C.sizeof(Survivor), SURVIVOR_NP, own) if n else np.zeros(0, SURVIVOR_NP)
        have = self.survivors["xmx_offset"] >= 0
        ext = int((self.survivors["xmx_offset"][have] + (self.survivors["L"][have].astype(np.int64) + 1) * 6).max()) if have.any() else 0
        if ext:
            self.fwd_xmx = _view(L.p7gpu_result_fwd_xmx(h), ext * 4, np.float32, own)
            self.bck_xmx = _view(L.p7gpu_result_bck_xmx(h), ext * 4, np.float32, own)
        else:
            self.fwd_xmx = self.bck_xmx = np.zeros(0, np.float32)
        hd = self.survivors["dec_offset"] >= 0 if n else np.zeros(0, bool)
        dext = int((self.survivors["dec_offset"][hd] + 4 * (self.survivors["L"][hd].astype(np.int64) + 1) + 2).max()) if hd.any() else 0
        self.decoding = _view(L.p7gpu_result_decoding(h), dext * 4, np.float32, own) if dext else np.zeros(0, np.float32)
        m = L.p7gpu_result_nmsvpass(h)
        self.msv_pass = _view(L.p7gpu_result_msvpass(h), m * C.sizeof(MsvSurvivor), MSV_NP, own) if m else np.zeros(0, MSV_NP)

    def close(self):
        """Release the library buffers now (the arrays of this result must not be used afterwards)."""
        self.survivors = self.fwd_xmx = self.bck_xmx = self.msv_pass = self.decoding = None
        self._own.destroy()


MSV_NP = np.dtype([("seqidx", "<i8"), ("L", "<i4"), ("msv_status", "<i4"), ("usc", "<f4"), ("stage", "<i4")])


def _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger=False, fb_exact=False, decode=False,
          stage_records=False):
    return PipelineOpts(F1, F2, F3, int(do_biasfilter), int(do_max), int(want_specials), int(stop_after_msv), int(region_trigger), int(fb_exact),
                        int(decode), int(stage_records))


def pipeline_run(ctx, prof, db, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False, want_specials=True,
                 stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """The batched p7_Pipeline() over a resident database (defaults: p7_pipeline.c:214-216)."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    h = C.c_void_p()
    _check(lib().p7gpu_pipeline_run(ctx.h, prof.h, db.h, C.byref(o), C.byref(h)))
    return PipelineResult(h)


def pipeline_run_many(ctx, profs, db, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False, want_specials=True,
                      stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """Many profiles against one resident database (p7gpu_pipeline_run_many) -> list of PipelineResult."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    n = len(profs)
    ph = (C.c_void_p * n)(*[p.h for p in profs])
    rh = (C.c_void_p * n)()
    _check(lib().p7gpu_pipeline_run_many(ctx.h, ph, n, db.h, C.byref(o), rh))
    return [PipelineResult(C.c_void_p(rh[i])) for i in range(n)]


def scan_msv_prefilter(ctx, oprofiles, db, lengths, F1=0.02):
    """hmmscan's MSV stage for a block of models in one launch (p7gpu_scan_msv_prefilter): per model, the number of targets of
    `db` that pass F1 by their SSV score or need the J recursion; 0 = pipeline_run(stop_after_msv) would return no survivor.
    `oprofiles`: objects with .view() -> p7gpu_oprofile_view (not uploaded); `lengths`: the target lengths of `db` (host)."""
    n = len(oprofiles)
    views = (OProfileView * max(n, 1))()
    for k, o in enumerate(oprofiles):
        v = o.view()
        if not isinstance(v, OProfileView):
            v = C.cast(C.pointer(v), C.POINTER(OProfileView)).contents
        C.memmove(C.byref(views, k * C.sizeof(OProfileView)), C.byref(v), C.sizeof(OProfileView))
    a = np.ascontiguousarray(lengths, np.int32)
    out = np.zeros(max(n, 1), np.int32)
    _check(lib().p7gpu_scan_msv_prefilter(ctx.h, views, n, db.h, _p(a), float(F1), _p(out)))
    return out[:n]


def scan_table(oprofile, L, F1=0.02):
    """Host-only diagnostic (p7gpu_scan_table_build): the block prefilter's table of one model -> (table[29][rowWords] uint32,
    P, npass, tjb_b, msv_xj_min for target length L).  Needs no device."""
    v = oprofile.view()
    if not isinstance(v, OProfileView):
        v = C.cast(C.pointer(v), C.POINTER(OProfileView)).contents
    geom = np.zeros(3, np.int32)
    _check(lib().p7gpu_scan_table_build(C.byref(v), _p(geom), None, int(L), float(F1), None))
    tab = np.zeros((29, int(geom[1])), np.uint32)          # P7GPU_KP rows
    w = np.zeros(1, np.uint32)
    _check(lib().p7gpu_scan_table_build(C.byref(v), _p(geom), _p(tab), int(L), float(F1), _p(w)))
    return tab, int(geom[0]), int(geom[2]), int(w[0] & 0xff), int(w[0] >> 16)


def pipeline_run_host(ctx, prof, packed, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False, want_specials=True,
                      stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """Same for a HOST PackedDB: the upload is streamed under the MSV stage."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    h = C.c_void_p()
    _check(lib().p7gpu_pipeline_run_host(ctx.h, prof.h, _p(packed.residues), packed.residues.shape[0], _p(packed.offsets),
                                         _p(packed.lengths), packed.nseq, C.byref(o), C.byref(h)))
    return PipelineResult(h)


def stage_threshold(kind, F, mu, lam):
    """Smallest float seq_score whose P-value passes F (p7gpu_stage_threshold); None when no verified threshold exists."""
    T = C.c_float(0)
    st = lib().p7gpu_stage_threshold({"gumbel": 0, "exp": 1}[kind], float(F), float(mu), float(lam), C.byref(T))
    if st == ERANGE:
        return None
    _check(st)
    return np.float32(T.value)


def pack5(residues, out=None, nthreads=0):
    """Residue bytes -> the 5-bit transfer format (p7gpu_pack5).  `out`: optional preallocated uint8 array (e.g. pinned)."""
    residues = np.ascontiguousarray(residues, np.uint8)
    nb = lib().p7gpu_pack5_bytes(residues.shape[0])
    if out is None:
        out = np.empty(nb, np.uint8)
    assert out.dtype == np.uint8 and out.shape[0] >= nb
    _check(lib().p7gpu_pack5(_p(residues), residues.shape[0], _p(out), int(nthreads)))
    return out


def pipeline_run_host5(ctx, prof, packed5, packed, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False, want_specials=True,
                       stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """pipeline_run_host with the residues in the 5-bit transfer format; `packed` supplies offsets, lengths and the unpacked size."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    h = C.c_void_p()
    _check(lib().p7gpu_pipeline_run_host5(ctx.h, prof.h, _p(packed5), packed.residues.shape[0], _p(packed.offsets), _p(packed.lengths),
                                          packed.nseq, C.byref(o), C.byref(h)))
    return PipelineResult(h)


class DbFile:
    """The packed on-disk target database (include/p7gpu.h section 8): built once from FASTA, mapped for every search."""

    @staticmethod
    def from_fasta(fasta_path, out_path, nthreads=0, align=16):
        """align=16: every target starts on a 16-byte boundary (recommended); align=1: exactly one sentinel between targets."""
        n = C.c_int64(0); r = C.c_int64(0)
        _check(lib().p7gpu_dbfile_from_fasta(str(fasta_path).encode(), str(out_path).encode(), int(nthreads), int(align), C.byref(n), C.byref(r)))
        return n.value, r.value

    def __init__(self, path):
        h = C.c_void_p()
        _check(lib().p7gpu_dbfile_open(str(path).encode(), C.byref(h)))
        self.h = h
        L = lib()
        self.nseq, self.nres, self.nbytes = L.p7gpu_dbfile_nseq(h), L.p7gpu_dbfile_nres(h), L.p7gpu_dbfile_nbytes(h)
        self.max_length = L.p7gpu_dbfile_maxlength(h)

    def _arr(self, fn, ct, n):
        return np.ctypeslib.as_array(C.cast(fn(self.h), C.POINTER(ct)), shape=(max(n, 0),)) if n else np.zeros(0, ct)

    def packed(self):
        """Views of the mapped arrays as a seqdb.PackedDB (valid until close())."""
        from .seqdb import PackedDB
        L = lib()
        return PackedDB(self._arr(L.p7gpu_dbfile_residues, C.c_uint8, self.nbytes), self._arr(L.p7gpu_dbfile_offsets, C.c_int64, self.nseq),
                        self._arr(L.p7gpu_dbfile_lengths, C.c_int32, self.nseq))

    def name(self, s): return lib().p7gpu_dbfile_name(self.h, s).decode()
    def desc(self, s): return lib().p7gpu_dbfile_desc(self.h, s).decode()
    def pin(self): _check(lib().p7gpu_dbfile_pin(self.h))

    def close(self):
        if self.h:
            lib().p7gpu_dbfile_close(self.h); self.h = None


def pipeline_run_file(ctx, prof, dbfile, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False, want_specials=True,
                      stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """The batched pipeline over a mapped packed database file (p7gpu_pipeline_run_file)."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    h = C.c_void_p()
    _check(lib().p7gpu_pipeline_run_file(ctx.h, prof.h, dbfile.h, C.byref(o), C.byref(h)))
    return PipelineResult(h)


def shard_plan(packed, nshards):
    """Contiguous shards balanced by residue count (p7gpu_shard_plan) -> int64[nshards + 1] boundaries."""
    first = np.zeros(nshards + 1, np.int64)
    _check(lib().p7gpu_shard_plan(_p(packed.offsets), _p(packed.lengths), packed.nseq, nshards, _p(first)))
    return first


def merge_counters(results):
    """Sum of the shards' P7_PIPELINE counters (p7gpu_counters_merge == p7_pipeline_Merge, p7_pipeline.c:597-636)."""
    tot = PipelineCounters()
    for r in results:
        c = PipelineCounters(*[int(x) for x in r.counters])
        lib().p7gpu_counters_merge(C.byref(tot), C.byref(c))
    return np.array([tot.nseqs, tot.nres, tot.n_past_msv, tot.n_past_bias, tot.n_past_vit, tot.n_past_fwd], np.uint64)


def pipeline_run_sharded(ctxs, profs, packed, packed5=None, dbfile=None, F1=0.02, F2=1e-3, F3=1e-5, do_biasfilter=True, do_max=False,
                         want_specials=True, stop_after_msv=False, region_trigger=False, fb_exact=False, decode=False, stage_records=False):
    """One HOST database over several GPUs of this process: one context, one contiguous shard and one host thread per GPU
    (p7gpu_pipeline_run_sharded / _file_sharded).  Returns (first[nshards+1], [PipelineResult per shard]); survivor seqidx are
    database indices.  Merge like the reference merges its workers: merge_counters() + the hit lists through p7_tophits_Merge."""
    o = _opts(F1, F2, F3, do_biasfilter, do_max, want_specials, stop_after_msv, region_trigger, fb_exact, decode, stage_records)
    n = len(ctxs)
    ch = (C.c_void_p * n)(*[c.h for c in ctxs]); ph = (C.c_void_p * n)(*[p.h for p in profs]); rh = (C.c_void_p * n)()
    first = np.zeros(n + 1, np.int64)
    if dbfile is not None:
        _check(lib().p7gpu_pipeline_run_file_sharded(ch, ph, n, dbfile.h, C.byref(o), _p(first), rh))
    else:
        _check(lib().p7gpu_pipeline_run_sharded(ch, ph, n, _p(packed.residues) if packed5 is None else None, _p(packed5),
                                                packed.residues.shape[0], _p(packed.offsets), _p(packed.lengths), packed.nseq,
                                                C.byref(o), _p(first), rh))
    return first, [PipelineResult(C.c_void_p(rh[i])) for i in range(n)]


# ---- signature-level wrappers (one target per call), same meaning as impl_sse.h:309-344 ----
def _one(fn, ctx, dsq, L, prof):
    sc = C.c_float(0)
    dsq = np.ascontiguousarray(dsq, np.uint8)
    st = fn(ctx.h, _p(dsq), L, prof.h, C.byref(sc))
    if st not in (OK, ERANGE, ENORESULT):
        _check(st)
    return st, np.float32(sc.value)


def p7_SSVFilter(ctx, dsq, L, om): return _one(lib().p7gpu_SSVFilter, ctx, dsq, L, om)
def p7_MSVFilter(ctx, dsq, L, om): return _one(lib().p7gpu_MSVFilter, ctx, dsq, L, om)
def p7_ViterbiFilter(ctx, dsq, L, om): return _one(lib().p7gpu_ViterbiFilter, ctx, dsq, L, om)


def p7_ForwardParser(ctx, dsq, L, om):
    sc = C.c_float(0); tot = C.c_float(0)
    xmx = np.zeros((L + 1, 6), np.float32)
    dsq = np.ascontiguousarray(dsq, np.uint8)
    st = lib().p7gpu_ForwardParser(ctx.h, _p(dsq), L, om.h, _p(xmx), C.byref(tot), C.byref(sc))
    if st not in (OK, ERANGE):
        _check(st)
    return st, np.float32(sc.value), xmx, np.float32(tot.value)


def p7_BackwardParser(ctx, dsq, L, om, fwd_xmx, want_own=False):
    sc = C.c_float(0); tot = C.c_float(0); own = C.c_int(0)
    bx = np.zeros((L + 1, 6), np.float32)
    dsq = np.ascontiguousarray(dsq, np.uint8)
    fwd_xmx = np.ascontiguousarray(fwd_xmx, np.float32)
    if want_own:
        st = lib().p7gpu_BackwardParser2(ctx.h, _p(dsq), L, om.h, _p(fwd_xmx), _p(bx), C.byref(tot), C.byref(sc), C.byref(own))
    else:
        st = lib().p7gpu_BackwardParser(ctx.h, _p(dsq), L, om.h, _p(fwd_xmx), _p(bx), C.byref(tot), C.byref(sc))
    if st not in (OK, ERANGE):
        _check(st)
    if want_own:
        return st, np.float32(sc.value), bx, np.float32(tot.value), own.value
    return st, np.float32(sc.value), bx, np.float32(tot.value)
