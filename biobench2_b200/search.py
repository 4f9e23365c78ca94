"""Sharded search: the host-side flow of hmmsearch's worker/merge loop for one rank per GPU.

What the reference does with threads (HMMER/src/hmmsearch.c:505-546: per-worker pipeline + tophits, then
p7_tophits_Merge / p7_pipeline_Merge) or MPI (hmmsearch.c:806-1283: byte-range blocks out, packed hits + counters
back, merged by the master at :1008-1029) is done here with torch.distributed: the packed database is cut into
contiguous ranges balanced by residue count, each rank runs the batched pipeline (libp7gpu.so) on its range, and rank 0
receives every rank's survivor records and the summed counters.  There is no collective on the data path; the exchange
is one all_gather of counts, one gather of fixed-size records and one reduce(sum) — NCCL on GPUs, gloo in CPU tests.
"""
import numpy as np

SURVIVOR_DTYPE = np.dtype([
    ("seqidx", "<i8"), ("L", "<i4"), ("msv_status", "<i4"), ("vit_status", "<i4"),
    ("usc", "<f4"), ("vfsc", "<f4"), ("fwdsc", "<f4"), ("bcksc", "<f4"), ("nullsc", "<f4"), ("filtersc", "<f4"),
    ("fwd_totscale", "<f4"), ("bck_totscale", "<f4"), ("bck_own_scales", "<i4"), ("xmx_offset", "<i8"),
    ("region_flag", "<i4"), ("mocc_max", "<f4"), ("dec_offset", "<i8")])                             # = p7gpu_survivor, 80 bytes
MSV_DTYPE = np.dtype([("seqidx", "<i8"), ("L", "<i4"), ("msv_status", "<i4"), ("usc", "<f4"), ("stage", "<i4")])
assert SURVIVOR_DTYPE.itemsize == 80 and MSV_DTYPE.itemsize == 24


def shard_ranges(lengths, world):
    """Contiguous target ranges [lo, hi) per rank, balanced by residue count (+1 sentinel per target) — SURVEY.md §8e."""
    l = np.asarray(lengths, np.int64) + 1
    cum = np.concatenate([[0], np.cumsum(l)])
    total = int(cum[-1])
    cuts = [int(np.searchsorted(cum, total * r // world, side="left")) for r in range(world)] + [len(l)]
    cuts = [min(c, len(l)) for c in cuts]
    return [(cuts[r], max(cuts[r], cuts[r + 1])) for r in range(world)]


def shard_db(packed, rank, world):
    """This rank's slice of a host PackedDB as its own PackedDB (views, no residue copy) and its first global index."""
    from .seqdb import PackedDB
    lo, hi = shard_ranges(packed.lengths, world)[rank]
    if hi <= lo:
        return PackedDB(np.full(1, 255, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32)), lo
    b_lo = int(packed.offsets[lo]) - 1
    b_hi = int(packed.offsets[hi - 1]) + int(packed.lengths[hi - 1]) + 1
    return PackedDB(packed.residues[b_lo:b_hi], packed.offsets[lo:hi] - b_lo, packed.lengths[lo:hi]), lo


def survivors_to_records(result, first_index=0):
    """engine.PipelineResult -> structured array with GLOBAL target indices."""
    rec = np.zeros(len(result.survivors), SURVIVOR_DTYPE)
    for k in SURVIVOR_DTYPE.names:
        if k != "_pad":
            rec[k] = result.survivors[k]
    rec["seqidx"] += first_index
    return rec


def _dev(dist_mod, device):
    if device is not None:
        return device
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist_mod.get_backend() == "nccl" else torch.device("cpu")


def gather_records(rec, dist_mod=None, device=None, dst=0):
    """Gather fixed-size records (any structured dtype) from every rank to `dst`, in rank order.

    One all_gather of per-rank counts, then one gather of the payload padded to the largest count (NCCL's gather needs
    equal sizes).  Returns the concatenated array on `dst`, None elsewhere.  Single-process: returns rec.
    """
    import torch
    if dist_mod is None or not dist_mod.is_initialized() or dist_mod.get_world_size() == 1:
        return rec
    world, rank = dist_mod.get_world_size(), dist_mod.get_rank()
    device = _dev(dist_mod, device)
    cnt = torch.tensor([rec.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist_mod.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    mx = max(max(counts), 1)
    isz = rec.dtype.itemsize
    pad = np.zeros(mx * isz, np.uint8)
    pad[: rec.shape[0] * isz] = np.ascontiguousarray(rec).view(np.uint8).reshape(-1)
    t = torch.from_numpy(pad).to(device)
    out = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
    dist_mod.gather(t, out, dst=dst)
    if rank != dst:
        return None
    parts = [out[r].cpu().numpy()[: counts[r] * isz].view(rec.dtype) for r in range(world)]
    return np.concatenate(parts) if parts else rec[:0]


def gather_device_records(d_rec, d_count, dist_mod=None, dst=0):
    """The same exchange for records that are still in HBM, with NO host synchronisation: `d_rec` is a fixed-capacity device
    tensor [cap, ...] of which the first d_count[0] rows are valid (d_count: int64[1] on the same device).  One all_gather of the
    counts and one gather of the fixed-size buffers are queued on the current stream; nothing is read back.  Returns
    (list of per-rank buffers on `dst` | None, per-rank count tensors).  The caller reads the counts once, after its last step
    (a search exchanges survivors once per query: hmmsearch.c:1242-1243; a per-step .item() cost 1.3 % of an MSV pass)."""
    import torch
    if dist_mod is None or not dist_mod.is_initialized() or dist_mod.get_world_size() == 1:
        return [d_rec], [d_count]
    world, rank = dist_mod.get_world_size(), dist_mod.get_rank()
    cnts = [torch.empty_like(d_count) for _ in range(world)]
    dist_mod.all_gather(cnts, d_count)
    out = [torch.empty_like(d_rec) for _ in range(world)] if rank == dst else None
    dist_mod.gather(d_rec, out, dst=dst)
    return out, cnts


def gather_floats(arr, dist_mod=None, device=None, dst=0):
    """Gather variable-length float32 arrays (e.g. the survivors' Forward/Backward specials) to `dst`: one all_gather of the
    lengths, one gather of the payload padded to the longest.  Returns (concatenated array, per-rank start offsets) on `dst`,
    (None, None) elsewhere."""
    import torch
    arr = np.ascontiguousarray(arr, np.float32)
    if dist_mod is None or not dist_mod.is_initialized() or dist_mod.get_world_size() == 1:
        return arr, np.zeros(1, np.int64)
    world, rank = dist_mod.get_world_size(), dist_mod.get_rank()
    device = _dev(dist_mod, device)
    cnt = torch.tensor([arr.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist_mod.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    mx = max(max(counts), 1)
    pad = np.zeros(mx, np.float32); pad[: arr.shape[0]] = arr
    t = torch.from_numpy(pad).to(device)
    out = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
    dist_mod.gather(t, out, dst=dst)
    if rank != dst:
        return None, None
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    return np.concatenate([out[r].cpu().numpy()[: counts[r]] for r in range(world)]), starts


def sharded_search_with_specials(ctx, prof, packed, dist_mod=None, **opts):
    """sharded_pipeline plus the survivors' specials on rank 0: what the master needs to build the hit list from every rank's
    survivors (the MPI master's role, hmmsearch.c:1008-1029).  Returns (counters, records, fwd_xmx, bck_xmx) on rank 0 with
    records['xmx_offset'] rebased into the concatenated arrays; four Nones elsewhere."""
    from . import engine
    world = dist_mod.get_world_size() if dist_mod is not None and dist_mod.is_initialized() else 1
    rank = dist_mod.get_rank() if world > 1 else 0
    shard, first = shard_db(packed, rank, world)
    res = engine.pipeline_run_host(ctx, prof, shard, want_specials=True, **opts)
    rec = survivors_to_records(res, first)
    nrec = np.array([rec.shape[0]], np.int64)
    fx, starts = gather_floats(res.fwd_xmx, dist_mod)
    bx, _ = gather_floats(res.bck_xmx, dist_mod)
    allrec = gather_records(rec, dist_mod)
    counts = gather_records(nrec.view([("n", "<i8")]), dist_mod)
    counters = reduce_counters(res.counters, dist_mod)
    if rank != 0:
        return None, None, None, None
    if world > 1:                                                   # each rank's offsets point into its own arrays: rebase
        lo = 0
        for r in range(world):
            n = int(counts["n"][r]); seg = allrec[lo: lo + n]
            seg["xmx_offset"] = np.where(seg["xmx_offset"] >= 0, seg["xmx_offset"] + starts[r], -1)
            lo += n
    return counters, allrec, fx, bx


def reduce_counters(counters, dist_mod=None, device=None, dst=0):
    """Sum the six pipeline counters over ranks (p7_pipeline_Merge, p7_pipeline.c:597-636).  Result valid on `dst`."""
    import torch
    c = np.asarray(counters, np.int64)
    if dist_mod is None or not dist_mod.is_initialized() or dist_mod.get_world_size() == 1:
        return c.astype(np.uint64)
    t = torch.from_numpy(c.copy()).to(_dev(dist_mod, device))
    dist_mod.reduce(t, dst=dst, op=dist_mod.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def sharded_pipeline(ctx, prof, packed, dist_mod=None, want_specials=False, **opts):
    """Run the batched pipeline on this rank's shard of `packed` and merge on rank 0.

    Returns (counters, survivor records with global indices) on rank 0 and (None, None) elsewhere.  Forward/Backward
    specials stay on the rank that produced them unless want_specials (then they are returned locally as well).
    """
    from . import engine
    world = dist_mod.get_world_size() if dist_mod is not None and dist_mod.is_initialized() else 1
    rank = dist_mod.get_rank() if world > 1 else 0
    shard, first = shard_db(packed, rank, world)
    res = engine.pipeline_run_host(ctx, prof, shard, want_specials=want_specials, **opts)
    rec = survivors_to_records(res, first) if not opts.get("stop_after_msv") else res.msv_pass.copy()
    if opts.get("stop_after_msv"):
        rec["seqidx"] += first
    allrec = gather_records(rec, dist_mod)
    counters = reduce_counters(res.counters, dist_mod)
    if rank != 0:
        return None, None
    return counters, allrec
