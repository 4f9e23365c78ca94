"""Host-side multi-rank logic on CPU: sharding by residue count, the survivor gather and the counter reduction over a
world-size-2 (and 3) gloo group — the N>1 path of biobench2_b200/search.py without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from biobench2_b200 import search, seqdb


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _fake_records(rank, n):
    rec = np.zeros(n, search.SURVIVOR_DTYPE)
    rec["seqidx"] = np.arange(n) * 7 + rank * 1000
    rec["L"] = 100 + rank
    rec["fwdsc"] = np.arange(n, dtype=np.float32) * 0.5 + rank
    rec["xmx_offset"] = -1
    return rec


def _worker(rank, world, port, counts, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = _fake_records(rank, counts[rank])
    allrec = search.gather_records(rec, dist)
    ctr = search.reduce_counters(np.array([10 + rank, 1000 * (rank + 1), 5, 4, 3, counts[rank]], np.uint64), dist)
    # MSV-only records use a different dtype through the same path
    m = np.zeros(counts[rank] + 1, search.MSV_DTYPE); m["seqidx"] = rank; m["usc"] = 1.5
    allm = search.gather_records(m, dist)
    if rank == 0:
        np.save(os.path.join(outdir, "rec.npy"), allrec); np.save(os.path.join(outdir, "ctr.npy"), ctr)
        np.save(os.path.join(outdir, "msv.npy"), allm)
    else:
        assert allrec is None and allm is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,counts", [(2, [5, 0]), (2, [3, 11]), (3, [0, 0, 2])])
def test_gather_and_reduce_over_gloo(tmp_path, world, counts):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, counts, str(tmp_path)), nprocs=world, join=True)
    rec = np.load(tmp_path / "rec.npy"); ctr = np.load(tmp_path / "ctr.npy"); msv = np.load(tmp_path / "msv.npy")
    want = np.concatenate([_fake_records(r, counts[r]) for r in range(world)])
    assert rec.dtype == search.SURVIVOR_DTYPE and np.array_equal(rec, want)          # rank order, nothing padded in
    assert list(ctr) == [sum(10 + r for r in range(world)), sum(1000 * (r + 1) for r in range(world)), 5 * world, 4 * world,
                         3 * world, sum(counts)]
    assert msv.shape[0] == sum(counts) + world and list(msv["seqidx"]) == sum(([r] * (counts[r] + 1) for r in range(world)), [])


def test_single_process_paths_need_no_group():
    rec = _fake_records(0, 4)
    assert search.gather_records(rec) is rec
    assert list(search.reduce_counters([1, 2, 3, 4, 5, 6])) == [1, 2, 3, 4, 5, 6]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_cover_the_database_and_balance_residues(world):
    db = seqdb.synth_db(2000, mean=200, seed=3)
    ranges = search.shard_ranges(db.lengths, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == db.nseq
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    for r in range(world):
        shard, first = search.shard_db(db, r, world)
        lo, hi = ranges[r]
        assert first == lo and shard.nseq == hi - lo
        assert abs(shard.nres - db.nres / world) <= 0.05 * db.nres + 5000
        for j in (0, shard.nseq - 1):
            if shard.nseq:
                assert np.array_equal(shard.dsq(j), db.dsq(lo + j))          # still sentinel-framed ESL_DSQs


def test_more_ranks_than_targets():
    db = seqdb.synth_db(3, mean=50, seed=1)
    got = []
    for r in range(8):
        shard, first = search.shard_db(db, r, 8)
        got.extend(range(first, first + shard.nseq))
    assert got == [0, 1, 2]


def _worker2(rank, world, port, counts, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = (np.arange(counts[rank], dtype=np.float32) + 100 * rank)
    allf, starts = search.gather_floats(f, dist)
    # records that are "still on the device": fixed capacity, a count tensor, no host read inside the exchange
    cap = 6
    rec = torch.zeros((cap, 2), dtype=torch.float64); n = min(cap, counts[rank])
    rec[:n, 0] = torch.arange(n, dtype=torch.float64) + 10 * rank; rec[:n, 1] = rank
    out, cnts = search.gather_device_records(rec, torch.tensor([n], dtype=torch.int64), dist)
    if rank == 0:
        np.save(os.path.join(outdir, "f.npy"), allf); np.save(os.path.join(outdir, "starts.npy"), starts)
        np.save(os.path.join(outdir, "dev.npy"), np.stack([o.numpy() for o in out])); np.save(os.path.join(outdir, "cnts.npy"), np.array([int(c.item()) for c in cnts]))
    else:
        assert allf is None and starts is None and out is None
        assert [int(c.item()) for c in cnts] == [min(cap, c) for c in counts]      # the counts are all-gathered: every rank has them
    dist.destroy_process_group()


@pytest.mark.parametrize("world,counts", [(2, [4, 9]), (2, [0, 3]), (3, [5, 0, 1])])
def test_float_and_device_record_gathers_over_gloo(tmp_path, world, counts):
    port = _free_port()
    mp.spawn(_worker2, args=(world, port, counts, str(tmp_path)), nprocs=world, join=True)
    f = np.load(tmp_path / "f.npy"); starts = np.load(tmp_path / "starts.npy")
    want = np.concatenate([np.arange(counts[r], dtype=np.float32) + 100 * r for r in range(world)])
    assert np.array_equal(f, want) and list(starts) == list(np.concatenate([[0], np.cumsum(counts)[:-1]]))
    dev = np.load(tmp_path / "dev.npy"); cnts = np.load(tmp_path / "cnts.npy")
    for r in range(world):
        n = min(6, counts[r])
        assert cnts[r] == n and np.array_equal(dev[r, :n, 0], np.arange(n) + 10 * r) and (dev[r, :n, 1] == r).all() and (dev[r, n:] == 0).all()


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
def test_c_shard_plan_balances_residues(nshards):
    """p7gpu_shard_plan (the C host's sharding, include/p7gpu.h section 9): contiguous, covering, balanced by residue count"""
    from biobench2_b200 import engine
    db = seqdb.synth_db(5000, mean=250, seed=9)
    packed = seqdb.pack([db.residues[db.offsets[s]: db.offsets[s] + db.lengths[s]] for s in range(db.nseq)], align=16)
    first = engine.shard_plan(packed, nshards)
    assert first[0] == 0 and first[-1] == packed.nseq and all(first[k] <= first[k + 1] for k in range(nshards))
    for k in range(nshards):
        res = int(packed.lengths[first[k]: first[k + 1]].astype(np.int64).sum())
        assert abs(res - packed.nres / nshards) <= 0.03 * packed.nres + 4000
