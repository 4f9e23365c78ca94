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
