#!/usr/bin/env python
"""BASELINE.json configs[4] — hmmscan-style batch: many Pfam-shaped synthetic profiles (M = 100..800, tests/golden/scan) against
one database sharded by sequence over the GPUs of one box, each profile run through the FULL pipeline (MSV -> bias -> Viterbi
-> Forward -> Backward, default thresholds) with p7gpu_pipeline_run_many on the rank's resident shard.

    python scripts/config5_bench.py --nseq 6250000 --passes 48                       (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/config5_bench.py --gpus N --nseq 6250000 --passes 48

--nseq is the shard per GPU (50M targets / 8 GPUs = 6.25M), --passes how many profile passes are timed (the 12 committed
profiles are cycled; a pass costs the same whichever profile of that length it is).  Timing: barrier + synchronize on both
sides, max over ranks.  Prints one JSON line with profile passes/s, whole-job GCUPS and the extrapolation to 1000 profiles."""
import argparse, glob, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--nseq", type=int, default=6_250_000)
    ap.add_argument("--passes", type=int, default=24); ap.add_argument("--mean-len", type=int, default=350)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench import synth_db_device
    from biobench2_b200 import engine, search
    from biobench2_b200.oprofile import OProfile
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = engine.Context(local)
    res, d_off, d_len, lengths, offsets, nbytes = synth_db_device(torch, dev, args.nseq, args.mean_len, 42 + rank)
    torch.cuda.synchronize()
    db = engine.SeqDB.adopt_torch(ctx, res, d_off, d_len, int(lengths.max()))
    nres = int(lengths.astype(np.int64).sum())
    paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scan", "prof_scan_M*.npz")))
    ops = [OProfile.load(p) for p in paths]
    t0 = time.perf_counter()
    profs = [engine.Profile(ctx, op) for op in ops]
    upload_ms = (time.perf_counter() - t0) * 1e3 / len(profs)
    batch = [profs[i % len(profs)] for i in range(args.passes)]
    Ms = [ops[i % len(ops)].M for i in range(args.passes)]
    rs = engine.pipeline_run_many(ctx, profs, db, want_specials=False)          # warm: per-L tables, allocator pools
    for r in rs:
        r.close()
    barrier(); t0 = time.perf_counter()
    rs = engine.pipeline_run_many(ctx, batch, db, want_specials=False)
    barrier(); dt = time.perf_counter() - t0
    counters = np.sum([r.counters.astype(np.int64) for r in rs], axis=0)
    nsurv = sum(len(r.survivors) for r in rs)
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        counters = search.reduce_counters(counters.astype(np.uint64), dist).astype(np.int64)
        t = torch.tensor([float(nres), float(nsurv)], dtype=torch.float64, device=dev); dist.all_reduce(t); tot_res, nsurv = float(t[0]), int(t[1])
    else:
        tot_res = float(nres)
    cells = float(sum(Ms)) * tot_res
    if rank == 0:
        per_pass = dt / args.passes
        print(json.dumps({
            "workload": "hmmscan-style batch (BASELINE.json configs[4]): %d profile passes (12 synthetic calibrated profiles, M=100..800, cycled) "
                        "x %d targets per GPU (mean L=%d), full pipeline with default thresholds, database resident and sharded by sequence"
                        % (args.passes, args.nseq, args.mean_len),
            "n_gpus": world, "passes": args.passes, "seconds": dt, "ms_per_profile_pass": per_pass * 1e3,
            "gcups_whole_pipeline": cells / dt / 1e9, "mean_M": float(np.mean(Ms)), "targets_all_gpus": args.nseq * world,
            "profile_upload_ms_each": upload_ms,
            "counters_sum": [int(x) for x in counters], "forward_survivors": int(nsurv),
            "extrapolated_seconds_for_1000_profiles": per_pass * 1000.0}))
    for r in rs:
        r.close()
    if world > 1:
        dist.destroy_process_group()


main()
