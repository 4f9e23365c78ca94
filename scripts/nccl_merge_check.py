#!/usr/bin/env python
"""Multi-rank search ending in the reference's own hit list (hmmsearch.c:1008-1029, p7_tophits.c:326,401,866).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_merge_check.py

Every rank searches its shard of one seeded database (planted homologs) on its own GPU; survivor records, counters and the
survivors' specials go to rank 0 over NCCL (biobench2_b200/search.py); rank 0 pushes them through the unmodified reference's
post-processing (p7_domaindef_ByPosteriorHeuristics, hit creation, p7_tophits_SortBySortkey, p7_tophits_Threshold: oracle/_ref
harness) and compares counters and the hit list with (a) the same search on one rank and (b) the reference's own p7_Pipeline
over the whole database.  Prints one JSON line; exit status 0 = identical."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from biobench2_b200 import engine, search, seqdb
    from biobench2_b200.oprofile import OProfile
    from oracle import ref
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dmod = dist if world > 1 else None
    model = sys.argv[1] if len(sys.argv) > 1 else "Pkinase"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    rp = ref.RefProfile.open(os.path.join(ROOT, "tests", "golden", model + ".hmm"))
    base = seqdb.synth_db(n, mean=350, seed=77, degenerate_frac=0.002)
    seqs = [base.residues[base.offsets[s]: base.offsets[s] + base.lengths[s]] for s in range(n)]
    rng = np.random.default_rng(5)
    for k, s in enumerate(rng.choice(n, size=240, replace=False)):
        h = rp.emit(3000 + k, Lcfg=350)
        if h is not None and len(h):
            seqs[int(s)] = h
    hdb = seqdb.pack(seqs, align=16)
    ctx = engine.Context(local)
    prof = engine.Profile(ctx, OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % model)))
    counters, rec, fx, bx = search.sharded_search_with_specials(ctx, prof, hdb, dmod, fb_exact=True)
    ok = True; out = {}
    if rank == 0:
        def hits_of(c, r, f, b):
            hs = rp.finish_from_specials(hdb.residues, hdb.offsets, hdb.lengths, r["seqidx"], r["fwdsc"], f, b, r["xmx_offset"],
                                         r["fwd_totscale"], r["bck_totscale"], int(c[0]), int(c[1]))
            return [(h.seqidx, np.float32(h.score).tobytes(), h.ndom, h.flags, np.float64(h.lnP).tobytes()) for h in hs]
        merged = hits_of(counters, rec, fx, bx)
        one = engine.pipeline_run_host(ctx, prof, hdb, want_specials=True, fb_exact=True)             # the same search, one rank
        r1 = search.survivors_to_records(one)
        single = hits_of(one.counters, r1, one.fwd_xmx, one.bck_xmx)
        rc, rhits = rp.pipeline(hdb.residues, hdb.offsets, hdb.lengths)                              # the reference itself
        refl = [(h.seqidx, np.float32(h.score).tobytes(), h.ndom, h.flags, np.float64(h.lnP).tobytes()) for h in rhits]
        ok = (list(counters) == list(one.counters) == list(rc)) and merged == single == refl and len(merged) > 100
        out = {"world": world, "targets": n, "counters": [int(x) for x in counters], "reference_counters": [int(x) for x in rc],
               "hits_merged": len(merged), "hits_one_rank": len(single), "hits_reference": len(refl),
               "survivor_records_gathered": int(rec.shape[0]), "specials_floats_gathered": int(fx.shape[0] * 2), "identical": bool(ok)}
        print(json.dumps(out), flush=True)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local)); dist.broadcast(flag, 0); ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
