#!/usr/bin/env python
"""bench.py — hmmsearch MSV-filter throughput (GCUPS) of libp7gpu.so on B200, with its roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's own SSE code on the host cores

Workload (BASELINE.json configs[1]): ONE synthetic calibrated profile, M=400, against 10,000,000 iid
background-frequency protein targets (log-normal lengths, mean 350) PER GPU (weak scaling: the database is sharded
by sequence, every rank filters its own shard, survivors are gathered to rank 0 with NCCL).  A "step" is one pass of
the MSV filter (p7_MSVFilter = SSV + the J-state recursion for eslENORESULT targets, plus the P-value test of
p7_pipeline.c:691-693) over the rank's whole shard.  One cell = one (target residue, model node) pair, the
reference's own Mc/s unit (p7_pipeline.c:1791-1792).

`value`  : device-resident GCUPS (database already in HBM), CUDA events, max over ranks.
`e2e`    : the same pass through the host-buffer C-ABI call p7gpu_pipeline_run_host(stop_after_msv=1): the packed
           database is copied host->device (streamed under the kernels) and the survivor records device->host inside the
           timed region.
The inputs (3.5 GB of residues per rank) are larger than L2 (126 MB), so no L2 flush is needed between steps.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LN2 = 0.69314718055994529


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nseq", type=int, default=10_000_000, help="targets per GPU")
    ap.add_argument("--mean-len", type=int, default=350)
    ap.add_argument("--profile", default="synth_M400")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the full-pipeline (config 3) and per-stage lines")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# synthetic database, generated on the device from a seeded stream (numpy cannot make 3.5 G residues in reasonable time)
# ---------------------------------------------------------------------------------------------------------------
def synth_lengths_np(n, mean, seed):
    from biobench2_b200 import seqdb
    return seqdb.synth_lengths(n, mean, np.random.default_rng(seed))


def synth_db_device(torch, dev, nseq, mean, seed, align=16):
    """Packed layout of include/p7gpu.h §4 built in HBM: (residues u8 [nbytes+32], offsets i64, lengths i32).
    align=16 (default): every target starts on a 16-byte boundary, the gap filled with sentinels (the layout the header
    recommends: the kernels read residues as aligned 16-byte vectors, so an aligned target has no masked leading rows);
    align=1: exactly one sentinel between targets."""
    from biobench2_b200 import seqdb
    lengths = synth_lengths_np(nseq, mean, seed)
    l64 = lengths.astype(np.int64)
    span = (l64 + 1 + align - 1) // align * align                       # residues + at least one sentinel, rounded up
    first = align if align > 1 else 1
    offsets = np.concatenate([[0], np.cumsum(span)[:-1]]) + first
    nbytes = int(offsets[-1] + span[-1]) if nseq else first
    g = torch.Generator(device=dev); g.manual_seed(seed)
    cdf = torch.tensor(np.cumsum(seqdb.AMINO_FREQ / seqdb.AMINO_FREQ.sum()), dtype=torch.float32, device=dev)
    res = torch.empty(nbytes + 32, dtype=torch.uint8, device=dev)
    chunk = 1 << 28
    for lo in range(0, nbytes, chunk):
        hi = min(nbytes, lo + chunk)
        u = torch.rand(hi - lo, generator=g, device=dev)
        res[lo:hi] = torch.searchsorted(cdf, u, right=True).clamp_(max=19).to(torch.uint8)
        del u
    res[nbytes:] = 255
    res[:first] = 255
    d_off = torch.from_numpy(offsets).to(dev)
    d_len = torch.from_numpy(lengths).to(dev)
    end = d_off + d_len                                                  # first byte after each target
    nxt = torch.from_numpy(offsets + span).to(dev)                       # start of the next target
    for k in range(align if align > 1 else 1):                          # 1..align sentinels after every target
        idx = end + k
        res[idx[idx < nxt]] = 255
    return res, d_off, d_len, lengths, offsets, nbytes


# ---------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (NVML)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref = unmodified HMMER 3.1b2 SSE build; else the C port)
# ---------------------------------------------------------------------------------------------------------------
class CpuMsv:
    """p7_MSVFilter over a packed sample with every host core, each thread owning an om clone (hmmsearch.c:505-516)."""

    def __init__(self, profile_name, M):
        from oracle import ref
        self.cores = os.cpu_count() or 1
        try:
            self.cores = len(os.sched_getaffinity(0))
        except Exception:
            pass
        self.M = M
        if ref.available():
            self.kind = "reference"
            self.p = ref.RefProfile.sample(M, 42, True) if profile_name.startswith("synth_M") else None
            if self.p is None:
                raise RuntimeError("reference arm supports the synthetic profiles only")
        else:
            from oracle import port
            from biobench2_b200.oprofile import OProfile
            self.kind = "port"
            op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % profile_name))
            self.p = port.OracleProfile(op.view(), op)
            self.cores = 1

    def run(self, hdb):
        """returns (seconds, cells)"""
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.p.stage_batch("msv", hdb.residues, hdb.offsets, hdb.lengths, threads=self.cores)
        else:
            self.p.stage_batch("msv", hdb)
        dt = time.perf_counter() - t0
        return dt, float(self.M) * float(hdb.nres)


def write_fasta(hdb, path):
    """Packed database -> FASTA (60 columns), vectorised."""
    from biobench2_b200 import seqdb
    sym = np.frombuffer(seqdb.AMINO_SYMS.encode(), np.uint8)
    with open(path, "wb") as fh:
        for s in range(hdb.nseq):
            o, L = int(hdb.offsets[s]), int(hdb.lengths[s])
            fh.write(b">seq%d\n" % s)
            txt = sym[hdb.residues[o:o + L]].tobytes()
            fh.write(b"\n".join(txt[i:i + 60] for i in range(0, L, 60)) + b"\n")


def hmmsearch_cli_baseline(cpu, mean, nseq=200_000):
    """The reference's own binary end to end (FASTA parse + whole pipeline), north_star's `hmmsearch --cpu N`:
    oracle/_ref/hmmsearch on a FASTA file of a sample, with --cpu 1 and --cpu <all cores>.  Reports its own Mc/s line."""
    import re, subprocess, tempfile
    from oracle import ref
    if cpu.kind != "reference" or not os.path.exists(ref.HMMSEARCH):
        return None
    tmp = tempfile.mkdtemp(prefix="p7bench_")
    hmm = os.path.join(tmp, "q.hmm"); fa = os.path.join(tmp, "db.fa")
    cpu.p.write_hmm(hmm)
    hdb = cpu_sample(nseq, mean, seed=99)
    write_fasta(hdb, fa)
    out = {"sample": "%d iid targets, %d residues, FASTA file; whole pipeline incl. parsing" % (nseq, hdb.nres)}
    for label, n in (("cpu1", 1), ("cpu%d" % cpu.cores, cpu.cores)):
        t0 = time.perf_counter()
        r = subprocess.run([ref.HMMSEARCH, "--cpu", str(n), "--tblout", os.path.join(tmp, "t.tbl"), hmm, fa],
                           capture_output=True, text=True)
        dt = time.perf_counter() - t0
        m = re.search(r"# Mc/sec:\s+([0-9.]+)", r.stdout)
        out[label] = {"elapsed_s": dt, "gcups": cpu.M * hdb.nres / dt / 1e9, "rc": r.returncode,
                      "own_mc_per_sec_line": float(m.group(1)) if m else None}
    # the input side of the same comparison: the reference's own digital FASTA reader (esl_sqio, the one
    # reader thread of hmmsearch) against this repo's parallel digitiser into the packed on-disk database, same file
    try:
        from biobench2_b200 import engine
        t0 = time.perf_counter()
        st, rdb, err = ref.read_fasta(fa, cap_bytes=hdb.residues.shape[0] + 64, cap_seqs=nseq + 1, want_names=False)
        dtr = time.perf_counter() - t0
        t0 = time.perf_counter()
        n2, r2 = engine.DbFile.from_fasta(fa, os.path.join(tmp, "db.p7db"), 0)
        dtd = time.perf_counter() - t0
        if st == 0 and n2 == nseq and r2 == hdb.nres and int(rdb["lengths"].astype(np.int64).sum()) == hdb.nres:
            out["fasta_reader"] = {"reference_esl_sqio_Mres_per_s": hdb.nres / dtr / 1e6, "p7gpu_dbfile_from_fasta_Mres_per_s": hdb.nres / dtd / 1e6,
                                   "threads": cpu.cores}
    except Exception as e:                                                # noqa: BLE001 — a baseline extra, never fatal
        out["fasta_reader"] = {"error": str(e)}
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    return out


def cpu_stage_rates(cpu, mean, nseq=24_000):
    """p7_ViterbiFilter and p7_ForwardParser of the unmodified reference on every host core (one om clone per thread), on a
    small sample: the CPU side of the per-stage kernel rates in `pipeline.stage_kernel_rates`."""
    if cpu.kind != "reference":
        return None
    hdb = cpu_sample(nseq, mean, seed=777)
    out = {"sample": "%d iid targets (%d residues), %d threads" % (nseq, hdb.nres, cpu.cores)}
    for stage in ("vit", "fwd"):
        cpu.p.stage_batch(stage, hdb.residues[:1 << 16], hdb.offsets[:64], hdb.lengths[:64], threads=1)     # warm
        t0 = time.perf_counter()
        cpu.p.stage_batch(stage, hdb.residues, hdb.offsets, hdb.lengths, threads=cpu.cores)
        out[stage + "_gcups"] = cpu.M * hdb.nres / (time.perf_counter() - t0) / 1e9
    return out


def cpu_sample(nseq, mean, seed=4242):
    from biobench2_b200 import seqdb
    return seqdb.synth_db(nseq, mean=mean, seed=seed)


def sized_cpu_sample(cpu, mean, target_seconds, nseq=400_000):
    """A fixed sample (nseq targets) and the number of passes over it that take about target_seconds on this host."""
    hdb = cpu_sample(nseq, mean)
    cpu.run(hdb)                                   # warm: page in the library, spin up threads
    dt, cells = cpu.run(hdb)
    reps = int(max(1, min(1000, round(target_seconds / max(dt, 1e-6)))))
    return hdb, nseq, reps


def cpu_timed(cpu, hdb, reps):
    tot_t = tot_c = 0.0
    for _ in range(reps):
        dt, c = cpu.run(hdb)
        tot_t += dt; tot_c += c
    return tot_t, tot_c


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    M = int(args.profile.split("M")[-1])
    cpu = CpuMsv(args.profile, M)
    per_step = max(2.0, min(15.0, 100.0 / max(1, args.steps + args.warmup)))
    hdb, n, reps = sized_cpu_sample(cpu, args.mean_len, per_step)
    for _ in range(args.warmup):
        cpu_timed(cpu, hdb, reps)
    tot_t = tot_c = 0.0
    for _ in range(args.steps):
        dt, cells = cpu_timed(cpu, hdb, reps)
        tot_t += dt; tot_c += cells
    gcups = tot_c / tot_t / 1e9
    sample = "%d passes per step over %d iid targets (mean L=%d, %d residues), p7_MSVFilter, %d threads" % (
        reps, n, args.mean_len, hdb.nres, cpu.cores)
    line = {"impl": "reference", "metric": "hmmsearch MSV-filter GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "MSV/SSV filter only: synthetic profile M=%s vs %d iid protein targets per GPU (mean L=%d), "
                        "BASELINE.json configs[1]" % (args.profile.split("M")[-1], args.nseq, args.mean_len),
            "profile": args.profile, "targets_per_gpu": args.nseq, "mean_length": args.mean_len,
            "sharding": "by sequence, %d rank(s)" % world, "l2": "inputs_larger_than_L2 (no flush needed)"}


# ---------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from biobench2_b200 import engine
    from biobench2_b200.oprofile import OProfile
    from biobench2_b200.seqdb import PackedDB

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libp7gpu has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    op = OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_%s.npz" % args.profile))
    M = op.M
    ctx = engine.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    prof = engine.Profile(ctx, op)

    with torch.cuda.stream(stream):
        res, d_off, d_len, lengths, offsets, nbytes = synth_db_device(torch, dev, args.nseq, args.mean_len, 42 + rank)
    stream.synchronize()
    db = engine.SeqDB.adopt_torch(ctx, res, d_off, d_len, int(lengths.max()))
    nres = int(lengths.astype(np.int64).sum())
    cells = float(M) * float(nres)
    d_sc = torch.empty(args.nseq, dtype=torch.float32, device=dev)
    d_st = torch.empty(args.nseq, dtype=torch.int32, device=dev)

    # survivors of the P-value test (p7_pipeline.c:691-693) are compacted by the filter kernels themselves
    F1 = 0.02
    d_pass = torch.empty(args.nseq, dtype=torch.int32, device=dev)
    d_pass_count = torch.zeros(1, dtype=torch.int32, device=dev)

    def gather_survivors():
        """The multi-GPU exchange: every rank's survivor records (global index, score) go to rank 0 with one NCCL gather
        (counts first, then the padded payload), like the MPI workers' hit return (hmmsearch.c:1242-1243)."""
        n = int(d_pass_count.item())
        if world == 1:
            return n
        idx = d_pass[:n].long()
        rec = torch.stack([(idx + rank * args.nseq).double(), d_sc[idx].double()], 1)
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        mx = int(max(int(c.item()) for c in cnts))
        pad = torch.zeros((mx, 2), dtype=torch.float64, device=dev); pad[: rec.shape[0]] = rec
        out = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, out, dst=0)
        return int(sum(int(c.item()) for c in cnts))

    def step():
        engine.msvfilter_device(ctx, prof, db, d_sc.data_ptr(), d_st.data_ptr(), F1, d_pass.data_ptr(), d_pass_count.data_ptr())
        return gather_survivors()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            nsurv = step()
        barrier()
        clocks = ClockSampler(physical_gpu_index(local_rank)); clocks.start()
        ctx.set_timing(True); ctx.reset_timing()
        l0 = ctx.launches
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            nsurv = step()
        e1.record(stream)
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launches - l0
        ssv_ms, ssv_n = ctx.kernel_time(engine.KERNEL_SSV)
        msvj_ms, msvj_n = ctx.kernel_time(engine.KERNEL_MSVJ)
        ctx.set_timing(False)
        clk = clocks.finish()
    total_cells = sum_over_ranks(cells)
    value = total_cells * args.steps / (ms_total * 1e-3) / 1e9
    statuses = torch.bincount(d_st.clamp(min=0), minlength=20).cpu().numpy()

    # ---- roofline of the dominant kernel (ssv_kernel), measured live with CUDA events on the launching stream ----
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ssv_avg_ms = ssv_ms / max(1, ssv_n)
    alg_bytes = nres + args.nseq * (8 + 4 + 4 + 4 + 4)      # dsq bytes + offset, length, order in; score, status out
    hbm_achieved = alg_bytes / (ssv_avg_ms * 1e-3) / 1e9
    int_peak, fp32_peak, lds_peak = ctx.measure_peaks()
    ssv_gcups = cells / (ssv_avg_ms * 1e-3) / 1e9
    # DRAM bytes per launch from the committed ncu --set full capture of this very launch (same profile and database size)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ssv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("profile") == args.profile and tj.get("nseq") == args.nseq and tj.get("mean_len") == args.mean_len:
            traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
    roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/ssv_traffic.json)",
                "algorithmic_bytes": alg_bytes, "kernel": "ssv_kernel", "avg_launch_ms": ssv_avg_ms,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "note": "the path moves 1 byte per M cells, so HBM is NOT the binding roof; see roofline_alu (integer issue)"}
    roofline_alu = {"bound": "int16x2-alu", "kernel": "ssv_kernel", "achieved": ssv_gcups * 2.0, "peak": int_peak,
                    "unit": "G int16 ops/s", "frac": ssv_gcups * 2.0 / int_peak if int_peak else None,
                    "ops_per_cell": 2, "kernel_gcups": ssv_gcups, "lds_peak_gbs": lds_peak,
                    "lds_achieved_gbs": ssv_gcups * 2.0, "fp32_nofma_peak_gflops": fp32_peak,
                    "peak_source": "measured live (p7gpu_measure_peaks: dependent-free VIADDMNMX.S16x2 chains)"}

    # ---- the whole filter pipeline on the same resident database (BASELINE.json configs[2]) and per-stage kernel rates ----
    pipeline = None
    if not args.no_stages and world == 1:
        KINDS = (("ssv", engine.KERNEL_SSV), ("msvj", engine.KERNEL_MSVJ), ("bias", engine.KERNEL_BIAS), ("vit", engine.KERNEL_VIT),
                 ("fwd", engine.KERNEL_FWD), ("bwd", engine.KERNEL_BWD), ("region_trigger", engine.KERNEL_TRIG))
        STAGE_OF = {"ssv": "msv", "msvj": None, "bias": "bias", "vit": "vit", "fwd": "fwd", "bwd": "bwd", "region_trigger": None}

        def timed_pipeline(dbx, prof=prof, M=M, **kw):
            engine.pipeline_run(ctx, prof, dbx, **kw)                       # warm (allocator pool, per-L tables)
            ctx.set_timing(True); ctx.reset_timing()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = engine.pipeline_run(ctx, prof, dbx, **kw)
            torch.cuda.synchronize(); wall = time.perf_counter() - t0
            km = {k: ctx.kernel_time(kind)[0] for k, kind in KINDS}
            ctx.set_timing(False)
            rates = {}
            for k, _ in KINDS:
                st_name = STAGE_OF[k]
                if st_name and km[k] > 0 and r.stage_residues[st_name] > 0:
                    rates[st_name + "_gcups"] = M * r.stage_residues[st_name] / (km[k] * 1e-3) / 1e9 if st_name != "bias" else None
                    if st_name == "bias":
                        rates["bias_gres_per_s"] = r.stage_residues["bias"] / (km[k] * 1e-3) / 1e9
            return r, wall, km, {k: v for k, v in rates.items() if v is not None}

        r, wall, km, rates = timed_pipeline(db, want_specials=True)
        pipeline = {"workload": "full pipeline MSV->bias->Viterbi->Forward->Backward, default thresholds (F1=0.02,F2=1e-3,F3=1e-5), "
                                "same profile and resident database (BASELINE.json configs[2])",
                    "ms": wall * 1e3, "gcups_whole_db": cells / wall / 1e9, "counters": [int(x) for x in r.counters],
                    "survivors": len(r.survivors), "kernel_ms": km, "stage_residues": r.stage_residues, "stage_kernel_rates": rates}
        # Forward/Backward see almost nothing of a random database at default thresholds: rate them with --max on a slice
        nsub = min(args.nseq, 100_000)
        sub_lengths = lengths[:nsub]
        sub_bytes = int(offsets[nsub - 1] + sub_lengths[-1] + 1) if nsub else 1
        sub_res = torch.empty(sub_bytes + 32, dtype=torch.uint8, device=dev); sub_res[:sub_bytes] = res[:sub_bytes]; sub_res[sub_bytes:] = 255
        sub_db = engine.SeqDB.adopt_torch(ctx, sub_res, d_off[:nsub].contiguous(), d_len[:nsub].contiguous(), int(sub_lengths.max()))
        r2, wall2, km2, rates2 = timed_pipeline(sub_db, do_max=True, want_specials=False)
        pipeline["max_mode_slice"] = {"workload": "--max (all filters off: Forward+Backward on every target) on the first %d targets" % nsub,
                                      "ms": wall2 * 1e3, "kernel_ms": km2, "stage_kernel_rates": rates2,
                                      "counters": [int(x) for x in r2.counters]}
        vsc = engine.vitfilter_batch  # Viterbi on every target of the slice (kernel time only)
        ctx.set_timing(True); ctx.reset_timing()
        vsc(ctx, prof, sub_db)
        vms = ctx.kernel_time(engine.KERNEL_VIT)[0]; ctx.set_timing(False)
        pipeline["max_mode_slice"]["stage_kernel_rates"]["vit_gcups_all_targets"] = M * float(sub_lengths.astype(np.int64).sum()) / (vms * 1e-3) / 1e9
        # BASELINE.json configs[3]: long-model stress, M=2000, --max (Forward+Backward on every target), with the region
        # trigger of the domain-definition step on the device so that specials come back only where a domain is possible
        p2000 = os.path.join(ROOT, "tests", "golden", "prof_synth_M2000.npz")
        if os.path.exists(p2000):
            op4 = OProfile.load(p2000); prof4 = engine.Profile(ctx, op4)
            r4, wall4, km4, rates4 = timed_pipeline(sub_db, prof=prof4, M=op4.M, do_max=True, want_specials=True, region_trigger=True)
            nflag = int((r4.survivors["region_flag"] == 1).sum())
            pipeline["config4_slice"] = {
                "workload": "long-model stress (BASELINE.json configs[3]): synth_M2000, --max on the first %d targets, region trigger "
                            "on the device, specials returned for flagged targets only" % nsub,
                "ms": wall4 * 1e3, "gcups_fwd_plus_bwd_wall": 2.0 * op4.M * r4.stage_residues["fwd"] / wall4 / 1e9,
                "kernel_ms": km4, "stage_kernel_rates": rates4, "counters": [int(x) for x in r4.counters],
                "region_flagged": nflag, "specials_floats_returned": int(r4.fwd_xmx.shape[0] * 2)}
            prof4.close()
        sub_db.close()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if args.e2e_steps > 0:
        h_res = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_res.copy_(res[:nbytes]); torch.cuda.synchronize()
        h_off = torch.empty(args.nseq, dtype=torch.int64, pin_memory=True); h_off.copy_(torch.from_numpy(offsets))
        h_len = torch.empty(args.nseq, dtype=torch.int32, pin_memory=True); h_len.copy_(torch.from_numpy(lengths))
        hdb = PackedDB(h_res.numpy(), h_off.numpy(), h_len.numpy())
        # the database in its 5-bit transfer format (built once per database, outside the timed region, like the packed byte
        # form itself): 5/8 of the bytes cross PCIe, which decides the rate where several GPUs share one uplink
        h_p5 = torch.empty(int(engine.lib().p7gpu_pack5_bytes(nbytes)), dtype=torch.uint8, pin_memory=True)
        engine.pack5(hdb.residues, out=h_p5.numpy())

        e2e_out = {}

        def e2e_step():
            r = engine.pipeline_run_host5(ctx, prof, h_p5.numpy(), hdb, F1=F1, stop_after_msv=True)
            e2e_out["survivors"] = len(r.msv_pass); e2e_out["counters"] = [int(x) for x in r.counters]
            r.close()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_cells * args.e2e_steps / dt / 1e9, "unit": "GCUPS", "steps": args.e2e_steps,
               "h2d_bytes_per_step": int(h_p5.numel() + args.nseq * 12), "d2h_bytes_per_step": int(e2e_out["survivors"] * 24 + 4),
               "ms_per_step": dt / args.e2e_steps * 1e3, "survivors": e2e_out["survivors"], "counters": e2e_out["counters"],
               "api": "p7gpu_pipeline_run_host5(stop_after_msv): pinned host packed DB in the 5-bit transfer format (8 residues in 5 bytes, expanded on the device) streamed in 16-256 MB chunks (database order; the last chunk longest-first) under the "
                      "SSV/MSV kernels; the MSV survivor records (index, L, status, score) come back to the host"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = CpuMsv(args.profile, M)
        sample_db, n, reps = sized_cpu_sample(cpu, args.mean_len, args.cpu_seconds)
        dt, c = cpu_timed(cpu, sample_db, reps)
        cpu_baseline = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": cpu.cores, "kind": cpu.kind,
                        "sample": "%d passes over %d iid targets (mean L=%d, %d residues), p7_MSVFilter on %d threads, %.1f s"
                                  % (reps, n, args.mean_len, sample_db.nres, cpu.cores, dt)}
        try:                                                          # the other stages of the path on the same host cores
            cpu_baseline["stages"] = cpu_stage_rates(cpu, args.mean_len)
        except Exception as e:                                       # informative only
            cpu_baseline["stages"] = {"error": repr(e)}
        try:
            cpu_baseline["hmmsearch_cli"] = hmmsearch_cli_baseline(cpu, args.mean_len)
        except Exception as e:                                       # the CLI leg is informative only
            cpu_baseline["hmmsearch_cli"] = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "hmmsearch MSV-filter GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "roofline_alu": roofline_alu,
                "cpu_baseline": cpu_baseline, "pipeline": pipeline,
                "detail": {"cells_per_step_per_gpu": cells, "residues_per_gpu": nres, "M": M,
                           "ssv_launches": int(ssv_n), "ssv_ms_total": ssv_ms, "msvj_launches": int(msvj_n), "msvj_ms_total": msvj_ms,
                           "msv_survivors_all_ranks": nsurv,
                           "status_counts": {"ok": int(statuses[0]), "erange": int(statuses[16]), "noresult": int(statuses[19])}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
