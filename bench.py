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
`e2e`    : the same pass through the host-buffer C-ABI call p7gpu_pipeline_run_host5(stop_after_msv=1): the packed
           database is copied host->device (streamed under the kernels) and the survivor records device->host inside the
           timed region.
`roofline`          : the dominant kernel (ssv_tm_kernel / ssv_kernel) against its BINDING roof, the packed-int16 issue rate measured live
                      (HBM, which the schema asks about and which is not binding, sits in roofline.hbm).
`roofline_kernels`  : the same fraction for every other kernel of the path (Viterbi, Forward, Backward, bias, the
                      exact-order Forward/Backward): SURVEY.md 8(d) ops per cell x measured cells/s / the live-measured peak.
`configs`           : BASELINE.json configs[2..4] as first-class entries: the full pipeline at default thresholds, the long-model
                      --max stress at the FULL database size, the many-profiles batch (200 profile passes).
`e2e_app`           : whole-application wall clock, FASTA in -> hit tables out: integration/_build/hmmsearch_b200 against the
                      reference's own hmmsearch (oracle/_ref) on the same files, outputs compared.
`strong`            : the same MSV pass over ONE fixed database sharded by residue count over the ranks (strong scaling).
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
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--strong-total", type=int, default=20_000_000, help="targets of the fixed database of the strong-scaling leg (0 = skip)")
    ap.add_argument("--config5-passes", type=int, default=200)
    ap.add_argument("--no-app", action="store_true", help="skip the whole-application (hmmsearch_b200 vs hmmsearch) leg")
    ap.add_argument("--app-reps", type=int, default=20, help="whole-application leg: also search this many concatenated copies of Swiss-Prot (0/1 = only the shipped database)")
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


def synth_db_device(torch, dev, nseq, mean, seed, align=16, lengths=None):
    """Packed layout of include/p7gpu.h §4 built in HBM: (residues u8 [nbytes+32], offsets i64, lengths i32).
    align=16 (default): every target starts on a 16-byte boundary, the gap filled with sentinels (the layout the header
    recommends: the kernels read residues as aligned 16-byte vectors, so an aligned target has no masked leading rows);
    align=1: exactly one sentinel between targets."""
    from biobench2_b200 import seqdb
    if lengths is None:
        lengths = synth_lengths_np(nseq, mean, seed)
    lengths = np.ascontiguousarray(lengths, np.int32)
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
        # the same query against the same database written out `reps` times (names made unique): the start-up cost of a GPU
        # process is paid once, the work scales — the size at which the whole application is faster than the reference
        if reps > 1:
            big = os.path.join(tmp, "sprot_x%d.fasta" % reps)
            with open(fa, "rb") as f:
                body = f.read()
            with open(big, "wb") as f:
                for r in range(reps):
                    f.write(body.replace(b">", b">r%02d_" % r) if r else body)
            del body
            os.sync()                                              # the 1.2 GB just written are not flushed under the timed runs
            tag = "tufa420_x_swissprot_x%d" % reps
            one(tag, hmm, big)
            out[tag]["workload"] = "tufa420.hmm vs %d concatenated copies of uniprot_sprot.fasta (%d targets, names prefixed per copy)" % (reps, 153325 * reps)
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
            "config": workload_config(args, world, sample),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world, sample=None):
    c = {"workload": "MSV/SSV filter only: synthetic profile M=%s vs %d iid protein targets per GPU (mean L=%d), "
                     "BASELINE.json configs[1]" % (args.profile.split("M")[-1], args.nseq, args.mean_len),
         "profile": args.profile, "targets_per_gpu": args.nseq, "mean_length": args.mean_len,
         "sharding": "by sequence, %d rank(s)" % world, "l2": "inputs_larger_than_L2 (no flush needed)"}
    if sample:
        c["reference_arm_sample"] = sample
    return c


# minimum arithmetic per DP cell and the unit it issues in
OPS_PER_CELL = {"ssv": 1.0,      # packed int16x2 instr per cell (2 int16 ops)
                "vit": 5.0,      # packed int16x2 instr per cell with fused add-max (14 int16 ops)
                "fwd": 16.0,     # fp32 flop per cell
                "bwd": 18.0}


def kernel_rooflines(M, rates, int_peak, fp32_peak, lds_peak):
    """rates: {'vit_gcups', 'fwd_gcups', 'bwd_gcups', 'fwdx_gcups', 'bwdx_gcups', 'bias_gres_per_s'} -> roofline_kernels"""
    out = {}
    pk_instr = int_peak / 2.0 if int_peak else None                 # G packed int16x2 instr/s (2 int16 ops each)
    if rates.get("vit_gcups") and pk_instr:
        a = rates["vit_gcups"] * OPS_PER_CELL["vit"]
        out["vitp_kernel"] = {"bound": "int16x2-issue", "achieved": a, "peak": pk_instr, "unit": "G packed int16x2 instr/s", "frac": a / pk_instr,
                              "gcups": rates["vit_gcups"], "instr_per_cell": OPS_PER_CELL["vit"]}
    for k, name, kern in (("fwd_gcups", "fwd", "fwd_mw_kernel"), ("bwd_gcups", "bwd", "bwd_mw_kernel")):
        if rates.get(k) and fp32_peak:
            a = rates[k] * OPS_PER_CELL[name]
            out[kern] = {"bound": "fp32-issue", "achieved": a, "peak": fp32_peak, "unit": "GFLOP/s (fp32 mul/add, one per issue slot)",
                         "frac": a / fp32_peak, "frac_of_fma_peak": a / (2.0 * fp32_peak), "gcups": rates[k], "flop_per_cell": OPS_PER_CELL[name]}
    # exact-order parsers: every state cell goes through the shared-memory pipe several times per row (fbx.cu header):
    # Forward 11 LDS/STS words per cell + 3 table words, Backward 17 + 3 -> bytes per cell against the LDS.32 peak
    for k, kern, words in (("fwdx_gcups", "fbx_fwd_kernel", 14.0), ("bwdx_gcups", "fbx_bwd_kernel", 20.0)):
        if rates.get(k) and lds_peak:
            a = rates[k] * words * 4.0
            out[kern] = {"bound": "smem-lsu", "achieved": a, "peak": lds_peak, "unit": "GB/s through the LSU (shared + L1)", "frac": a / lds_peak,
                         "gcups": rates[k], "lsu_words_per_cell": words}
    if rates.get("bias_gres_per_s") and fp32_peak:
        a = rates["bias_gres_per_s"] * 14.0                          # 2-state Forward: 4 mul + 2 add per state, renormalisation
        out["bias_kernel"] = {"bound": "fp32-issue", "achieved": a, "peak": fp32_peak, "unit": "GFLOP/s", "frac": a / fp32_peak,
                              "gres_per_s": rates["bias_gres_per_s"], "flop_per_residue": 14.0,
                              "note": "one thread per MSV survivor: latency-bound by design (2.3 ms of a 121 ms search)"}
    return out


def app_e2e(ROOT, cores, reps=1):
    """Whole application, FASTA in -> hit tables out: hmmsearch_b200 against the reference's hmmsearch on the same files."""
    import re, subprocess, tempfile
    b200 = os.path.join(ROOT, "integration", "_build", "hmmsearch_b200")
    refbin = os.path.join(ROOT, "oracle", "_ref", "hmmsearch")
    hmm = os.path.join(ROOT, "tests", "_refdata", "tufa420.hmm"); fa = os.path.join(ROOT, "tests", "_refdata", "uniprot_sprot.fasta")
    if not (os.path.exists(b200) and os.path.exists(refbin)):
        return {"unavailable": "integration/_build/hmmsearch_b200 or oracle/_ref/hmmsearch not built"}
    out = {}
    tmp = tempfile.mkdtemp(prefix="p7app_")

    def payload(path):
        with open(path, errors="replace") as f:
            return [l for l in f.read().split("\n") if not l.startswith("#")]

    def one(tag, hmmf, faf, extra=()):
        rec = {}
        for name, exe, more in (("reference", refbin, ["--cpu", str(cores)]), ("b200", b200, ["--cpu", str(cores), "--times"])):
            best = None; walls = []
            for rep in range(3):                                   # later runs: page cache warm for both
                # a process that used the GPU is still being torn down by the driver for a moment after it exited (hmmsearch_b200
                # leaves with _exit); the next run should not start inside that — same pause before either binary
                time.sleep(0.5)
                t0 = time.perf_counter()
                p = subprocess.run([exe, "-o", os.path.join(tmp, name + ".out"), "--tblout", os.path.join(tmp, name + ".tbl"),
                                    "--domtblout", os.path.join(tmp, name + ".dom")] + list(extra) + more + [hmmf, faf], capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if p.returncode != 0:
                    rec[name] = {"rc": p.returncode, "stderr": p.stderr[-400:]}
                    break
                best = dt if best is None else min(best, dt)
                walls.append(round(dt, 3))
            if best is not None:
                rec[name] = {"wall_s": best, "threads": cores}
                rec[name]["wall_s_all_runs"] = walls
                if name == "b200":
                    rec[name]["phases"] = [l[2:] for l in p.stderr.split("\n") if l.startswith("# ")]
        if "wall_s" in rec.get("reference", {}) and "wall_s" in rec.get("b200", {}):
            same = all(payload(os.path.join(tmp, "reference" + e)) == payload(os.path.join(tmp, "b200" + e)) for e in (".out", ".tbl", ".dom"))
            rec["outputs_identical"] = bool(same)
            rec["speedup"] = rec["reference"]["wall_s"] / rec["b200"]["wall_s"]
            m = re.findall(r"^Passed (MSV|bias|Vit|Fwd) filter:\s+(\d+)", open(os.path.join(tmp, "b200.out")).read(), re.M)
            rec["counters_past_msv_bias_vit_fwd"] = [int(x[1]) for x in m]
        out[tag] = rec
    if os.path.exists(hmm) and os.path.exists(fa):
        one("config1_tufa420_x_swissprot", hmm, fa)
        out["config1_tufa420_x_swissprot"]["workload"] = "BASELINE.json configs[0], runBench.pl:115: tufa420.hmm (M=317) vs uniprot_sprot.fasta, 153,325 targets"
        # the same query against the same database written out `reps` times (names made unique): the start-up cost of a GPU
        # process is paid once, the work scales — the size at which the whole application is faster than the reference
        if reps > 1:
            big = os.path.join(tmp, "sprot_x%d.fasta" % reps)
            with open(fa, "rb") as f:
                body = f.read()
            with open(big, "wb") as f:
                for r in range(reps):
                    f.write(body.replace(b">", b">r%02d_" % r) if r else body)
            del body
            os.sync()                                              # the 1.2 GB just written are not flushed under the timed runs
            tag = "tufa420_x_swissprot_x%d" % reps
            one(tag, hmm, big)
            out[tag]["workload"] = "tufa420.hmm vs %d concatenated copies of uniprot_sprot.fasta (%d targets, names prefixed per copy)" % (reps, 153325 * reps)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    return out


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
    from biobench2_b200 import engine, search
    from biobench2_b200.oprofile import OProfile
    from biobench2_b200.seqdb import PackedDB

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libp7gpu has no CPU fallback (use --impl reference for the CPU arm)")
    # the whole-application leg runs two other processes; it goes first, before this process holds a CUDA context and
    # gigabytes of device memory (a second context on a busy device starts up to a second slower)
    e2e_app = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_app:
        try:
            cores = os.cpu_count() or 1
            if hasattr(os, "sched_getaffinity"):
                cores = len(os.sched_getaffinity(0))
            e2e_app = app_e2e(ROOT, cores, args.app_reps)
        except Exception as e:
            e2e_app = {"error": repr(e)}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dmod = dist if world > 1 else None

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
    assert db.nres == nres, (db.nres, nres)                      # counters.nres == sum of the lengths, whatever the padding
    cells = float(M) * float(nres)
    # survivors of the P-value test (p7_pipeline.c:691-693) are compacted by the filter kernels themselves
    F1 = 0.02
    last = {}

    class MsvBuffers:
        """device outputs of one database: score + status per target, the pass list + its count, the exchange capacity"""
        def __init__(self, n):
            self.n = n
            self.sc = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
            self.st = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            self.passed = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            self.count = torch.zeros(1, dtype=torch.int32, device=dev)
            self.cap = min(max(n, 1), max(1024, int(n * 0.06)))    # survivor records per rank and step: fixed capacity (F1 = 2 %)
            self.ar = torch.arange(self.cap, device=dev)

    def msv_step(dbx, b, first_global):
        """One MSV pass over dbx and the multi-GPU exchange of its survivors (index, score) to rank 0: the product's
        search.gather_device_records — fixed-capacity buffers, no host synchronisation inside the step."""
        engine.msvfilter_device(ctx, prof, dbx, b.sc.data_ptr(), b.st.data_ptr(), F1, b.passed.data_ptr(), b.count.data_ptr())
        cnt = b.count.to(torch.int64)
        idx = b.passed[:b.cap].long().clamp_(0, max(b.n - 1, 0))
        valid = (b.ar < cnt).unsqueeze(1)
        rec = torch.stack([(idx + first_global).double(), b.sc[idx].double()], 1) * valid
        last["out"], last["cnts"] = search.gather_device_records(rec, cnt, dmod)

    def timed_steps(dbx, b, first_global, steps, warmup):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                msv_step(dbx, b, first_global)
            barrier()
            ctx.set_timing(True); ctx.reset_timing()
            l0 = ctx.launches
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                msv_step(dbx, b, first_global)
            e1.record(stream)
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            km = (ctx.kernel_time(engine.KERNEL_SSV), ctx.kernel_time(engine.KERNEL_MSVJ))
            ctx.set_timing(False)
            nsurv = int(sum(int(c.item()) for c in last["cnts"]))  # read once, after the last step
            assert all(int(c.item()) <= b.cap for c in last["cnts"]), "survivor capacity exceeded"
        return ms, ctx.launches - l0, km, nsurv

    bufs = MsvBuffers(args.nseq)
    clocks = ClockSampler(physical_gpu_index(local_rank)); clocks.start()
    ms_total, launches, ((ssv_ms, ssv_n), (msvj_ms, msvj_n)), nsurv = timed_steps(db, bufs, rank * args.nseq, args.steps, args.warmup)
    clk = clocks.finish()
    total_cells = sum_over_ranks(cells)
    value = total_cells * args.steps / (ms_total * 1e-3) / 1e9
    statuses = torch.bincount(bufs.st.clamp(min=0), minlength=20).cpu().numpy()

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
    pk_instr = int_peak / 2.0
    tm_path = 193 <= M <= 1024           # ssv_msv.cu: launch_ssv takes the Tensor Memory table there (option ssv_tmem, default on)
    roofline = {"bound": "int16x2-issue", "kernel": "ssv_tm_kernel (gain table in Tensor Memory, one tcgen05.ld per row)" if tm_path else "ssv_kernel (gain table in shared memory)",
                "achieved": ssv_gcups * OPS_PER_CELL["ssv"], "peak": pk_instr, "unit": "G packed int16x2 instr/s (VIADDMNMX.S16x2)",
                "frac": ssv_gcups * OPS_PER_CELL["ssv"] / pk_instr if pk_instr else None,
                "instr_per_cell": OPS_PER_CELL["ssv"], "int16_ops_per_cell": 2, "kernel_gcups": ssv_gcups, "avg_launch_ms": ssv_avg_ms,
                "peak_source": "measured live in this run (p7gpu_measure_peaks: dependent-free VIADDMNMX.S16x2 chains; %.1f T int16 ops/s)" % (int_peak / 1e3),
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/ssv_traffic.json)",
                "smem": {"bound": "smem-lds", "achieved": ssv_gcups * 2.0, "peak": lds_peak, "unit": "GB/s (2 table bytes per cell, LDS.32)",
                         "frac": ssv_gcups * 2.0 / lds_peak if lds_peak else None,
                         "note": ("not this kernel's roof: the table is read from Tensor Memory (64 B/clk per SM sub-partition, 3 % busy); shown as the "
                                  "roof the shared-memory kernel sat on") if tm_path else "binding together with the integer pipe"},
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes": alg_bytes, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                        "note": "the path moves 1 byte per M cells: HBM is not the binding roof (0.6 % of peak)"}}

    # ---- strong scaling: ONE fixed database sharded by residue count over the ranks ----
    strong = None
    if args.strong_total > 0:
        tot_n = args.strong_total
        all_len = synth_lengths_np(tot_n, args.mean_len, 4242)
        span = (all_len.astype(np.int64) + 16) // 16 * 16
        all_off = np.concatenate([[0], np.cumsum(span)[:-1]]) + 16
        first = engine.shard_plan(PackedDB(np.zeros(1, np.uint8), all_off, all_len), world)
        lo, hi = int(first[rank]), int(first[rank + 1])
        with torch.cuda.stream(stream):
            # the shard's residues are generated in place from the shard's own seeded stream (iid background residues)
            s_res, s_off, s_len, s_lengths, _, _ = synth_db_device(torch, dev, hi - lo, args.mean_len, 4242 + 1000 * world + rank,
                                                                   lengths=all_len[lo:hi])
        stream.synchronize()
        sdb = engine.SeqDB.adopt_torch(ctx, s_res, s_off, s_len, int(s_lengths.max()))
        s_cells = float(M) * float(all_len.astype(np.int64).sum())
        s_ms, _, _, s_surv = timed_steps(sdb, MsvBuffers(hi - lo), lo, args.steps, args.warmup)
        strong = {"scaling": "strong", "targets_total": tot_n, "targets_this_rank0": int(first[1] - first[0]), "residues_total": int(all_len.astype(np.int64).sum()),
                  "value": s_cells * args.steps / (s_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": s_ms / args.steps,
                  "shards": "contiguous, balanced by residue count (p7gpu_shard_plan)", "msv_survivors": s_surv}
        sdb.close(); del s_res, s_off, s_len

    # ---- the configs of BASELINE.json beyond the MSV pass, per-stage kernel rates, per-kernel rooflines ----
    KINDS = (("ssv", engine.KERNEL_SSV), ("msvj", engine.KERNEL_MSVJ), ("bias", engine.KERNEL_BIAS), ("vit", engine.KERNEL_VIT),
             ("fwd", engine.KERNEL_FWD), ("bwd", engine.KERNEL_BWD), ("region_trigger", engine.KERNEL_TRIG),
             ("fwdx", engine.KERNEL_FWDX), ("bwdx", engine.KERNEL_BWDX))
    STAGE_OF = {"ssv": "msv", "bias": "bias", "vit": "vit", "fwd": "fwd", "bwd": "bwd"}

    def timed_pipeline(dbx, prof=prof, M=M, warm=True, **kw):
        if warm:
            engine.pipeline_run(ctx, prof, dbx, **kw).close()               # allocator pool, per-L tables
        barrier()
        ctx.set_timing(True); ctx.reset_timing()
        t0 = time.perf_counter()
        r = engine.pipeline_run(ctx, prof, dbx, **kw)
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        km = {k: ctx.kernel_time(kind)[0] for k, kind in KINDS}
        ctx.set_timing(False)
        rates = {}
        for k, st_name in STAGE_OF.items():
            if km[k] > 0 and r.stage_residues[st_name] > 0:
                if st_name == "bias":
                    rates["bias_gres_per_s"] = r.stage_residues["bias"] / (km[k] * 1e-3) / 1e9
                else:
                    rates[st_name + "_gcups"] = M * r.stage_residues[st_name] / (km[k] * 1e-3) / 1e9
        wall = max_over_ranks(wall)
        return r, wall, km, rates

    configs = None; roofline_kernels = None
    if not args.no_stages:
        configs = {}
        rates_all = {}
        # configs[2]: the full pipeline at default thresholds, same profile, same resident database; fb_exact = product default
        r, wall, km, rates = timed_pipeline(db, want_specials=True, region_trigger=True, fb_exact=True)
        rf, wallf, kmf, ratesf = timed_pipeline(db, want_specials=True, fb_exact=False)
        rates_all.update(ratesf); rates_all.update({k: v for k, v in rates.items() if k in ("vit_gcups", "bias_gres_per_s")})
        configs["config3_full_pipeline"] = {
            "workload": "full pipeline MSV->bias->Viterbi->Forward->Backward, default thresholds (F1=0.02,F2=1e-3,F3=1e-5), M=%d, "
                        "%d resident targets per GPU (BASELINE.json configs[2]); Forward/Backward results that reach the host bit-identical to the reference (fb_exact)" % (M, args.nseq),
            "ms": wall * 1e3, "gcups_whole_db": sum_over_ranks(cells) / wall / 1e9, "kernel_ms": km, "kernel_ms_sum": sum(km.values()),
            "counters": [int(x) for x in r.counters], "survivors": len(r.survivors), "stage_residues": r.stage_residues, "stage_kernel_rates": rates,
            "fast_specials_variant": {"ms": wallf * 1e3, "kernel_ms": kmf, "counters": [int(x) for x in rf.counters],
                                      "same_counters": [int(x) for x in rf.counters] == [int(x) for x in r.counters]}}
        r.close(); rf.close()
        # a slice for the stages a random database starves at default thresholds
        nsub = min(args.nseq, 100_000)
        sub_lengths = lengths[:nsub]
        sub_bytes = int(offsets[nsub - 1] + sub_lengths[-1] + 1) if nsub else 1
        sub_res = torch.empty(sub_bytes + 32, dtype=torch.uint8, device=dev); sub_res[:sub_bytes] = res[:sub_bytes]; sub_res[sub_bytes:] = 255
        sub_db = engine.SeqDB.adopt_torch(ctx, sub_res, d_off[:nsub].contiguous(), d_len[:nsub].contiguous(), int(sub_lengths.max()))
        sub_resid = float(sub_lengths.astype(np.int64).sum())
        ctx.set_timing(True); ctx.reset_timing()
        engine.vitfilter_batch(ctx, prof, sub_db)
        vms = ctx.kernel_time(engine.KERNEL_VIT)[0]; ctx.set_timing(False)
        rates_all["vit_gcups"] = M * sub_resid / (vms * 1e-3) / 1e9
        for mode, exact in (("fast", False), ("exact", True)):               # Forward+Backward on every target of the slice
            r2, wall2, km2, rates2 = timed_pipeline(sub_db, do_max=True, want_specials=False, region_trigger=True, fb_exact=exact)
            if exact:
                if km2["fwdx"] > 0: rates_all["fwdx_gcups"] = M * sub_resid / (km2["fwdx"] * 1e-3) / 1e9
                if km2["bwdx"] > 0: rates_all["bwdx_gcups"] = M * sub_resid / (km2["bwdx"] * 1e-3) / 1e9
            else:
                rates_all.update({k: v for k, v in rates2.items() if k in ("fwd_gcups", "bwd_gcups")})
            r2.close()
        roofline_kernels = {"M": M, "sample": "Viterbi/Forward/Backward: every target of a %d-target slice (--max); bias: the MSV survivors of the full database" % nsub,
                            "peaks": {"int16_gops": int_peak, "fp32_nofma_gflops": fp32_peak, "fp32_fma_gflops": 2.0 * fp32_peak, "lds_gbs": lds_peak,
                                      "source": "measured live in this run (p7gpu_measure_peaks)"}}
        roofline_kernels.update(kernel_rooflines(M, rates_all, int_peak, fp32_peak, lds_peak))
        # configs[3]: long-model stress, M=2000, --max, at the FULL database size: Forward + Backward + region trigger on every
        # target; survivor records (scores, region flags) come back, specials do not (they would be 170 GB)
        p2000 = os.path.join(ROOT, "tests", "golden", "prof_synth_M2000.npz")
        if os.path.exists(p2000):
            op4 = OProfile.load(p2000); prof4 = engine.Profile(ctx, op4)
            r4, wall4, km4, rates4 = timed_pipeline(db, prof=prof4, M=op4.M, warm=False, do_max=True, want_specials=False, region_trigger=True, fb_exact=False)
            ksum = km4["fwd"] + km4["bwd"] + km4["ssv"] + km4["msvj"] + km4["region_trigger"]
            c4 = {"workload": "long-model stress (BASELINE.json configs[3]): synth_M2000, --max (all filters off: MSV, Forward, Backward and the region "
                              "trigger on every one of the %d targets per GPU); survivor records returned, specials stay on the device" % args.nseq,
                  "ms": wall4 * 1e3, "kernel_ms": km4, "kernel_ms_sum": ksum, "wall_over_kernel": wall4 * 1e3 / ksum if ksum else None,
                  "gcups_fwd_plus_bwd_wall": sum_over_ranks(2.0 * op4.M * r4.stage_residues["fwd"]) / wall4 / 1e9,
                  "stage_kernel_rates": rates4, "counters": [int(x) for x in r4.counters],
                  "region_flagged": int((r4.survivors["region_flag"] == 1).sum()), "fb": "fast kernels (<= 1e-4 relative)"}
            c4["roofline"] = kernel_rooflines(op4.M, rates4, int_peak, fp32_peak, lds_peak)
            r4.close()
            # the same with the exact-order kernels and the specials of the flagged targets returned, on the slice
            r5, wall5, km5, _ = timed_pipeline(sub_db, prof=prof4, M=op4.M, do_max=True, want_specials=True, region_trigger=True, fb_exact=True)
            c4["exact_slice"] = {"workload": "same on the first %d targets with fb_exact: specials bit-identical to the reference, returned for flagged targets" % nsub,
                                 "ms": wall5 * 1e3, "kernel_ms": km5,
                                 "fwdx_gcups": op4.M * sub_resid / (km5["fwdx"] * 1e-3) / 1e9 if km5["fwdx"] > 0 else None,
                                 "bwdx_gcups": op4.M * sub_resid / (km5["bwdx"] * 1e-3) / 1e9 if km5["bwdx"] > 0 else None,
                                 "region_flagged": int((r5.survivors["region_flag"] == 1).sum()), "specials_floats_returned": int(r5.fwd_xmx.shape[0] * 2)}
            r5.close()
            configs["config4_long_model_max"] = c4
            prof4.close()
        sub_db.close()
        # configs[4]: hmmscan-style batch — many Pfam-shaped profiles (M=100..800) against the resident, sharded database
        import glob
        paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "scan", "prof_scan_M*.npz")))
        if paths and args.config5_passes > 0:
            ops5 = [OProfile.load(p) for p in paths]
            t0 = time.perf_counter(); profs5 = [engine.Profile(ctx, o5) for o5 in ops5]; upload_ms = (time.perf_counter() - t0) * 1e3 / len(profs5)
            for r_ in engine.pipeline_run_many(ctx, profs5, db, want_specials=False, region_trigger=True, fb_exact=True):
                r_.close()
            batch = [profs5[i % len(profs5)] for i in range(args.config5_passes)]
            Ms = [ops5[i % len(ops5)].M for i in range(args.config5_passes)]
            barrier(); t0 = time.perf_counter()
            rs = engine.pipeline_run_many(ctx, batch, db, want_specials=False, region_trigger=True, fb_exact=True)
            barrier(); dt5 = max_over_ranks(time.perf_counter() - t0)
            c5 = np.sum([r_.counters.astype(np.int64) for r_ in rs], axis=0)
            configs["config5_many_profiles"] = {
                "workload": "hmmscan-style batch (BASELINE.json configs[4]): %d profile passes (%d calibrated profiles, M=%d..%d, cycled) x %d resident "
                            "targets per GPU, full pipeline at default thresholds, database sharded by sequence (%d rank(s): %d targets in all)"
                            % (args.config5_passes, len(ops5), min(o5.M for o5 in ops5), max(o5.M for o5 in ops5), args.nseq, world, args.nseq * world),
                "passes": args.config5_passes, "seconds": dt5, "ms_per_profile_pass": dt5 / args.config5_passes * 1e3,
                "gcups_whole_pipeline": sum_over_ranks(float(sum(Ms)) * nres) / dt5 / 1e9, "mean_M": float(np.mean(Ms)),
                "profile_upload_ms_each": upload_ms, "counters_sum_this_rank": [int(x) for x in c5],
                "extrapolated_seconds_for_1000_profiles": dt5 / args.config5_passes * 1000.0}
            for r_ in rs:
                r_.close()
            for p5 in profs5:
                p5.close()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if args.e2e_steps > 0:
        h_res = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_res.copy_(res[:nbytes]); torch.cuda.synchronize()
        h_off = torch.empty(args.nseq, dtype=torch.int64, pin_memory=True); h_off.copy_(torch.from_numpy(offsets))
        h_len = torch.empty(args.nseq, dtype=torch.int32, pin_memory=True); h_len.copy_(torch.from_numpy(lengths))
        hdb = PackedDB(h_res.numpy(), h_off.numpy(), h_len.numpy())
        # the database in its 5-bit transfer format: built ONCE PER DATABASE (like the packed byte form itself, like makehmmerdb),
        # outside the timed region; from FASTA text the one-off preparation costs cpu_baseline.hmmsearch_cli.fasta_reader's rate
        t0 = time.perf_counter()
        h_p5 = torch.empty(int(engine.lib().p7gpu_pack5_bytes(nbytes)), dtype=torch.uint8, pin_memory=True)
        engine.pack5(hdb.residues, out=h_p5.numpy())
        pack5_s = time.perf_counter() - t0

        e2e_out = {}

        def e2e_step():
            r = engine.pipeline_run_host5(ctx, prof, h_p5.numpy(), hdb, F1=F1, stop_after_msv=True)
            e2e_out["survivors"] = len(r.msv_pass); e2e_out["counters"] = [int(x) for x in r.counters]
            r.close()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # one more, untimed step with per-launch CUDA events: how much of the streamed step the SSV launches themselves take
        ctx.set_timing(True); ctx.reset_timing()
        t1 = time.perf_counter(); e2e_step(); wall_probe = time.perf_counter() - t1
        ssv_ms, ssv_n = ctx.kernel_time(engine.KERNEL_SSV); msvj_ms, _ = ctx.kernel_time(engine.KERNEL_MSVJ)
        ctx.set_timing(False)
        h2d = int(h_p5.numel() + args.nseq * 12)
        e2e = {"value": total_cells * args.e2e_steps / dt / 1e9, "unit": "GCUPS", "steps": args.e2e_steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(e2e_out["survivors"] * 24 + 4),
               "ms_per_step": dt / args.e2e_steps * 1e3, "h2d_gbs_per_rank": h2d / (dt / args.e2e_steps) / 1e9,
               "survivors": e2e_out["survivors"], "counters": e2e_out["counters"],
               "kernel_ms_one_step": {"ssv_sum_over_chunk_launches": ssv_ms, "ssv_launches": ssv_n, "msvj": msvj_ms, "wall_ms": wall_probe * 1e3,
                                      "note": "launches of consecutive chunks run on two alternating streams and may overlap, so the sum can exceed the busy time"},
               "once_per_database_prep": {"pack5_seconds": pack5_s, "note": "5-bit transfer form built once per database, outside the timed region; "
                                          "digitising FASTA text is a further one-off (see cpu_baseline.hmmsearch_cli.fasta_reader and e2e_app, which include it)"},
               "api": "p7gpu_pipeline_run_host5(stop_after_msv): pinned host packed DB in the 5-bit transfer format (8 residues in 5 bytes, expanded on the device) streamed in 16-256 MB chunks (database order; the last chunk longest-first) under the "
                      "SSV/MSV kernels; the MSV survivor records (index, L, status, score) come back to the host"}
        del h_res, h_p5

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
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "roofline_kernels": roofline_kernels,
                "cpu_baseline": cpu_baseline, "configs": configs, "strong": strong, "e2e_app": e2e_app,
                "detail": {"cells_per_step_per_gpu": cells, "residues_per_gpu": nres, "M": M,
                           "ssv_launches": int(ssv_n), "ssv_ms_total": ssv_ms, "msvj_launches": int(msvj_n), "msvj_ms_total": msvj_ms,
                           "msv_survivors_all_ranks_last_step": nsurv,
                           "exchange": "search.gather_device_records: fixed-capacity survivor buffers, one all_gather + one gather per step, no host sync in the timed region",
                           "status_counts": {"ok": int(statuses[0]), "erange": int(statuses[16]), "noresult": int(statuses[19])}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
