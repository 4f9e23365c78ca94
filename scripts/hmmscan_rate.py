#!/usr/bin/env python
"""hmmscan_b200 against the reference's hmmscan on a pressed database of N sampled, calibrated profiles and a few queries:
wall clock of both, outputs compared, and the engine's own per-model time (--times).
Usage: hmmscan_rate.py [nmodels] [nqueries] [extra hmmscan options, e.g. --F1 1e-7 so that only the first stage runs]
P7SCAN_DB=<dir>: keep / reuse the pressed database in that directory (built once, e.g. in the authoring container).
P7SCAN_B200_OPTS="--prefilter 0": extra options for hmmscan_b200 only."""
import os, sys, subprocess, tempfile, time, json, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

def main():
    nm = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    more = sys.argv[3:]
    from oracle import ref
    tmp = tempfile.mkdtemp(prefix="p7scan_")
    keep = os.environ.get("P7SCAN_DB")
    if keep:
        os.makedirs(keep, exist_ok=True)
    db = os.path.join(keep or tmp, "scan%d.hmm" % nm)
    rng = np.random.default_rng(3)
    t0 = time.perf_counter()
    if not os.path.exists(db + ".h3f"):
        with open(db, "wb") as f:
            for k in range(nm):
                M = int(rng.integers(40, 420))
                rp = ref.RefProfile.sample(M, 1000 + k, calibrate=True)
                one = os.path.join(tmp, "m.hmm"); rp.write_hmm(one); rp.close()
                txt = open(one, "rb").read().replace(b"NAME  ", b"NAME  m%05d_" % k, 1)
                f.write(txt)
        p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "hmmpress"), db], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
    t_build = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    fa = os.path.join(tmp, "q.fa")
    with open(fa, "w") as f:
        aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
        if os.environ.get("P7SCAN_UNIFORM"):                  # uniform residue frequencies: far from the null model the filters are
            for q in range(nq):                               # calibrated on, so most models get past MSV (a stress case)
                L = int(rng.integers(150, 600))
                f.write(">q%d\n%s\n" % (q, "".join(aa[rng.integers(0, 20, L)])))
        else:                                                 # background (Swiss-Prot) frequencies: the null model's own composition
            from biobench2_b200 import seqdb
            qdb = seqdb.synth_db(nq, mean=350, seed=5)
            for q in range(nq):
                o, L = int(qdb.offsets[q]), int(qdb.lengths[q])
                f.write(">q%d\n%s\n" % (q, "".join(aa[qdb.residues[o:o + L]])))
    out = {"nmodels": nm, "nqueries": nq, "options": more, "profile_db_build_s": round(t_build, 1)}
    res = {}
    for name, exe, extra in (("reference", os.path.join(ROOT, "oracle", "_ref", "hmmscan"), ["--cpu", "16"]),
                             ("b200", os.path.join(ROOT, "integration", "_build", "hmmscan_b200"), ["--times"] + os.environ.get("P7SCAN_B200_OPTS", "").split())):
        best = None
        for rep in range(2):
            t = time.perf_counter()
            p = subprocess.run([exe, "-o", os.path.join(tmp, name + ".out"), "--tblout", os.path.join(tmp, name + ".tbl")] + extra + more + [db, fa],
                               capture_output=True, text=True)
            dt = time.perf_counter() - t
            assert p.returncode == 0, (name, p.stderr[-800:])
            best = dt if best is None else min(best, dt)
        out[name + "_wall_s"] = round(best, 3)
        if name == "b200":
            out["b200_phases"] = [l[2:] for l in p.stderr.split("\n") if l.startswith("# ")]
            m = re.search(r"models: ([0-9.]+) s", p.stderr)
            if m:
                out["b200_us_per_model"] = round(float(m.group(1)) / nm * 1e6, 1)
        res[name] = [l for l in open(os.path.join(tmp, name + ".tbl")).read().split("\n") if not l.startswith("#")]
    out["b200_options"] = os.environ.get("P7SCAN_B200_OPTS", "")
    out["queries"] = "uniform residues" if os.environ.get("P7SCAN_UNIFORM") else "background frequencies, mean length 350"
    out["tblout_identical"] = res["reference"] == res["b200"]
    print(json.dumps(out))

main()
