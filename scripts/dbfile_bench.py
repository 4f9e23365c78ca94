#!/usr/bin/env python
"""Measurement for the packed on-disk database: FASTA digitiser on all host cores and on one, then a search
from the mapped file.  (The comparison with the reference's own reader is part of bench.py's cpu_baseline leg.)
Usage: dbfile_bench.py [nseq]"""
import os, sys, json, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biobench2_b200 import engine, seqdb
from biobench2_b200.oprofile import OProfile
from bench import write_fasta


def main():
    nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    hdb = seqdb.synth_db(nseq, mean=350, seed=99)
    tmp = tempfile.mkdtemp(prefix="p7db_")
    fa = os.path.join(tmp, "db.fa"); out = os.path.join(tmp, "db.p7db")
    write_fasta(hdb, fa)
    fbytes = os.path.getsize(fa)
    res = {"nseq": nseq, "nres": hdb.nres, "fasta_bytes": fbytes, "host_threads": os.cpu_count()}
    t0 = time.perf_counter(); n, r = engine.DbFile.from_fasta(fa, out, 0); dt = time.perf_counter() - t0
    assert n == nseq and r == hdb.nres
    res["digitise_s"] = dt; res["digitise_Mres_per_s"] = hdb.nres / dt / 1e6; res["digitise_MB_per_s"] = fbytes / dt / 1e6
    t0 = time.perf_counter(); n1, r1 = engine.DbFile.from_fasta(fa, out + "1", 1); dt1 = time.perf_counter() - t0
    res["digitise_1thread_Mres_per_s"] = hdb.nres / dt1 / 1e6
    d = engine.DbFile(out)
    assert np.array_equal(d.packed().lengths, hdb.lengths) and (d.packed().offsets % 16 == 0).all()
    if engine.lib().p7gpu_device_count() > 0:
        ctx = engine.Context(0)
        prof = engine.Profile(ctx, OProfile.load(os.path.join(ROOT, "tests", "golden", "prof_synth_M400.npz")))
        t0 = time.perf_counter(); d.pin(); res["pin_s"] = time.perf_counter() - t0
        engine.pipeline_run_file(ctx, prof, d, want_specials=False).close()
        t0 = time.perf_counter(); rr = engine.pipeline_run_file(ctx, prof, d, want_specials=False); dts = time.perf_counter() - t0
        res["search_from_file_s"] = dts; res["search_gcups"] = 400.0 * hdb.nres / dts / 1e9; res["counters"] = [int(x) for x in rr.counters]
        rr.close(); prof.close(); ctx.close()
    d.close()
    print(json.dumps(res))


main()
