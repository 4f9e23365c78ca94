#!/usr/bin/env python
"""Generate tests/golden/scan/prof_scan_M*.npz: calibrated synthetic profiles spread over M = 100..800 for the
hmmscan-style batch (BASELINE.json configs[4]), by running the UNMODIFIED reference (oracle/_ref/libp7ref.so:
p7_hmm_Sample + p7_hmm_SetComposition + p7_Calibrate, SURVEY.md §8d recipe (i)).  Run in the authoring container:
    make -C oracle ref && python tests/golden/make_scan_profiles.py
The outputs are committed; nothing at test/bench time reads /root/reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref                        # noqa: E402
from biobench2_b200.oprofile import OProfile  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "scan")
MS = [100, 127, 160, 200, 255, 320, 383, 450, 520, 610, 700, 800]

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for M in MS:
        p = ref.RefProfile.sample(M, 1000 + M, calibrate=True)
        t, pi, eo = p.filter_hmm()
        op = OProfile.from_view(p.view(), name="scan_M%d" % M, bg_filter=(t, pi, eo, p.omega()))
        op.save(os.path.join(OUT, "prof_scan_M%d.npz" % M))
        print("scan_M%d  evparam %s" % (M, [round(x, 3) for x in op.evparam]))
        p.close()
