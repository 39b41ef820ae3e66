#!/usr/bin/env python
"""k_score per-phase cycle shares (variant build with -DSAGE_B200_PHASE_CLOCKS=1; SAGE_B200_LIB must point at it).

    SAGE_B200_LIB=$PWD/sage_b200/lib/ab/phase.so python tools/phase_cycles.py [cfg2|cfg5]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sage_b200 import IndexedDatabase, Scorer, Tolerance, api, synth   # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
pep = synth.make_peptides(2_000_000)
spectra = synth.make_spectra(pep, 50_000, seed=0xB202, chimeric=(wl == "cfg5"))
db = IndexedDatabase.build_from_peptides(pep)
kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20))
if wl == "cfg5":
    kw.update(chimera=True, report_psms=5)
sc = Scorer(db, **kw)
sc.upload(spectra)
for _ in range(3):
    sc.run()
lib = api.load_library()
buf = (C.c_ulonglong * 16)()
lib.sage_b200_debug_phase_cycles(buf, 1)
steps = 5
for _ in range(steps):
    sc.run()
c = sc.counters()
lib.sage_b200_debug_phase_cycles(buf, 0)
v = np.array(list(buf), dtype=np.float64)
names = ["prologue (hits fold, headers)", "wait peaks (bulk copy)", "verify + LUT", "phase A (cand headers, scan)", "phase B + B' (lookups, hit pass)",
         "fold (+ records, hyperscore)", "rank + features + tail"]
tot = v[:7].sum()
print(f"{wl}: k_score {c['ms_score']:.3f} ms per step; cycles per spectrum-CTA {tot / steps / len(spectra):.0f}")
for i, n in enumerate(names):
    print(f"  {n:38s} {100 * v[i] / tot:5.1f}%  {v[i] / steps / len(spectra):8.0f} cyc")
