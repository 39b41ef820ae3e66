"""GPU debug helper: re-runs one configuration of tests/test_gpu_random.py with the fast and the generic k_score task bodies and prints
where the f64 fields differ from the oracle (usage: python tools/debug_seed.py 23)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_random as T
from helpers import oracle_cfg, oracle_db_from_peptides
from sage_b200 import IndexedDatabase, Scorer, synth

seed = int(sys.argv[1])
rng = np.random.default_rng(7000 + seed)
pep = synth.make_peptides(int(rng.choice([600, 3000, 9000])), seed=100 + seed, static_c=bool(rng.integers(2)), var_mod_m=bool(rng.integers(2)))
kinds = [("b", "y"), ("b", "y"), ("a", "b", "y"), ("c", "z"), ("y",), ("b", "x", "y")][int(rng.integers(6))]
bucket = int(rng.choice([16, 256, 4096, 8192, 32768])); min_ion = int(rng.choice([0, 1, 2, 2, 3]))
odb = oracle_db_from_peptides(pep, bucket_size=bucket, ion_kinds=kinds, min_ion_index=min_ion)
rng.random()
gdb = IndexedDatabase.build_from_peptides(pep, bucket_size=bucket, ion_kinds=kinds, min_ion_index=min_ion)
spectra = T.random_spectra(pep, rng, 160)
iso = [(0, 0), (0, 0), (-1, 3), (0, 1), (2, 2)][int(rng.integers(5))]
kw = dict(precursor_tol=T.random_tolerance(rng, True), fragment_tol=T.random_tolerance(rng, False), min_matched_peaks=int(rng.choice([0, 1, 4, 6])),
          min_isotope_err=iso[0], max_isotope_err=iso[1], min_precursor_charge=int(rng.choice([1, 2])), max_precursor_charge=int(rng.choice([3, 4, 5])),
          override_precursor_charge=bool(rng.random() < 0.2), max_fragment_charge=[None, None, 1, 2, 3][int(rng.integers(5))],
          chimera=bool(rng.random() < 0.3), report_psms=int(rng.choice([1, 2, 5, 30])), wide_window=bool(rng.random() < 0.2),
          score_type=int(rng.random() < 0.2))
print(kinds, bucket, min_ion, kw)
of, oc, _, _ = odb.score_batch(oracle_cfg(**kw), spectra.as_dict())
r = kw["report_psms"]
for fast in (1, 0):
    sc = Scorer(gdb, **kw); sc.set_option("score_fast", fast)
    gf, gc = sc.score_batch(spectra)
    sel = (np.arange(len(gf)) % r) < np.repeat(gc, r)
    print("fast", fast, "counts equal", np.array_equal(gc, oc))
    for f in ("hyperscore", "delta_next", "delta_best", "poisson"):
        a, b = gf[f][sel], of[f][sel]
        d = np.nonzero(a.view(np.uint64) != b.view(np.uint64))[0]
        print("  ", f, "bitwise differing rows", len(d), [(float(a[i]).hex(), float(b[i]).hex()) for i in d[:4]])
