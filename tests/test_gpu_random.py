"""Randomised parity sweep: random index shapes (bucket size, ion kinds, min_ion_index, mods), random Scorer settings and ragged
spectra, CUDA path vs oracle. Seeds are fixed, so failures reproduce."""
import numpy as np
import pytest

from sage_b200 import IndexedDatabase, Scorer, SpectraBatch, Tolerance, synth

from helpers import assert_features_equal, f64_exact_default, oracle_cfg, oracle_db_from_peptides

pytestmark = pytest.mark.gpu


def random_spectra(pep, rng, n):
    """Ragged spectra: peak counts 0..400 (sub-sampled / padded), some unknown charges, some isolation windows."""
    base = synth.make_spectra(pep, n, seed=int(rng.integers(1 << 30)), n_peaks=200)
    m = base.masses.reshape(n, 200)
    it = base.intensities.reshape(n, 200)
    masses, intens, off = [], [], [0]
    for i in range(n):
        k = int(rng.choice([0, 1, 7, 60, 200, 200, 200, 333]))
        if k <= 200:
            sel = np.sort(rng.choice(200, size=k, replace=False))
            mm, ii = m[i][sel], it[i][sel]
        else:
            extra = np.sort(rng.uniform(100, 1800, k - 200).astype(np.float32))
            mm = np.concatenate([m[i], extra])
            ii = np.concatenate([it[i], rng.lognormal(7, 1, k - 200).astype(np.float32)])
            o = np.argsort(mm, kind="stable")
            mm, ii = mm[o], ii[o]
        masses.append(mm)
        intens.append(ii)
        off.append(off[-1] + len(mm))
    masses = np.concatenate(masses).astype(np.float32) if off[-1] else np.zeros(0, np.float32)
    intens = np.concatenate(intens).astype(np.float32) if off[-1] else np.zeros(0, np.float32)
    tic = np.array([np.cumsum(intens[off[i]:off[i + 1]], dtype=np.float32)[-1] if off[i + 1] > off[i] else 0.0 for i in range(n)], np.float32)
    chg = np.where(rng.random(n) < 0.25, 0, base.prec_charge).astype(np.uint8)
    iso = rng.random(n) < 0.5
    ilo = np.where(iso, -rng.uniform(0.5, 3.0, n), np.nan).astype(np.float32)
    ihi = np.where(iso, rng.uniform(0.5, 3.0, n), np.nan).astype(np.float32)
    return SpectraBatch(peak_off=np.array(off, np.uint64), masses=masses, intensities=intens, prec_mz=base.prec_mz, prec_charge=chg, iso_lo=ilo,
                        iso_hi=ihi, tic=tic, level=base.level, rt=base.rt, ims=np.where(rng.random(n) < 0.3, rng.random(n), np.nan).astype(np.float32))


def random_tolerance(rng, precursor):
    kind = rng.choice(["ppm", "da", "pct"], p=[0.6, 0.3, 0.1])
    if kind == "ppm":
        w = float(rng.choice([5, 10, 20, 50, 100])) if not precursor else float(rng.choice([5, 20, 50, 500]))
        return Tolerance.ppm(-w * rng.uniform(0.5, 1.0), w * rng.uniform(0.5, 1.0))
    if kind == "da":
        w = float(rng.choice([0.01, 0.05, 0.5])) if not precursor else float(rng.choice([0.5, 3.0, 50.0, 600.0]))
        return Tolerance.da(-w * rng.uniform(0.2, 1.0), w * rng.uniform(0.2, 1.0))
    w = 0.002 if not precursor else float(rng.choice([0.01, 0.5]))
    return Tolerance.pct(-w, w)


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed):
    rng = np.random.default_rng(7000 + seed)
    pep = synth.make_peptides(int(rng.choice([600, 3000, 9000])), seed=100 + seed, static_c=bool(rng.integers(2)), var_mod_m=bool(rng.integers(2)))
    kinds = [("b", "y"), ("b", "y"), ("a", "b", "y"), ("c", "z"), ("y",), ("b", "x", "y")][int(rng.integers(6))]
    bucket = int(rng.choice([16, 256, 4096, 8192, 32768]))
    min_ion = int(rng.choice([0, 1, 2, 2, 3]))
    odb = oracle_db_from_peptides(pep, bucket_size=bucket, ion_kinds=kinds, min_ion_index=min_ion)
    if rng.random() < 0.5:
        gdb = IndexedDatabase.build_from_peptides(pep, bucket_size=bucket, ion_kinds=kinds, min_ion_index=min_ion)
    else:
        e = odb.export()
        gdb = IndexedDatabase.from_reference_layout(pep, e["frag_pep"], e["frag_mz"], e["bucket_min"], e["bucket_size"], ion_kinds=kinds)
    spectra = random_spectra(pep, rng, 160)
    iso = [(0, 0), (0, 0), (-1, 3), (0, 1), (2, 2)][int(rng.integers(5))]
    kw = dict(precursor_tol=random_tolerance(rng, True), fragment_tol=random_tolerance(rng, False), min_matched_peaks=int(rng.choice([0, 1, 4, 6])),
              min_isotope_err=iso[0], max_isotope_err=iso[1], min_precursor_charge=int(rng.choice([1, 2])), max_precursor_charge=int(rng.choice([3, 4, 5])),
              override_precursor_charge=bool(rng.random() < 0.2), max_fragment_charge=[None, None, 1, 2, 3][int(rng.integers(5))],
              chimera=bool(rng.random() < 0.3), report_psms=int(rng.choice([1, 2, 5, 30])), wide_window=bool(rng.random() < 0.2),
              score_type=int(rng.random() < 0.2))
    sc = Scorer(gdb, **kw)
    sc.set_option("pep_cap", int(rng.choice([0, 64, 8192])))
    sc.set_option("wide_tile", int(rng.choice([512, 4096, 32768])))
    sc.set_option("pipeline_chunks", int(rng.choice([1, 2, 7])))
    sc.set_option("score_fast", int(rng.integers(2)))   # straight-line vs generic task body of k_score
    gf, gc = sc.score_batch(spectra)
    of, oc, _, _ = odb.score_batch(oracle_cfg(**kw), spectra.as_dict())
    assert_features_equal(gf, gc, of, oc, kw["report_psms"], what=f"seed {seed}: kinds={kinds} bucket={bucket} min_ion={min_ion} {kw}",
                          f64_exact=f64_exact_default(kw["score_type"]))
    # idempotence: a second pass over the same resident batch returns the same bytes
    gf2, gc2 = sc.score_batch(spectra)
    sel = (np.arange(len(gf)) % kw["report_psms"]) < np.repeat(gc, kw["report_psms"])
    assert np.array_equal(gc, gc2) and gf[sel].tobytes() == gf2[sel].tobytes()
