"""SpectrumProcessor::process (spectrum.rs:263-412) on the device vs the oracle's CPU restatement (which is pinned to the reference's exact
deisotope vectors in test_oracle_known_answers.py)."""
import numpy as np
import pytest

from sage_b200 import SpectrumProcessor
from oracle import oracle as O

pytestmark = pytest.mark.gpu
NEUTRON = 1.00335


def raw_spectra(rng, n):
    """Centroided raw spectra with isotope envelopes (z = 1..3), duplicates and noise; ascending m/z per spectrum."""
    mzs, its, off, chg = [], [], [0], []
    for _ in range(n):
        k = int(rng.choice([0, 1, 2, 40, 150, 300, 600]))
        mono = rng.uniform(150, 1800, max(k // 3, 1))
        z = rng.integers(1, 4, len(mono))
        mz = [mono]
        it = [rng.lognormal(9, 1, len(mono))]
        for iso in (1, 2):
            sel = rng.random(len(mono)) < (0.8 if iso == 1 else 0.4)
            mz.append((mono + iso * NEUTRON / z)[sel] * (1 + rng.normal(0, 2e-6, int(sel.sum()))))
            it.append(it[0][sel] * (0.6 if iso == 1 else 0.25) * rng.uniform(0.7, 1.2, int(sel.sum())))
        noise = k - sum(len(x) for x in mz)
        if noise > 0:
            mz.append(rng.uniform(150, 1800, noise))
            it.append(rng.lognormal(7, 1, noise))
        mz, it = np.concatenate(mz)[:k].astype(np.float32), np.concatenate(it)[:k].astype(np.float32)
        if k > 10:
            mz[5], it[5] = mz[4], it[4]          # exact duplicate peak
            it[7] = it[8]                         # equal intensities
        o = np.argsort(mz, kind="stable")
        mzs.append(mz[o]); its.append(it[o]); off.append(off[-1] + k); chg.append(int(rng.choice([0, 2, 3, 4])))
    return np.array(off, np.uint64), np.concatenate(mzs), np.concatenate(its), np.array(chg, np.uint8)


@pytest.mark.parametrize("top_n,deiso,min_mz", [(100, True, 0.0), (150, True, 0.0), (150, True, 400.0), (100, False, 0.0), (30, False, 0.0), (5000, False, 0.0),
                                                (1, True, 0.0)])
def test_process_matches_oracle(top_n, deiso, min_mz):
    rng = np.random.default_rng(500 + top_n + int(deiso))
    off, mz, it, chg = raw_spectra(rng, 120)
    sp = SpectrumProcessor(top_n, deiso, min_mz)
    goff, gm, gi, gt = sp.process_batch(off, mz, it, chg)
    for s in range(len(chg)):
        a, b = int(off[s]), int(off[s + 1])
        om, oi, ot = O.process_ms2(mz[a:b], it[a:b], int(chg[s]), top_n, deiso, min_mz)
        ga, gb = int(goff[s]), int(goff[s + 1])
        assert gb - ga == len(om), (s, gb - ga, len(om))
        assert np.array_equal(gm[ga:gb].view(np.uint32), om.view(np.uint32)), s
        assert np.array_equal(gi[ga:gb].view(np.uint32), oi.view(np.uint32)), s
        assert np.float32(gt[s]).view(np.uint32) == np.float32(ot).view(np.uint32), s


def test_process_config1_fixture(config1):
    # crates/sage-cli/tests/integration.rs:24-26: SpectrumProcessor::new(100, true, 0.0).process(spectra[0])
    off = np.array([0, len(config1["mz"])], np.uint64)
    chg = np.array([config1["precursor_charge"]], np.uint8)
    for top_n in (100, 150):
        goff, gm, gi, gt = SpectrumProcessor(top_n, True, 0.0).process_batch(off, config1["mz"], config1["intensity"], chg)
        om, oi, ot = O.process_ms2(config1["mz"], config1["intensity"], config1["precursor_charge"], top_n, True, 0.0)
        assert len(gm) == len(om) == top_n and np.array_equal(gm, om) and np.array_equal(gi, oi) and gt[0] == ot


def test_find_reporter_ions_matches_oracle():
    # tmt.rs:193-211; the -PROTON offset semantics are pinned by the reference test spectrum.rs:589-605 (in test_oracle_known_answers.py)
    from sage_b200 import Tolerance
    from sage_b200.api import TMT6PLEX, find_reporter_ions
    rng = np.random.default_rng(77)
    n = 300
    off, masses, intens = [0], [], []
    for i in range(n):
        k = int(rng.choice([0, 3, 50, 200]))
        m = rng.uniform(100, 1500, k)
        lab = TMT6PLEX[rng.random(6) < 0.7].astype(np.float64) - 1.0072764
        m = np.concatenate([m, lab * (1 + rng.normal(0, 3e-6, len(lab))), lab * (1 + rng.normal(0, 8e-6, len(lab)))])
        m = np.sort(m.astype(np.float32))
        masses.append(m)
        intens.append(rng.lognormal(8, 1, len(m)).astype(np.float32))
        off.append(off[-1] + len(m))
    masses, intens, off = np.concatenate(masses), np.concatenate(intens), np.array(off, np.uint64)
    for tol, otol in ((Tolerance.ppm(-20, 20), (O.PPM, -20.0, 20.0)), (Tolerance.da(-0.003, 0.003), (O.DA, -0.003, 0.003))):
        g = find_reporter_ions(off, masses, intens, TMT6PLEX, tol)
        o = O.find_reporter_ions(off, masses, intens, TMT6PLEX, otol)
        assert (g > 0).sum() > 500 and np.array_equal(g.view(np.uint32), o.view(np.uint32))
