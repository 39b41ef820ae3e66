"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c). Each test names the reference test it restates."""
import numpy as np
import pytest

from oracle import oracle as O

NEUTRON = np.float32(1.00335)
PROTON = np.float32(1.0072764)
f32 = np.float32


def test_tolerances():
    # mass.rs:143-157 (exact f32 equality)
    assert O.tolerance_bounds(O.PPM, -10.0, 20.0, 1000.0) == (float(f32(999.99)), float(f32(1000.02)))
    assert O.tolerance_bounds(O.PPM, -10.0, 10.0, 487.0) == (float(f32(486.99513)), float(f32(487.00487)))
    assert O.tolerance_bounds(O.PPM, -50.0, 50.0, 1000.0) == (float(f32(999.95)), float(f32(1000.05)))
    assert O.tolerance_bounds(O.DA, -2.5, 1.5, 100.0) == (97.5, 101.5)
    assert O.tolerance_bounds(O.PCT, -1.0, 1.0, 200.0) == (198.0, 202.0)


def test_binary_search_slice_smoke():
    # database.rs:569-580
    data = np.array([1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0])
    assert O.binary_search_slice(data, 1.75, 3.5) == (1, 6)
    assert O.binary_search_slice(data, 0.0, 5.0) == (0, len(data))


def test_binary_search_slice_run():
    # database.rs:582-593
    data = np.array([1.0, 1.5, 1.5, 1.5, 1.5, 2.0, 2.5, 3.0, 3.0, 3.5, 4.0])
    l, r = O.binary_search_slice(data, 1.5, 3.25)
    assert data[l] <= 1.5 and data[r] > 3.25
    assert list(data[l:r]) == [1.0, 1.5, 1.5, 1.5, 1.5, 2.0, 2.5, 3.0, 3.0]


def test_max_fragment_charge():
    # scoring.rs:820-830
    assert O.max_fragment_charge(None, 1) == 2
    assert O.max_fragment_charge(None, 2) == 2
    assert O.max_fragment_charge(None, 3) == 3
    assert O.max_fragment_charge(None, 4) == 4
    assert O.max_fragment_charge(1, 2) == 2
    assert O.max_fragment_charge(1, 3) == 2
    assert O.max_fragment_charge(2, 4) == 3
    assert O.max_fragment_charge(4, 1) == 2


def test_longest_series():
    # scoring.rs:799-818
    r = O.run_ladder([1, 2, 3, 3, 3])
    assert r["length"] == 3 and r["longest"] == 3
    r = O.run_ladder([1, 2, 3, 3, 3, 5, 5])
    assert r["length"] == 1 and r["longest"] == 3
    r = O.run_ladder([1, 2, 3, 3, 3, 5, 5, 6])
    assert r["length"] == 2


@pytest.mark.parametrize("seed", range(20))
def test_heap_quickcheck(seed):
    # heap.rs:64-88: top-k set equals a full sort, heap property on the first k
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 400))
    k = int(rng.integers(0, 500))
    data = rng.integers(-1000, 1000, size=n).astype(np.int32)
    k = min(k, n)
    out = O.bounded_min_heapify(data.copy(), k)
    top = out[:k]
    for i in range(1, k):
        assert top[(i - 1) // 2] <= top[i] or k == n
    assert sorted(top.tolist(), reverse=True) == sorted(data.tolist(), reverse=True)[:k]


def test_heap_smoke():
    # heap.rs:94-100
    for arr in (np.arange(500), np.arange(500)[::-1]):
        out = O.bounded_min_heapify(arr.astype(np.int32), 50)
        assert sorted(out[:50].tolist()) == list(range(450, 500))


def _mz(ions, charge):
    return (ions + f32(charge) * PROTON) / f32(charge)


def test_ion_series_abc_xyz():
    # ion_series.rs:157-175 (within 0.005)
    exp = {
        "a": [70.065, 199.108, 296.160, 397.208, 510.292, 625.32],
        "b": [98.0600, 227.1026, 324.155, 425.2030, 538.287, 653.314],
        "c": [115.086, 244.129, 341.182, 442.229, 555.314, 670.341],
        "x": [729.294, 600.251, 503.198, 402.151, 289.066, 174.039],
        "y": [703.314, 574.2719, 477.219, 376.171, 263.0874, 148.0604],
        "z": [686.288, 557.245, 460.193, 359.145, 246.061, 131.034],
    }
    for kind, e in exp.items():
        ions, _ = O.ion_series("PEPTIDE", kind)
        assert len(ions) == 6
        assert np.all(np.abs(_mz(ions, 1) - np.array(e, dtype=np.float32)) < 0.005), kind


def test_ion_series_decoy_charge2():
    # ion_series.rs:253-271
    ions, _ = O.ion_series("PEPTIDE", "y")
    assert np.all(np.abs(_mz(ions, 2) - f32([352.16087, 287.6396, 239.11319, 188.58935, 132.04732, 74.53385])) < 0.005)
    ions, _ = O.ion_series("EDITPEP", "y")
    assert np.all(np.abs(_mz(ions, 2) - f32([336.16596, 278.6525, 222.11046, 171.58662, 123.060237, 58.53894])) < 0.005)


def test_ion_series_mods():
    # ion_series.rs:273-327 (nterm / cterm / internal static mods)
    b0 = f32([98.06004, 227.10263, 324.1554, 425.20306, 538.2872, 653.3141])
    y0 = f32([703.31447, 574.27188, 477.21912, 376.17144, 263.08737, 148.06043])
    b, _ = O.ion_series("PEPTIDE", "b", nterm=229.01)
    y, _ = O.ion_series("PEPTIDE", "y", nterm=229.01)
    assert np.all(np.abs(_mz(b, 1) - (b0 + f32(229.01))) < 0.005) and np.all(np.abs(_mz(y, 1) - y0) < 0.005)
    b, mono = O.ion_series("PEPTIDE", "b", cterm=229.01)
    y, _ = O.ion_series("PEPTIDE", "y", cterm=229.01)
    assert abs(mono - 1028.37) < 0.001
    assert np.all(np.abs(_mz(b, 1) - b0) < 0.005) and np.all(np.abs(_mz(y, 1) - (y0 + f32(229.01))) < 0.005)
    mods = [0, 0, 0, 0, 29.0, 0, 0]
    b, _ = O.ion_series("PEPTIDE", "b", mods=mods)
    y, _ = O.ion_series("PEPTIDE", "y", mods=mods)
    assert np.all(np.abs(_mz(b, 1) - (b0 + f32([0, 0, 0, 0, 29, 29]))) < 0.005)
    assert np.all(np.abs(_mz(y, 1) - (y0 + f32([29, 29, 29, 29, 0, 0]))) < 0.005)


def test_select_most_intense_peak():
    # spectrum.rs:570-605
    masses = [99.0, 100.0, 100.01, 100.02, 101.0]
    intens = [10.0, 20.0, 50.0, 30.0, 100.0]
    assert O.select_most_intense_peak(masses, intens, 100.01, (O.DA, -0.02, 0.02)) == 2
    label = f32(126.127726)
    masses = [label - PROTON - f32(0.01), label - PROTON, label - PROTON + f32(0.01)]
    assert O.select_most_intense_peak(masses, [10.0, 100.0, 50.0], float(label), (O.DA, -0.005, 0.005), offset=float(-PROTON)) == 1
    assert O.select_most_intense_peak([1.0, 2.0], [1.0, 1.0], 50.0, (O.PPM, -10, 10)) is None


def test_deisotope():
    # spectrum.rs:419-567 (exact Deisotoped vectors, before and after path compression)
    mz = np.array([800.9, f32(800.9) + NEUTRON * f32(1.0), f32(800.9) + NEUTRON * f32(2.0), 803.4080, 804.4108, 805.4106, 806.4116, 810.0, 812.0,
                   f32(812.0) + NEUTRON / f32(2.0)], dtype=np.float32)
    inten = f32([2., 1.5, 1., 4., 3., 2., 1., 1., 9.0, 4.5])
    oi, oc, oe = O.deisotope(mz, inten, 2, 5.0, 800.91)
    assert oi.tolist() == [2.0, 2.5, 1.0, 10.0, 6.0, 3.0, 1.0, 1.0, 13.5, 4.5]
    assert oc.tolist() == [-1, 1, 1, 1, 1, 1, 1, -1, 2, 2]
    assert oe.tolist() == [-1, -1, 1, -1, 3, 4, 5, -1, -1, 8]
    oi, oc, oe = O.deisotope(mz, inten, 2, 5.0, 800.91, compress=True)
    assert oi.tolist() == [2.0, 2.5, 0.0, 10.0, 0.0, 0.0, 0.0, 1.0, 13.5, 0.0]
    assert oe.tolist() == [-1, -1, 1, -1, 3, 3, 3, -1, -1, 8]


def test_digestion_order():
    # database.rs:595-671
    fasta = """
        >sp|AAAAA
        MEWKLEQSMREQALLKAQLTQLK
        >sp|BBBBB
        RMEWKLEQSMREQALLKAQLTQLK
        """
    db = O.OracleDB.from_fasta(fasta, bucket_size=128, missed_cleavages=1, min_len=6, max_len=10, peptide_min_mass=150.0,
                               variable_mods={"[": [42.0]}, max_variable_mods=2, generate_decoys=False)
    seqs = [db.peptide_string(i)[0] for i in range(db.n_peptides)]
    assert seqs == ["EQALLK", "LEQSMR", "AQLTQLK", "MEWKLEQSMR", "[+42]-MEWKLEQSMR"]
    for i in range(4):
        assert db.peptide_string(i)[1] == 2
    assert db.peptide_proteins(4) == ["sp|AAAAA"]


Q99536 = """
>sp|Q99536|VAT1_HUMAN Synaptic vesicle membrane protein VAT-1 homolog OS=Homo sapiens OX=9606 GN=VAT1 PE=1 SV=2
MSDEREVAEAATGEDASSPPPKTEAASDPQHPAASEGAAAAAASPPLLRCLVLTGFGGYD
KVKLQSRPAAPPAPGPGQLTLRLRACGLNFADLMARQGLYDRLPPLPVTPGMEGAGVVIA
VGEGVSDRKAGDRVMVLNRSGMWQEEVTVPSVQTFLIPEAMTFEEAAALLVNYITAYMVL
FDFGNLQPGHSVLVHMAAGGVGMAAVQLCRTVENVTVFGTASASKHEALKENGVTHPIDY
HTTDYVDEIKKISPKGVDIVMDPLGGSDTAKGYNLLKPMGKVVTYGMANLLTGPKRNLMA
LARTWWNQFSVTALQLLQANRAVCGFHLGYLDGEVELVSGVVARLLALYNQGHIKPHIDS
VWPFEKVADAMKQMQEKKNVGKVLLVPGPEKEN
"""


@pytest.mark.parametrize("seed", range(40))
def test_check_all_ions_visited(seed):
    # crates/sage/tests/integration.rs:30-70 (quickcheck over target_fragment_mz and bucket_size)
    rng = np.random.default_rng(1000 + seed)
    bucket_size = int(rng.integers(1, 8193)) if seed % 4 else int(2 ** rng.integers(0, 14))
    target = f32(rng.choice([rng.uniform(0, 3000), rng.uniform(-100, 100), rng.uniform(200, 900)]))
    db = O.OracleDB.from_fasta(Q99536, bucket_size=bucket_size, generate_decoys=False)
    e = db.export()
    bs = db.bucket_size
    assert bs & (bs - 1) == 0  # next_power_of_two
    flo, fhi = O.tolerance_bounds(O.DA, -100.0, 100.0, target)
    expected = np.zeros(db.n_peptides, dtype=np.int64)
    for c in range(db.n_buckets):
        pep, mz = e["frag_pep"][c * bs:(c + 1) * bs], e["frag_mz"][c * bs:(c + 1) * bs]
        assert np.all(np.diff(pep.astype(np.int64)) >= 0)
        assert np.all(mz >= e["bucket_min"][c])
        if c + 1 < db.n_buckets:
            assert np.all(mz <= e["bucket_min"][c + 1])
        sel = (mz >= f32(flo)) & (mz <= f32(fhi))
        np.add.at(expected, pep[sel], 1)
    pep, mz, _ = db.page_search(1000.0, (O.DA, -5000.0, 5000.0), (O.DA, -100.0, 100.0), float(target))
    visited = np.bincount(pep, minlength=db.n_peptides)
    assert np.array_equal(expected, visited)


def _process(config1, top_n):
    return O.process_ms2(config1["mz"], config1["intensity"], config1["precursor_charge"], top_n, True, 0.0)


def _one_spectrum(config1, masses, intens, tic):
    return dict(peak_off=np.array([0, len(masses)], np.uint64), masses=masses, intensities=intens,
                prec_mz=f32([config1["precursor_mz"]]), prec_charge=np.array([config1["precursor_charge"]], np.uint8),
                iso_lo=f32([config1["isolation_window_da"][0]]), iso_hi=f32([config1["isolation_window_da"][1]]), tic=f32([tic]))


def test_integration_matched_peaks_21(config1):
    # crates/sage-cli/tests/integration.rs:7-52 — the reference's only end-to-end known answer
    db = O.OracleDB.from_fasta(config1["fasta"])  # Builder::default(): trypsin, 0 missed, decoys, bucket 8192
    masses, intens, tic = _process(config1, 100)
    assert len(masses) <= 300
    cfg = O.ScorerConfig(precursor_tol=(O.PPM, -50.0, 50.0), fragment_tol=(O.PPM, -10.0, 10.0), min_matched_peaks=4, min_isotope_err=-1,
                         max_isotope_err=3, min_precursor_charge=2, max_precursor_charge=4, max_fragment_charge=1, report_psms=1)
    feats, counts, _, _ = db.score_batch(cfg, _one_spectrum(config1, masses, intens, tic))
    assert counts[0] == 1
    assert feats[0]["matched_peaks"] == 21
    assert db.sequence(int(feats[0]["peptide_idx"])) == "LQSRPAAPPAPGPGQLTLR"
    assert feats[0]["rank"] == 1 and feats[0]["label"] == 1 and feats[0]["charge"] == 3


def test_config1_tests_config_json(config1):
    # tests/config.json run by CI (.github/workflows/rust.yml:22-34); values derived in SURVEY.md §4/§8d
    db = O.OracleDB.from_fasta(config1["fasta"], bucket_size=16384, missed_cleavages=1, static_mods={"C": 57.0216})
    assert db.n_peptides == 102 and db.n_fragments == 2872
    masses, intens, tic = _process(config1, 150)
    assert len(masses) == 150
    cfg = O.ScorerConfig(precursor_tol=(O.PPM, -50.0, 50.0), fragment_tol=(O.PPM, -10.0, 10.0), min_isotope_err=-1, max_isotope_err=3,
                         max_fragment_charge=1, report_psms=1)
    feats, counts, _, _ = db.score_batch(cfg, _one_spectrum(config1, masses, intens, tic))
    assert counts[0] == 1
    assert feats[0]["matched_peaks"] == 22
    assert db.sequence(int(feats[0]["peptide_idx"])) == "LQSRPAAPPAPGPGQLTLR"
    assert abs(feats[0]["hyperscore"] - 72.26591574) < 1e-6


def test_hyperscore_against_high_precision_arithmetic():
    """The f64 scores pinned by something other than the oracle: Score::hyperscore (scoring.rs:179-201) re-evaluated with 60-digit arithmetic
    (mpmath) from the same f32 inputs. The oracle evaluates ln through the host libm (< 1 ulp) and sums four f64 terms, so it must agree with the
    exactly rounded value to a few ulp of the result — and the ORDER of two candidates whose exact scores differ by more than that must be the
    exact order (the rank is a stable sort on this value, scoring.rs:495)."""
    mpmath = pytest.importorskip("mpmath")
    import ctypes
    import ctypes.util
    log1pf = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6").log1pf
    log1pf.restype, log1pf.argtypes = ctypes.c_float, [ctypes.c_float]
    mp = mpmath.mp
    mp.prec = 200
    rng = np.random.default_rng(0x5C0E)

    def exact_lnfact(n):   # scoring.rs:170-177 (Stirling-style approximation; n == 0 -> 1.0), every f64 operation replaced by an exact one
        if n == 0:
            return mpmath.mpf(1)
        x = mpmath.mpf(n)
        return x * mpmath.log(x) - x + mpmath.mpf(0.5) * mpmath.log(x) + mpmath.mpf(0.5) * mpmath.log(mpmath.pi * 2 * x)

    def exact_score(score_type, mb, my, sb, sy):
        sb32, sy32 = np.float32(sb), np.float32(sy)
        if score_type == 0:
            i = mpmath.mpf(float(np.float32(sb32 + np.float32(1.0)))) * mpmath.mpf(float(np.float32(sy32 + np.float32(1.0))))   # the f64 product of two f32 is exact
            return mpmath.log(i) + exact_lnfact(mb) + exact_lnfact(my)
        si = float(np.float32(sb32 + sy32))
        return mpmath.mpf(float(log1pf(ctypes.c_float(si)))) + exact_lnfact(mb) + exact_lnfact(my)   # ln_1p is evaluated in f32 by the reference (libm log1pf)

    cases = []
    for _ in range(4000):
        mb, my = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        sb, sy = float(np.float32(rng.uniform(0, 5e6))), float(np.float32(rng.uniform(0, 5e6)))
        cases.append((mb, my, sb, sy))
    worst = 0.0
    scored = []
    for mb, my, sb, sy in cases:
        got = O.hyperscore(0, mb, my, sb, sy)
        want = exact_score(0, mb, my, sb, sy)
        ulp = float(np.spacing(np.float64(abs(float(want)))))
        err = abs(float(mpmath.mpf(got) - want)) / ulp
        worst = max(worst, err)
        scored.append((got, want))
    # ln < 1 ulp of its own result (<= 1 ulp of the sum, which is larger), each lnfact carries its own libm roundings (4 of them at most 1 ulp of
    # terms no larger than the sum) and three f64 additions add 0.5 ulp each
    assert worst <= 8.0, worst
    # order: pairs whose exact scores are further apart than that bound must sort the same way
    idx = rng.integers(0, len(scored), (20000, 2))
    for a, b in idx:
        ga, wa = scored[a]
        gb, wb = scored[b]
        gap = abs(float(wa - wb))
        if gap > 16 * float(np.spacing(np.float64(max(abs(float(wa)), abs(float(wb)))))):
            assert (ga < gb) == (wa < wb)
    # lnfact itself (its large terms cancel, so the bound is relative to the largest term, not to the result)
    for n in (1, 2, 3, 10, 59, 255, 1000):
        want = exact_lnfact(n)
        big = float(n) * float(np.log(max(n, 2)))
        assert abs(O.lnfact(n) - float(want)) <= 8 * np.spacing(np.float64(max(big, 1.0))), n
    # the OpenMS-style score goes through f32 ln_1p: checked bit for bit against libm elsewhere (tests/test_glibc_log.py); here only that the
    # remaining f64 arithmetic agrees
    for mb, my, sb, sy in cases[:500]:
        got = O.hyperscore(1, mb, my, sb, sy)
        want = exact_score(1, mb, my, sb, sy)
        assert abs(got - float(want)) <= 8 * np.spacing(np.float64(abs(float(want)))), (mb, my, sb, sy)
