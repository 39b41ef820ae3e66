"""BASELINE.json-size checks (cfg2 / cfg4 / cfg5: ~1.9 M peptides, 49 M fragments; cfg3: ~16 M peptides, ~680 M fragments): oracle parity on samples the CPU finishes in seconds,
and size-independent properties over the full 50 k-spectrum batch."""
import numpy as np
import pytest

from sage_b200 import IndexedDatabase, Scorer, Tolerance, synth

from helpers import assert_features_equal, oracle_cfg, oracle_db_from_peptides, valid_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    pep = synth.make_peptides(2_000_000)
    spectra = synth.make_spectra(pep, 50_000, seed=0xB202)
    gdb = IndexedDatabase.build_from_peptides(pep)
    odb = oracle_db_from_peptides(pep)
    return pep, spectra, gdb, odb


def test_index_matches_oracle_at_scale(full):
    pep, _, gdb, odb = full
    fp, fm, bm = gdb.export_index()
    e = odb.export()
    assert np.array_equal(bm.view(np.uint32), e["bucket_min"].view(np.uint32))
    assert np.array_equal(fp, e["frag_pep"]) and np.array_equal(fm.view(np.uint32), e["frag_mz"].view(np.uint32))
    # index invariants (crates/sage/tests/integration.rs:39-52): PeptideIx ascending inside a bucket, bucket minima ascending
    bs = gdb.info["bucket_size"]
    d = np.diff(fp.astype(np.int64))
    assert np.all(d[np.arange(len(d)) % bs != bs - 1] >= 0) and np.all(np.diff(bm) >= 0)


def test_cfg2_sample_parity_and_properties(full):
    pep, spectra, gdb, odb = full
    kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20), report_psms=2)
    sc = Scorer(gdb, **kw)
    gf, gc = sc.score_batch(spectra)
    # oracle parity on a 3000-spectrum sample spread over the batch
    for a in (0, 23_000, 47_000):
        sub = spectra.slice(a, a + 1000)
        of, oc, _, _ = odb.score_batch(oracle_cfg(**kw), sub.as_dict())
        of = of.copy()
        of["spectrum"] += np.uint32(a)
        assert_features_equal(gf[2 * a:2 * (a + 1000)], gc[a:a + 1000], of, oc, 2, what=f"cfg2 sample @{a}")
    # properties over all 50k spectra
    sel = (np.arange(len(gf)) % 2) < np.repeat(gc, 2)
    g = gf[sel]
    assert gc.sum() > 45_000
    assert np.all(g["matched_peaks"] >= 4) and np.all(g["rank"] >= 1) and np.all(g["rank"] <= 2)
    two = gc == 2
    assert np.all(gf[0::2]["hyperscore"][two] >= gf[1::2]["hyperscore"][two])          # sorted by hyperscore
    assert np.allclose(gf[0::2]["delta_next"][two], gf[0::2]["hyperscore"][two] - gf[1::2]["hyperscore"][two], rtol=0, atol=1e-9)
    assert np.abs(g["delta_mass"]).max() <= 20.2, np.abs(g["delta_mass"]).max()         # inside the +-20 ppm precursor tolerance (f32 bounds)
    # sharding invariance: halves scored separately give the same rows (spectrum index rebased)
    h0, c0 = sc.score_batch(spectra.slice(0, 25_000))
    h1, c1 = sc.score_batch(spectra.slice(25_000, 50_000))
    h1 = h1.copy()
    h1["spectrum"] += np.uint32(25_000)
    both, cb = np.concatenate([h0, h1]), np.concatenate([c0, c1])
    assert np.array_equal(cb, gc) and both[sel].tobytes() == gf[sel].tobytes()
    # chunking invariance: the same batch cut into 3 pipelined chunks (two lanes alternate) gives the same bytes
    sc.set_option("pipeline_chunks", 3)
    g3, c3 = sc.score_batch(spectra)
    assert np.array_equal(c3, gc) and g3[sel].tobytes() == gf[sel].tobytes()


def test_host_paths_agree_bitwise(full):
    """The same 50k batch through every host path of sage_b200_score_batch gives the same bytes: pageable caller arrays (staged through the lane's
    pinned buffers by the pool + helper thread), pinned caller arrays with the peak-mass copy in 1, 2 and 4 parts (one counting launch per part),
    and two pipelined chunks."""
    from sage_b200 import SpectraBatch, api
    pep, spectra, gdb, _ = full
    kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20))
    sc = Scorer(gdb, **kw)
    ref_f, ref_c = sc.score_batch(spectra)      # pageable numpy arrays
    ref_f, ref_c = valid_rows(ref_f, ref_c, 1).copy(), ref_c.copy()
    assert ref_c.sum() > 45_000
    pm, pi = api.pinned_empty(spectra.masses.shape, np.float32), api.pinned_empty(spectra.intensities.shape, np.float32)
    pm[...] = spectra.masses
    pi[...] = spectra.intensities
    pinned = SpectraBatch(**{**spectra.__dict__, "masses": pm, "intensities": pi})
    try:
        for parts in (1, 2, 4):
            sc.set_option("mass_parts", parts)
            f, c = sc.score_batch(pinned)
            assert np.array_equal(c, ref_c) and valid_rows(f, c, 1).tobytes() == ref_f.tobytes(), f"pinned, {parts} part(s)"
        sc.set_option("pipeline_chunks", 2)
        for batch, what in ((pinned, "pinned"), (spectra, "pageable")):
            f, c = sc.score_batch(batch)
            assert np.array_equal(c, ref_c) and valid_rows(f, c, 1).tobytes() == ref_f.tobytes(), f"{what}, two chunks"
    finally:
        api.pinned_free(pm)
        api.pinned_free(pi)


def test_cfg4_sample_parity(full):
    pep, spectra, gdb, odb = full
    kw = dict(precursor_tol=Tolerance.da(-500, 500), fragment_tol=Tolerance.ppm(-20, 20))
    sub = spectra.slice(1000, 1096)
    sc = Scorer(gdb, **kw)
    gf, gc = sc.score_batch(sub)
    of, oc, _, octr = odb.score_batch(oracle_cfg(**kw), sub.as_dict(), counters=True)
    assert_features_equal(gf, gc, of, oc, 1, what="cfg4 sample")
    c = sc.counters()
    assert c["wide_queries"] == 96 and c["entries_scanned"] == octr["entries_scanned"] and c["pages"] == octr["pages"]


def test_cfg5_at_size_chimeric_sample_parity(full):
    """BASELINE.json configs[4] at its stated size: 100k co-fragmenting spectra, chimera = true, report_psms = 5, against the ~2M-peptide index.
    Oracle parity on 3 x 1000-spectrum samples; properties over all 100k (rank == round, at most 5 PSMs, peaks are consumed)."""
    pep, _, gdb, odb = full
    chim = synth.make_spectra(pep, 100_000, seed=0xB205, chimeric=True)
    kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20), chimera=True, report_psms=5)
    sc = Scorer(gdb, **kw)
    gf, gc = sc.score_batch(chim)
    total = 0
    for a in (0, 48_000, 99_000):
        sub = chim.slice(a, a + 1000)
        of, oc, _, _ = odb.score_batch(oracle_cfg(**kw), sub.as_dict())
        of = of.copy()
        of["spectrum"] += np.uint32(a)
        total += assert_features_equal(gf[5 * a:5 * (a + 1000)], gc[a:a + 1000], of, oc, 5, what=f"cfg5 sample @{a}")
    assert total > 3500
    sel = (np.arange(len(gf)) % 5) < np.repeat(gc, 5)
    g = gf[sel]
    assert gc.max() <= 5 and gc.sum() > 120_000
    assert np.array_equal(g["rank"], (np.arange(len(gf)) % 5)[sel] + 1)          # chimera: rank == round (scoring.rs:662)
    assert np.all(g["matched_peaks"] >= 4)


@pytest.fixture(scope="module")
def cfg3():
    """BASELINE.json configs[2]: human-tryptic-scale digest + 2 variable modifications (M+15.9949, STY+79.9663, up to 2 per peptide) + static C:
    ~16 M peptides, ~680 M fragments (the same table bench.py's cfg3 uses)."""
    pep = synth.make_peptides(2_000_000, var_mods=(("M", 15.9949), ("STY", 79.9663)), max_variable_mods=2, static_c=True)
    gdb = IndexedDatabase.build_from_peptides(pep)
    odb = oracle_db_from_peptides(pep)
    return pep, gdb, odb


def test_cfg3_index_matches_oracle_at_size(cfg3):
    pep, gdb, odb = cfg3
    assert len(pep) > 14_000_000 and gdb.info["n_fragments"] > 400_000_000
    fp, fm, bm = gdb.export_index()
    e = odb.export()
    assert np.array_equal(bm.view(np.uint32), e["bucket_min"].view(np.uint32))
    step = 1 << 26   # compare in slices: the arrays are 2.7 GB each
    for a in range(0, len(fp), step):
        assert np.array_equal(fp[a:a + step], e["frag_pep"][a:a + step]), a
        assert np.array_equal(fm[a:a + step].view(np.uint32), e["frag_mz"][a:a + step].view(np.uint32)), a


def test_cfg3_at_size_sample_parity(cfg3):
    """200k spectra sharded over 8 GPUs = 25k per GPU: one shard here, oracle parity on 3 x 1000 spectra of it, plus an isotope-error /
    report_psms variant on a smaller sample (the variable-mod index puts many isobaric positional isomers into every precursor window)."""
    pep, gdb, odb = cfg3
    spectra = synth.make_spectra(pep, 25_000, seed=0xB203)
    kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20))
    sc = Scorer(gdb, **kw)
    gf, gc = sc.score_batch(spectra)
    total = 0
    for a in (0, 12_000, 24_000):
        sub = spectra.slice(a, a + 1000)
        of, oc, _, _ = odb.score_batch(oracle_cfg(**kw), sub.as_dict())
        of = of.copy()
        of["spectrum"] += np.uint32(a)
        total += assert_features_equal(gf[a:a + 1000], gc[a:a + 1000], of, oc, 1, what=f"cfg3 sample @{a}")
    assert total > 2500
    kw2 = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20), report_psms=5, min_isotope_err=-1, max_isotope_err=2)
    sub = spectra.slice(5000, 5400)
    g2, c2 = Scorer(gdb, **kw2).score_batch(sub)
    o2, oc2, _, _ = odb.score_batch(oracle_cfg(**kw2), sub.as_dict())
    assert assert_features_equal(g2, c2, o2, oc2, 5, what="cfg3 isotope errors, report_psms 5") > 800
