"""Shared helpers for the parity tests: build the oracle DB and the device DB from the same peptide table, and compare
Feature tables field by field. Integer and f32 fields are always compared bit for bit. The f64 scores (hyperscore, delta_next,
delta_best, poisson) are compared BIT FOR BIT too whenever the library reports that it reproduces this host's libm log()
(sage_b200_host_log_variant() >= 0, and sage_b200_host_log1pf_exact() for the OpenMS score type: the kernels evaluate glibc's algorithms,
glibc_log.cuh); only hosts with an unknown libm fall back to north_star's 1e-6 relative tolerance."""
import numpy as np

from oracle import oracle as O

EXACT_FIELDS = ["peptide_idx", "peptide_len", "rank", "label", "expmass", "calcmass", "charge", "delta_mass", "isotope_error", "average_ppm",
                "matched_peaks", "longest_b", "longest_y", "longest_y_pct", "missed_cleavages", "matched_intensity_pct", "scored_candidates",
                "ms2_intensity"]
F64_FIELDS = ["hyperscore", "delta_next", "delta_best", "poisson"]
F64_RTOL = 1e-6   # north_star: within 1e-6 relative on hyperscore
F64_ATOL = 1e-9   # differences of two hyperscores (delta_*) inherit an absolute error of ~1 ulp(hyperscore)


def oracle_db_from_peptides(pep, bucket_size=8192, ion_kinds=("b", "y"), min_ion_index=2):
    return O.OracleDB.from_peptides(pep.seq_off, pep.seq, pep.mods, pep.nterm, pep.mono, pep.decoy, pep.missed, bucket_size=bucket_size,
                                    ion_kinds=ion_kinds, min_ion_index=min_ion_index)


def peptides_from_oracle(odb):
    from sage_b200 import Peptides
    e = odb.export()
    return Peptides(seq_off=e["seq_off"], seq=e["seq"], mods=e["mods"], nterm=e["nterm"], mono=e["pep_mono"], decoy=e["decoy"], missed=e["missed"])


def oracle_cfg(**kw):
    """ScorerConfig from Scorer-style kwargs with sage_b200.Tolerance values."""
    kw = dict(kw)
    for k in ("precursor_tol", "fragment_tol"):
        if k in kw and hasattr(kw[k], "as_tuple"):
            kw[k] = kw[k].as_tuple()
    return O.ScorerConfig(**kw)


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def f64_exact_default(score_type=0):
    from sage_b200 import api
    return api.host_log_variant() >= 0 and (score_type == 0 or api.host_log1pf_exact())


def assert_features_equal(gf, gc, of, oc, report_psms, what="", f64_exact=None):
    if f64_exact is None:
        f64_exact = f64_exact_default()
    assert np.array_equal(gc, oc), f"{what}: PSM counts differ at spectra {np.nonzero(gc != oc)[0][:10]}"
    n = len(gc)
    sel = (np.arange(n * report_psms) % report_psms) < np.repeat(gc, report_psms)
    g, o = gf[sel], of[sel]
    assert np.array_equal(g["spectrum"], o["spectrum"])
    for f in EXACT_FIELDS:
        a, b = g[f], o[f]
        if a.dtype.kind == "f":
            same = (bits(a.astype(np.float32)) == bits(b.astype(np.float32))) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        if not np.all(same):
            bad = np.nonzero(~same)[0][:5]
            raise AssertionError(f"{what}: field {f} differs at rows {bad}: gpu={a[bad]} oracle={b[bad]} (spectra {g['spectrum'][bad]})")
    for f in F64_FIELDS:
        a, b = g[f], o[f]
        if f64_exact:
            ok = (np.ascontiguousarray(a).view(np.uint64) == np.ascontiguousarray(b).view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        else:
            ok = np.isclose(a, b, rtol=F64_RTOL, atol=F64_ATOL) | (a == b) | (np.isnan(a) & np.isnan(b))
        if not np.all(ok):
            bad = np.nonzero(~ok)[0][:5]
            raise AssertionError(f"{what}: field {f} differs at rows {bad} ({'bit-exact' if f64_exact else 'rtol 1e-6'} compare, {int((~ok).sum())} rows): "
                                 f"gpu={[float(x).hex() for x in a[bad]]} oracle={[float(x).hex() for x in b[bad]]}")
    return int(sel.sum())


def valid_rows(features, counts, report_psms):
    """The rows of a score_batch result that hold PSMs (spectrum i: rows i*report_psms .. +counts[i]); the rest of the array is never written."""
    import numpy as np
    k = np.arange(len(features)) % report_psms
    return features[k < np.repeat(np.asarray(counts), report_psms)]
