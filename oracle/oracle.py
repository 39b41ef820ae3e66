"""ctypes binding of the CPU oracle (oracle/sage_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Never imported by the sage_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsage_oracle.so")

PPM, PCT, DA = 0, 1, 2
KIND = {"a": 0, "b": 1, "c": 2, "x": 3, "y": 4, "z": 5}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sage_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Tol(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lo", C.c_float), ("hi", C.c_float)]


class ScorerParams(C.Structure):
    _fields_ = [
        ("precursor_tol", Tol), ("fragment_tol", Tol),
        ("min_matched_peaks", C.c_uint16), ("min_isotope_err", C.c_int8), ("max_isotope_err", C.c_int8),
        ("min_precursor_charge", C.c_uint8), ("max_precursor_charge", C.c_uint8), ("override_precursor_charge", C.c_uint8),
        ("max_fragment_charge", C.c_int8),
        ("chimera", C.c_uint8), ("wide_window", C.c_uint8), ("annotate_matches", C.c_uint8), ("score_type", C.c_uint8),
        ("report_psms", C.c_uint32),
    ]


FEATURE_DTYPE = np.dtype([
    ("spectrum", "<u4"), ("peptide_idx", "<u4"), ("peptide_len", "<u4"), ("rank", "<u4"), ("label", "<i4"),
    ("expmass", "<f4"), ("calcmass", "<f4"), ("charge", "<u4"), ("delta_mass", "<f4"), ("isotope_error", "<f4"), ("average_ppm", "<f4"),
    ("_pad0", "<u4"),
    ("hyperscore", "<f8"), ("delta_next", "<f8"), ("delta_best", "<f8"),
    ("matched_peaks", "<u4"), ("longest_b", "<u4"), ("longest_y", "<u4"), ("longest_y_pct", "<f4"), ("missed_cleavages", "<u4"),
    ("matched_intensity_pct", "<f4"), ("scored_candidates", "<u4"), ("_pad1", "<u4"), ("poisson", "<f8"), ("ms2_intensity", "<f4"),
    ("frag_offset", "<u4"), ("frag_count", "<u4"), ("_pad2", "<u4"),
])
FRAGMENT_DTYPE = np.dtype([("kind", "<i4"), ("charge", "<i4"), ("ordinal", "<i4"), ("intensity", "<f4"),
                           ("mz_calculated", "<f4"), ("mz_experimental", "<f4")])
COUNTER_FIELDS = ["queries", "probes_pep", "probes_bucket", "pages", "probes_page", "entries_scanned", "candidates_scored", "psms",
                  "peptide_record_floats"]


class BuildParams(C.Structure):
    _fields_ = [
        ("bucket_size", C.c_uint64), ("missed_cleavages", C.c_uint8), ("min_len", C.c_uint64), ("max_len", C.c_uint64),
        ("cleave_at", C.c_char_p), ("restrict_", C.c_char_p), ("c_terminal", C.c_uint8), ("semi_enzymatic", C.c_uint8),
        ("peptide_min_mass", C.c_float), ("peptide_max_mass", C.c_float),
        ("ion_kinds", C.POINTER(C.c_uint8)), ("n_kinds", C.c_uint64), ("min_ion_index", C.c_uint64),
        ("static_mod_specs", C.POINTER(C.c_char_p)), ("static_mod_masses", C.POINTER(C.c_float)), ("n_static", C.c_uint64),
        ("var_mod_specs", C.POINTER(C.c_char_p)), ("var_mod_masses", C.POINTER(C.c_float)), ("n_var", C.c_uint64),
        ("max_variable_mods", C.c_uint64), ("decoy_tag", C.c_char_p), ("generate_decoys", C.c_uint8),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        assert _lib.so_db_n_peptides  # symbol check
        _lib.so_db_from_fasta.restype = C.c_void_p
        _lib.so_db_from_peptides.restype = C.c_void_p
        for f in ("so_db_n_peptides", "so_db_n_fragments", "so_db_n_buckets", "so_db_bucket_size", "so_db_total_residues",
                  "so_ion_series", "so_process_ms2", "so_page_search"):
            getattr(_lib, f).restype = C.c_uint64
        _lib.so_score_batch.restype = C.c_int64
        _lib.so_quick_score.restype = C.c_int64
        _lib.so_initial_hits.restype = C.c_int64
        _lib.so_max_fragment_charge.restype = C.c_uint8
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


@dataclass
class ScorerConfig:
    """Mirror of sage-core's Scorer public fields (scoring.rs:210-232); tolerances are (kind, lo, hi)."""
    precursor_tol: tuple = (PPM, -50.0, 50.0)
    fragment_tol: tuple = (PPM, -10.0, 10.0)
    min_matched_peaks: int = 4
    min_isotope_err: int = 0
    max_isotope_err: int = 0
    min_precursor_charge: int = 2
    max_precursor_charge: int = 4
    override_precursor_charge: bool = False
    max_fragment_charge: int | None = None
    chimera: bool = False
    report_psms: int = 1
    wide_window: bool = False
    annotate_matches: bool = False
    score_type: int = 0

    def to_c(self) -> ScorerParams:
        p = ScorerParams()
        p.precursor_tol = Tol(*self.precursor_tol)
        p.fragment_tol = Tol(*self.fragment_tol)
        p.min_matched_peaks = self.min_matched_peaks
        p.min_isotope_err = self.min_isotope_err
        p.max_isotope_err = self.max_isotope_err
        p.min_precursor_charge = self.min_precursor_charge
        p.max_precursor_charge = self.max_precursor_charge
        p.override_precursor_charge = int(self.override_precursor_charge)
        p.max_fragment_charge = -1 if self.max_fragment_charge is None else self.max_fragment_charge
        p.chimera = int(self.chimera)
        p.wide_window = int(self.wide_window)
        p.annotate_matches = int(self.annotate_matches)
        p.score_type = self.score_type
        p.report_psms = self.report_psms
        return p


class OracleDB:
    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        L = lib()
        self.n_peptides = L.so_db_n_peptides(self.h)
        self.n_fragments = L.so_db_n_fragments(self.h)
        self.n_buckets = L.so_db_n_buckets(self.h)
        self.bucket_size = L.so_db_bucket_size(self.h)
        self._export = None

    def __del__(self):
        try:
            lib().so_db_free(self.h)
        except Exception:
            pass

    @staticmethod
    def from_fasta(fasta_text: str, *, bucket_size=8192, missed_cleavages=0, min_len=5, max_len=50, cleave_at="KR", restrict="P",
                   c_terminal=True, semi_enzymatic=False, peptide_min_mass=500.0, peptide_max_mass=5000.0, ion_kinds=("b", "y"),
                   min_ion_index=2, static_mods=None, variable_mods=None, max_variable_mods=2, decoy_tag="rev_", generate_decoys=True):
        """Builder::default()/make_parameters() + Parameters::build (database.rs:29-41,96-115,260)."""
        bp = BuildParams()
        bp.bucket_size = bucket_size
        bp.missed_cleavages = missed_cleavages
        bp.min_len, bp.max_len = min_len, max_len
        bp.cleave_at, bp.restrict_ = cleave_at.encode(), restrict.encode()
        bp.c_terminal, bp.semi_enzymatic = int(c_terminal), int(semi_enzymatic)
        bp.peptide_min_mass, bp.peptide_max_mass = peptide_min_mass, peptide_max_mass
        kinds = (C.c_uint8 * len(ion_kinds))(*[KIND[k] for k in ion_kinds])
        bp.ion_kinds, bp.n_kinds = kinds, len(ion_kinds)
        bp.min_ion_index = min_ion_index
        sm = list((static_mods or {}).items())
        sspec = (C.c_char_p * max(1, len(sm)))(*[k.encode() for k, _ in sm])
        smass = (C.c_float * max(1, len(sm)))(*[v for _, v in sm])
        bp.static_mod_specs, bp.static_mod_masses, bp.n_static = sspec, smass, len(sm)
        vm = [(k, m) for k, ms in (variable_mods or {}).items() for m in ms]
        vspec = (C.c_char_p * max(1, len(vm)))(*[k.encode() for k, _ in vm])
        vmass = (C.c_float * max(1, len(vm)))(*[v for _, v in vm])
        bp.var_mod_specs, bp.var_mod_masses, bp.n_var = vspec, vmass, len(vm)
        bp.max_variable_mods = max_variable_mods
        bp.decoy_tag = decoy_tag.encode()
        bp.generate_decoys = int(generate_decoys)
        return OracleDB(lib().so_db_from_fasta(fasta_text.encode(), C.byref(bp)))

    @staticmethod
    def from_peptides(seq_off, seq, mods, nterm, mono, decoy, missed, *, bucket_size=8192, ion_kinds=("b", "y"), min_ion_index=2):
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint32)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        mods, nterm, mono = _f32(mods), _f32(nterm), _f32(mono)
        decoy = np.ascontiguousarray(decoy, dtype=np.uint8)
        missed = np.ascontiguousarray(missed, dtype=np.uint8)
        kinds = np.array([KIND[k] if isinstance(k, str) else int(k) for k in ion_kinds], dtype=np.uint8)
        h = lib().so_db_from_peptides(C.c_uint64(len(mono)), _p(seq_off), _p(seq), _p(mods), _p(nterm), _p(mono), _p(decoy), _p(missed),
                                      C.c_uint64(bucket_size), _p(kinds), C.c_uint64(len(kinds)), C.c_uint64(min_ion_index))
        return OracleDB(h)

    def export(self) -> dict:
        """Arrays in the reference layout (IndexedDatabase fields, database.rs:384-395)."""
        if self._export is None:
            L = lib()
            nres = L.so_db_total_residues(self.h)
            e = dict(
                frag_pep=np.empty(self.n_fragments, np.uint32), frag_mz=np.empty(self.n_fragments, np.float32),
                bucket_min=np.empty(self.n_buckets, np.float32), pep_mono=np.empty(self.n_peptides, np.float32),
                seq_off=np.empty(self.n_peptides + 1, np.uint32), seq=np.empty(nres, np.uint8), mods=np.empty(nres, np.float32),
                nterm=np.empty(self.n_peptides, np.float32), cterm=np.empty(self.n_peptides, np.float32),
                decoy=np.empty(self.n_peptides, np.uint8), missed=np.empty(self.n_peptides, np.uint8),
            )
            L.so_db_export(self.h, _p(e["frag_pep"]), _p(e["frag_mz"]), _p(e["bucket_min"]), _p(e["pep_mono"]), _p(e["seq_off"]), _p(e["seq"]),
                           _p(e["mods"]), _p(e["nterm"]), _p(e["cterm"]), _p(e["decoy"]), _p(e["missed"]))
            e["bucket_size"] = self.bucket_size
            self._export = e
        return self._export

    def peptide_string(self, i: int):
        buf = C.create_string_buffer(512)
        n = lib().so_db_peptide_string(self.h, C.c_uint64(i), buf, 512)
        return buf.value.decode(), n

    def peptide_proteins(self, i: int):
        out, j = [], 0
        buf = C.create_string_buffer(512)
        while lib().so_db_peptide_protein(self.h, C.c_uint64(i), C.c_uint64(j), buf, 512) == 0:
            out.append(buf.value.decode())
            j += 1
        return out

    def sequence(self, i: int) -> str:
        e = self.export()
        return bytes(e["seq"][e["seq_off"][i]:e["seq_off"][i + 1]]).decode()

    def page_search(self, precursor_mass, ptol, ftol, mass):
        cap = max(1, self.n_fragments)
        pep, mz = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        lohi = np.zeros(2, np.uint64)
        n = lib().so_page_search(self.h, C.c_float(precursor_mass), Tol(*ptol), Tol(*ftol), C.c_float(mass), _p(pep), _p(mz), C.c_uint64(cap), _p(lohi))
        return pep[:n].copy(), mz[:n].copy(), (int(lohi[0]), int(lohi[1]))

    def score_batch(self, cfg: ScorerConfig, spectra: dict, nthreads: int = 0, counters: bool = False):
        """spectra: dict with peak_off(u64,n+1) masses intensities prec_mz prec_charge iso_lo iso_hi tic [level] [ims].
        Returns (features[n*report_psms] structured array, counts[n], fragments|None, counters|None)."""
        n = len(spectra["prec_mz"])
        peak_off = np.ascontiguousarray(spectra["peak_off"], dtype=np.uint64)
        masses, intens = _f32(spectra["masses"]), _f32(spectra["intensities"])
        prec_mz, tic = _f32(spectra["prec_mz"]), _f32(spectra["tic"])
        prec_charge = np.ascontiguousarray(spectra["prec_charge"], dtype=np.uint8)
        iso_lo, iso_hi = _f32(spectra["iso_lo"]), _f32(spectra["iso_hi"])
        level = None if spectra.get("level") is None else np.ascontiguousarray(spectra["level"], dtype=np.uint8)
        ims = None if spectra.get("ims") is None else _f32(spectra["ims"])
        out = np.zeros(n * cfg.report_psms, FEATURE_DTYPE)
        counts = np.zeros(n, np.uint32)
        sp = cfg.to_c()
        frag, frag_used = None, C.c_uint64(0)
        cap = 0
        if cfg.annotate_matches:
            cap = int(n * cfg.report_psms * 600)
            frag = np.zeros(cap, FRAGMENT_DTYPE)
        ctr = (C.c_uint64 * len(COUNTER_FIELDS))()
        rc = lib().so_score_batch(self.h, C.byref(sp), C.c_uint64(n), _p(peak_off), _p(masses), _p(intens), _p(prec_mz), _p(prec_charge), _p(iso_lo),
                                  _p(iso_hi), _p(tic), _p(level), _p(ims), C.c_int(nthreads), _p(out), _p(counts), _p(frag), C.c_uint64(cap),
                                  C.byref(frag_used), ctr if counters else None)
        if rc != 0:
            raise RuntimeError(f"oracle: reference would panic (code {rc})")
        if frag is not None:
            frag = frag[:frag_used.value]
        return out, counts, frag, (dict(zip(COUNTER_FIELDS, list(ctr))) if counters else None)

    def quick_score(self, cfg: ScorerConfig, spectra: dict, low_memory: bool) -> np.ndarray:
        """Scorer::quick_score (scoring.rs:255-298) over a batch -> keep[n_peptides] (uint8)."""
        n = len(spectra["prec_mz"])
        keep = np.zeros(self.n_peptides, np.uint8)
        peak_off = np.ascontiguousarray(spectra["peak_off"], dtype=np.uint64)
        masses, intens = _f32(spectra["masses"]), _f32(spectra["intensities"])
        prec_mz, tic = _f32(spectra["prec_mz"]), _f32(spectra["tic"])
        prec_charge = np.ascontiguousarray(spectra["prec_charge"], dtype=np.uint8)
        iso_lo, iso_hi = _f32(spectra["iso_lo"]), _f32(spectra["iso_hi"])
        sp = cfg.to_c()
        rc = lib().so_quick_score(self.h, C.byref(sp), C.c_uint64(n), _p(peak_off), _p(masses), _p(intens), _p(prec_mz), _p(prec_charge), _p(iso_lo),
                                  _p(iso_hi), _p(tic), C.c_int(int(low_memory)), _p(keep))
        if rc != 0:
            raise RuntimeError(f"oracle: reference would panic (code {rc})")
        return keep

    def initial_hits(self, cfg: ScorerConfig, masses, intens, prec_mz, prec_charge=0, iso_lo=np.nan, iso_hi=np.nan):
        cap = 1 << 22
        m, p = np.zeros(cap, np.uint16), np.zeros(cap, np.uint32)
        c, i = np.zeros(cap, np.uint8), np.zeros(cap, np.int8)
        mp, scd = C.c_uint64(0), C.c_uint64(0)
        masses, intens = _f32(masses), _f32(intens)
        sp = cfg.to_c()
        n = lib().so_initial_hits(self.h, C.byref(sp), _p(masses), _p(intens), C.c_uint64(len(masses)), C.c_float(prec_mz), C.c_uint8(prec_charge),
                                  C.c_float(iso_lo), C.c_float(iso_hi), _p(m), _p(p), _p(c), _p(i), C.c_uint64(cap), C.byref(mp), C.byref(scd))
        n = min(n, cap)
        return dict(matched=m[:n].copy(), peptide=p[:n].copy(), charge=c[:n].copy(), iso=i[:n].copy(), matched_peaks=mp.value,
                    scored_candidates=scd.value)


# ---------------------------------------------------------------- unit helpers
def tolerance_bounds(kind, lo, hi, center):
    out = np.zeros(2, np.float32)
    lib().so_tolerance_bounds(C.c_int(kind), C.c_float(lo), C.c_float(hi), C.c_float(center), _p(out))
    return float(out[0]), float(out[1])


def binary_search_slice(data, low, high):
    out = np.zeros(2, np.uint64)
    data = np.ascontiguousarray(data)
    if data.dtype == np.float32:
        lib().so_binary_search_slice_f32(_p(data), C.c_uint64(len(data)), C.c_float(low), C.c_float(high), _p(out))
    else:
        data = data.astype(np.float64)
        lib().so_binary_search_slice_f64(_p(data), C.c_uint64(len(data)), C.c_double(low), C.c_double(high), _p(out))
    return int(out[0]), int(out[1])


def max_fragment_charge(opt, z):
    return int(lib().so_max_fragment_charge(C.c_int(-1 if opt is None else opt), C.c_uint8(z)))


def bounded_min_heapify(data, k):
    data = np.ascontiguousarray(data)
    if data.dtype == np.uint64:
        lib().so_bounded_min_heapify_u64(_p(data), C.c_uint64(len(data)), C.c_uint64(k))
    else:
        data = data.astype(np.int32)
        lib().so_bounded_min_heapify_i32(_p(data), C.c_uint64(len(data)), C.c_uint64(k))
    return data


def run_ladder(indices):
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.zeros(4, np.uint64)
    lib().so_run(_p(idx), C.c_uint64(len(idx)), _p(out))
    return dict(start=int(out[0]), length=int(out[1]), last=int(out[2]), longest=int(out[3]))


def ion_series(seq: str, kind: str, mods=None, nterm=np.nan, cterm=np.nan):
    out = np.zeros(max(1, len(seq)), np.float32)
    mono = C.c_float(0)
    m = None if mods is None else _f32(mods)
    n = lib().so_ion_series(seq.encode(), _p(m), C.c_float(nterm), C.c_float(cterm), C.c_int(KIND[kind]), _p(out), C.byref(mono))
    return out[:n].copy(), mono.value


def select_most_intense_peak(masses, intens, center, tol, offset=np.nan):
    masses, intens = _f32(masses), _f32(intens)
    r = lib().so_select_most_intense_peak(_p(masses), _p(intens), C.c_uint64(len(masses)), C.c_float(center), C.c_int(tol[0]), C.c_float(tol[1]),
                                          C.c_float(tol[2]), C.c_float(offset))
    return None if r < 0 else r


def deisotope(mz, inten, max_charge, ppm, min_mz, compress=False):
    mz, inten = _f32(mz), _f32(inten)
    n = len(mz)
    oi, oc, oe = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.int64)
    lib().so_deisotope(_p(mz), _p(inten), C.c_uint64(n), C.c_uint8(max_charge), C.c_float(ppm), C.c_float(min_mz), C.c_int(int(compress)), _p(oi),
                       _p(oc), _p(oe))
    return oi, oc, oe


def process_ms2(mz, inten, precursor_charge, take_top_n, deisotope_, min_deisotope_mz=0.0):
    """SpectrumProcessor::new(take_top_n, deisotope, min_deisotope_mz).process(MS2 RawSpectrum) (spectrum.rs:271-412)."""
    mz, inten = _f32(mz), _f32(inten)
    n = len(mz)
    om, oi = np.zeros(n, np.float32), np.zeros(n, np.float32)
    tic = C.c_float(0)
    k = lib().so_process_ms2(_p(mz), _p(inten), C.c_uint64(n), C.c_int(precursor_charge or 0), C.c_uint64(take_top_n), C.c_int(int(deisotope_)),
                             C.c_float(min_deisotope_mz), _p(om), _p(oi), C.byref(tic))
    return om[:k].copy(), oi[:k].copy(), np.float32(tic.value)


def find_reporter_ions(peak_off, masses, intens, labels, tol):
    """tmt.rs:193-211 over a batch -> float32 [n, n_labels] (0 where no peak is within tolerance)."""
    peak_off = np.ascontiguousarray(peak_off, np.uint64)
    masses, intens, labels = _f32(masses), _f32(intens), _f32(labels)
    n = len(peak_off) - 1
    out = np.zeros((n, len(labels)), np.float32)
    lib().so_find_reporter_ions(C.c_uint64(n), _p(peak_off), _p(masses), _p(intens), _p(labels), C.c_uint64(len(labels)), Tol(*tol), _p(out))
    return out


def hyperscore(score_type: int, matched_b: int, matched_y: int, summed_b, summed_y) -> float:
    """Score::hyperscore (scoring.rs:179-201) for one candidate."""
    L = lib()
    L.so_hyperscore.restype = C.c_double
    return float(L.so_hyperscore(C.c_int32(score_type), C.c_uint32(matched_b), C.c_uint32(matched_y), C.c_float(float(summed_b)), C.c_float(float(summed_y))))


def lnfact(n: int) -> float:
    L = lib()
    L.so_lnfact.restype = C.c_double
    return float(L.so_lnfact(C.c_uint32(n)))


def num_threads() -> int:
    return int(lib().so_num_threads())
