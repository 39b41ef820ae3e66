"""Synthetic workloads of the shapes BASELINE.json names: a human-tryptic-scale peptide table and
200-peak MS2 spectra. Product-side (bench.py, tests, smoke) — independent of oracle/.

Peptides follow the reference's digest conventions (trypsin KR|P, 1 missed cleavage, length 5-50, mass 500-5000,
reversed decoys `Peptide::reverse`, sort by (monoisotopic, sequence), dedup) but nothing here needs to be bit-faithful
to the reference's digest: both the oracle and the CUDA path consume exactly these arrays.
"""
from __future__ import annotations

import itertools
import os

import numpy as np

from .api import Peptides, SpectraBatch

H2O = np.float32(18.010565)
PROTON = np.float32(1.0072764)
# mass.rs:64-76
MONO = np.zeros(256, np.float32)
for _aa, _m in zip("ACDEFGHIKLMNPQRSTVWY", [71.03711, 103.00919, 115.02694, 129.04259, 147.0684, 57.02146, 137.05891, 113.08406, 128.09496,
                                           113.08406, 131.0405, 114.04293, 97.05276, 128.05858, 156.1011, 87.03203, 101.04768, 99.06841,
                                           186.07932, 163.06332]):
    MONO[ord(_aa)] = _m
# approximate SwissProt residue frequencies (%), order ACDEFGHIKLMNPQRSTVWY
_FREQ = np.array([8.25, 1.37, 5.45, 6.75, 3.86, 7.07, 2.27, 5.96, 5.84, 9.66, 2.42, 4.06, 4.70, 3.93, 5.53, 6.56, 5.34, 6.87, 1.08, 2.92])
_AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
MAX_LEN = 50


def _digest(prot: np.ndarray, prot_off: np.ndarray, missed: int, min_len: int, max_len: int):
    """Tryptic (KR, not before P) segments with up to `missed` missed cleavages -> (start, end, n_missed)."""
    n = len(prot)
    is_end = np.zeros(n, bool)
    is_end[prot_off[1:] - 1] = True
    nxt_p = np.zeros(n, bool)
    nxt_p[:-1] = prot[1:] == ord("P")
    cut = (((prot == ord("K")) | (prot == ord("R"))) & ~nxt_p) | is_end  # cleave after these positions
    ends = np.nonzero(cut)[0] + 1
    starts = np.concatenate([[0], ends[:-1]])
    seg_prot = np.searchsorted(prot_off, starts, side="right") - 1
    S, E, M = [starts], [ends], [np.zeros(len(starts), np.uint8)]
    for m in range(1, missed + 1):
        ok = seg_prot[:-m] == seg_prot[m:]
        S.append(starts[:-m][ok])
        E.append(ends[m:][ok])
        M.append(np.full(int(ok.sum()), m, np.uint8))
    s, e, mm = np.concatenate(S), np.concatenate(E), np.concatenate(M)
    ln = e - s
    keep = (ln >= min_len) & (ln <= max_len)
    return s[keep], e[keep], mm[keep]


def _gather(prot, s, e):
    ln = (e - s).astype(np.int64)
    idx = s[:, None] + np.arange(MAX_LEN)[None, :]
    mat = prot[np.minimum(idx, len(prot) - 1)]
    mat[np.arange(MAX_LEN)[None, :] >= ln[:, None]] = 0
    return mat, ln


def _reverse_inner(mat, ln):
    """Peptide::reverse (peptide.rs:307-318): reverse residues 1..len-1, keep first and last."""
    out = mat.copy()
    j = np.arange(MAX_LEN)[None, :]
    src = np.where((j >= 1) & (j < ln[:, None] - 1), ln[:, None] - 1 - j, j)
    inner = ln > 2
    out[inner] = np.take_along_axis(mat[inner], src[inner], axis=1)
    return out


def _expand_variable_mods(site_mass, max_variable_mods):
    """Peptide::apply (peptide.rs:258-305): every combination of up to `max_variable_mods` variable modifications on distinct sites
    (site_mass != 0 marks a site). Returns the base row of every form (the unmodified forms first) and the (form, column) pairs to modify."""
    n = len(site_mass)
    is_site = site_mass != 0
    nsite = is_site.sum(axis=1)
    maxs = int(nsite.max()) if n else 0
    order = np.argsort(~is_site, axis=1, kind="stable")[:, :max(maxs, 1)]   # order[i, k] = column of the k-th site of row i
    base_rows, add_rows, add_cols = [np.arange(n)], [], []
    form_count = n
    for k in range(1, max_variable_mods + 1):
        for slots in itertools.combinations(range(maxs), k):
            rows = np.nonzero(nsite > slots[-1])[0]
            if len(rows) == 0:
                continue
            base_rows.append(rows)
            ids = form_count + np.arange(len(rows))
            for sl in slots:
                add_rows.append(ids)
                add_cols.append(order[rows, sl])
            form_count += len(rows)
    base = np.concatenate(base_rows)
    if add_rows:
        return base, np.concatenate(add_rows), np.concatenate(add_cols)
    return base, np.zeros(0, np.int64), np.zeros(0, np.int64)


_NATIVE = None


def _native():
    """libsage_synth.so (sage_b200/csrc/synth_expand.cpp, built by sage_b200.build): same result as the numpy path, ~100x faster."""
    global _NATIVE
    if _NATIVE is None:
        import ctypes as C
        from .build import synth_library_path
        path = synth_library_path()
        if os.environ.get("SAGE_B200_SYNTH_NUMPY") == "1" or not os.path.exists(path):
            _NATIVE = False
        else:
            lib = C.CDLL(path)
            lib.synth_expand.restype = C.c_void_p
            lib.synth_expand.argtypes = [C.c_uint64, C.c_uint32] + [C.c_void_p] * 7 + [C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
            lib.synth_expand_fetch.argtypes = [C.c_void_p] * 7
            lib.synth_expand_free.argtypes = [C.c_void_p]
            _NATIVE = lib
    return _NATIVE or None


def _expand_native(lib, mat, ln, static, site_mass, base_mono, decoy, missed, max_mods):
    import ctypes as C
    mat, ln = np.ascontiguousarray(mat, np.uint8), np.ascontiguousarray(ln, np.int64)
    static, site_mass, base_mono = (np.ascontiguousarray(x, np.float32) for x in (static, site_mass, base_mono))
    decoy, missed = np.ascontiguousarray(decoy, np.uint8), np.ascontiguousarray(missed, np.uint8)
    n_out, n_res = C.c_uint64(0), C.c_uint64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    h = lib.synth_expand(len(mat), mat.shape[1], p(mat), p(ln), p(static), p(site_mass), p(base_mono), p(decoy), p(missed), max_mods,
                         C.c_float(500.0), C.c_float(5000.0), C.byref(n_out), C.byref(n_res))
    if not h:
        raise RuntimeError("synth_expand failed (too many forms / residues for u32 offsets)")
    n, r = n_out.value, n_res.value
    seq_off, seq, mods = np.zeros(n + 1, np.uint32), np.zeros(r, np.uint8), np.zeros(r, np.float32)
    mono, dec, mis = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    lib.synth_expand_fetch(h, p(seq_off), p(seq), p(mods), p(mono), p(dec), p(mis))
    lib.synth_expand_free(h)
    return Peptides(seq_off=seq_off, seq=seq, mods=mods, nterm=np.full(n, np.nan, np.float32), mono=mono, decoy=dec, missed=mis)


def make_peptides(n_target: int = 2_000_000, seed: int = 0x5A6E, missed: int = 1, static_c: bool = False, il_twin_fraction: float = 0.02,
                  var_mod_m: bool = False, var_mods=(), max_variable_mods: int = 2) -> Peptides:
    """Peptide table sorted like reorder_peptides (database.rs:221-258). `n_target` ~ rows (targets + reversed decoys) BEFORE the variable
    modifications are enumerated; var_mods = ((residues, mass), ...) adds every combination of up to max_variable_mods modified sites
    (peptide.rs:258-305), e.g. (("M", 15.9949), ("STY", 79.9663)) turns ~1.9 M rows into ~15 M. var_mod_m is shorthand for M+15.9949 with
    at most one modified site per form."""
    rng = np.random.default_rng(seed)
    if var_mod_m and not var_mods:
        var_mods, max_variable_mods = (("M", 15.9949),), 1
    # ~1 unique peptide (0+1 missed, len 5-50, after dedup) per 5.3 residues; decoys double it
    n_res = int(n_target * 2.75) + 4096
    lens = np.maximum(30, rng.lognormal(np.log(375.0), 0.6, size=max(8, n_res // 430))).astype(np.int64)
    prot_off = np.concatenate([[0], np.cumsum(lens)])
    prot = rng.choice(_AA, size=int(prot_off[-1]), p=_FREQ / _FREQ.sum())
    if il_twin_fraction > 0:  # I/L-swapped protein copies -> isobaric twin peptides (exercise tie order)
        k = max(1, int(len(lens) * il_twin_fraction))
        extra, extra_len = [], []
        for pi in rng.choice(len(lens), size=k, replace=False):
            seg = prot[prot_off[pi]:prot_off[pi + 1]].copy()
            il = np.nonzero((seg == ord("I")) | (seg == ord("L")))[0]
            flip = il[rng.random(len(il)) < 0.3]
            seg[flip] = np.where(seg[flip] == ord("I"), ord("L"), ord("I")).astype(np.uint8)
            extra.append(seg)
            extra_len.append(len(seg))
        prot = np.concatenate([prot] + extra)
        prot_off = np.concatenate([prot_off, prot_off[-1] + np.cumsum(extra_len)])
    s, e, mm = _digest(prot, prot_off, missed, 5, MAX_LEN)
    mat, ln = _gather(prot, s, e)
    # unique target sequences
    key = np.ascontiguousarray(mat).view(f"S{MAX_LEN}").ravel()
    _, first = np.unique(key, return_index=True)
    mat, ln, mm = mat[first], ln[first], mm[first]
    tkey = np.ascontiguousarray(mat).view(f"S{MAX_LEN}").ravel()
    dmat = _reverse_inner(mat, ln)
    dkey = np.ascontiguousarray(dmat).view(f"S{MAX_LEN}").ravel()
    _, dfirst = np.unique(dkey, return_index=True)
    dkeep = dfirst[~np.isin(dkey[dfirst], tkey)]  # decoys equal to a target sequence are dropped (database.rs:212)
    allmat = np.concatenate([mat, dmat[dkeep]])
    allln = np.concatenate([ln, ln[dkeep]])
    allmm = np.concatenate([mm, mm[dkeep]])
    decoy = np.concatenate([np.zeros(len(mat), np.uint8), np.ones(len(dkeep), np.uint8)])
    # rows in sequence order: the final stable sort by mass then yields (monoisotopic, sequence, modifications-in-enumeration-order)
    sorder = np.argsort(np.ascontiguousarray(allmat).view(f"S{MAX_LEN}").ravel(), kind="stable")
    allmat, allln, allmm, decoy = allmat[sorder], allln[sorder], allmm[sorder], decoy[sorder]
    res = MONO[allmat]
    base = np.cumsum(np.concatenate([np.full((len(allmat), 1), H2O, np.float32), res], axis=1), axis=1, dtype=np.float32)[:, -1]
    valid = np.arange(MAX_LEN)[None, :] < allln[:, None]
    static = np.zeros(allmat.shape, np.float32)
    if static_c:
        static[allmat == ord("C")] = np.float32(57.0216)
    if var_mods:
        # variable mods first, then static mods on the sites still unmodified (peptide.rs:293-300); disjoint residue sets here
        site_mass = np.zeros(allmat.shape, np.float32)
        for residues, mass in var_mods:
            for r in residues:
                site_mass[(allmat == ord(r)) & (static == 0) & valid] = np.float32(mass)
        lib = _native()
        if lib is not None:
            return _expand_native(lib, allmat, allln, static, site_mass, base, decoy, allmm, int(max_variable_mods))
        form_base, ar, ac = _expand_variable_mods(site_mass, int(max_variable_mods))
        nforms = len(form_base)
        if nforms > 4_000_000:
            raise RuntimeError("variable-modification tables of this size need the native helper (python -m sage_b200.build)")
        fmods = static[form_base]
        fmods[ar, ac] = site_mass[form_base[ar], ac]
        # modification_mass (peptide.rs:129-133): sequential f32 sum over the residues
        mono = (base[form_base] + np.cumsum(fmods, axis=1, dtype=np.float32)[:, -1]).astype(np.float32)
        keep = np.nonzero((mono >= np.float32(500.0)) & (mono <= np.float32(5000.0)))[0]
        # reorder_peptides: (monoisotopic, sequence == base row, modifications lexicographic); mods >= 0, so big-endian bit patterns sort like values
        mkey = np.ascontiguousarray(fmods[keep].view(np.uint32).astype(">u4")).view(f"S{4 * MAX_LEN}").ravel()
        order = keep[np.lexsort((mkey, form_base[keep], mono[keep]))]
        fb = form_base[order]
        fvalid = valid[fb]
        seq_off = np.concatenate([[0], np.cumsum(allln[fb])]).astype(np.uint32)
        return Peptides(seq_off=seq_off, seq=allmat[fb][fvalid].astype(np.uint8), mods=fmods[order][fvalid].astype(np.float32),
                        nterm=np.full(len(order), np.nan, np.float32), mono=mono[order], decoy=decoy[fb], missed=allmm[fb])
    mods = static
    # monoisotopic = H2O + sum(residues) (sequential f32, peptide.rs:361-373) + modification_mass (peptide.rs:129-133)
    modsum = np.cumsum(mods, axis=1, dtype=np.float32)[:, -1]
    mono = (base + modsum).astype(np.float32)
    keep = (mono >= np.float32(500.0)) & (mono <= np.float32(5000.0))
    allmat, allln, allmm, decoy, mods, mono = allmat[keep], allln[keep], allmm[keep], decoy[keep], mods[keep], mono[keep]
    order = np.argsort(mono, kind="stable")   # rows are in sequence order already: (monoisotopic, sequence)
    allmat, allln, allmm, decoy, mods, mono = allmat[order], allln[order], allmm[order], decoy[order], mods[order], mono[order]
    valid = np.arange(MAX_LEN)[None, :] < allln[:, None]
    seq_off = np.concatenate([[0], np.cumsum(allln)]).astype(np.uint32)
    return Peptides(seq_off=seq_off, seq=allmat[valid].astype(np.uint8), mods=mods[valid].astype(np.float32),
                    nterm=np.full(len(mono), np.nan, np.float32), mono=mono, decoy=decoy, missed=allmm)


def make_spectra(pep: Peptides, n: int = 50_000, seed: int = 0xB202, n_peaks: int = 200, chimeric: bool = False, charge_known: bool = True) -> SpectraBatch:
    """Synthetic MS2 spectra: a target peptide's b/y ions at z=1 (and z=2 for 30% of 3+ precursors), 50% dropout,
    4 ppm jitter, padded with uniform noise to exactly n_peaks; lognormal intensities (signal x3); masses sorted ascending."""
    rng = np.random.default_rng(seed)
    targets = np.nonzero(pep.decoy == 0)[0]
    ln_all = np.diff(pep.seq_off.astype(np.int64))

    def one_component(choice):
        ln = ln_all[choice]
        idx = pep.seq_off[choice].astype(np.int64)[:, None] + np.arange(MAX_LEN)[None, :]
        valid = np.arange(MAX_LEN)[None, :] < ln[:, None]
        idx = np.minimum(idx, len(pep.seq) - 1)
        rm = np.where(valid, MONO[pep.seq[idx]] + pep.mods[idx], np.float32(0)).astype(np.float32)
        b = np.cumsum(rm, axis=1, dtype=np.float32)  # b_i, i = 1..L (last one is not an ion)
        mono = pep.mono[choice]
        y = mono[:, None] - b
        ion_ok = np.arange(MAX_LEN)[None, :] < (ln[:, None] - 1)
        return b, y, ion_ok, mono

    choice = rng.choice(targets, size=n)
    z = np.where(rng.random(n) < 0.6, 2, 3).astype(np.uint8)
    b, y, ion_ok, mono = one_component(choice)
    prec_mz = ((mono.astype(np.float64) + z * float(PROTON)) / z * (1.0 + rng.normal(0, 3e-6, n))).astype(np.float32)
    comps = [(b, y, ion_ok, np.ones(n, bool))]
    if chimeric:  # second, co-isolated peptide whose precursor m/z is within 1 Th (a neighbour in the mass-sorted table)
        mass_rank = np.searchsorted(pep.mono[targets], mono)
        other = targets[np.clip(mass_rank + rng.integers(-200, 200, n), 0, len(targets) - 1)]
        b2, y2, ok2, _ = one_component(other)
        comps.append((b2, y2, ok2, np.ones(n, bool)))
    sig_m, sig_ok, sig_w = [], [], []
    for ci, (bb, yy, ok, _) in enumerate(comps):
        z2 = (z == 3) & (rng.random(n) < 0.3)
        for ions in (bb, yy):
            for fc, allow in ((1, np.ones(n, bool)), (2, z2)):
                keep = ok & allow[:, None] & (rng.random(ions.shape) < 0.5)
                m = ions.astype(np.float64) / fc * (1.0 + rng.normal(0, 4e-6, ions.shape))
                sig_m.append(m)
                sig_ok.append(keep & (m > 50.0))
                sig_w.append(np.full(ions.shape, 3.0 * (0.6 if (chimeric and ci == 0) else (0.4 if chimeric else 1.0)) / (1.0 if not chimeric else 0.6)))
    sig_m, sig_ok, sig_w = np.concatenate(sig_m, axis=1), np.concatenate(sig_ok, axis=1), np.concatenate(sig_w, axis=1)
    hi = np.minimum(2000.0, mono.astype(np.float64))
    masses = (150.0 + rng.random((n, n_peaks)) * (hi - 150.0)[:, None]) - float(PROTON)
    intens = rng.lognormal(8.0, 1.2, (n, n_peaks))
    # overwrite the first k_i slots of each row with that row's signal peaks
    rank = np.cumsum(sig_ok, axis=1) - 1
    take = sig_ok & (rank < n_peaks)
    rows = np.nonzero(take)[0]
    cols = rank[take]
    masses[rows, cols] = sig_m[take]
    intens[rows, cols] = intens[rows, cols] * sig_w[take]
    masses = masses.astype(np.float32)
    intens = intens.astype(np.float32)
    order = np.argsort(masses, axis=1, kind="stable")
    masses = np.take_along_axis(masses, order, axis=1)
    intens = np.take_along_axis(intens, order, axis=1)
    tic = np.cumsum(intens, axis=1, dtype=np.float32)[:, -1]  # sequential f32 sum (spectrum.rs:398)
    peak_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(n_peaks))
    chg = z if charge_known else np.zeros(n, np.uint8)
    return SpectraBatch(peak_off=peak_off, masses=masses.ravel(), intensities=intens.ravel(), prec_mz=prec_mz, prec_charge=chg,
                        iso_lo=np.full(n, np.nan, np.float32), iso_hi=np.full(n, np.nan, np.float32), tic=tic, level=np.full(n, 2, np.uint8),
                        rt=(np.arange(n, dtype=np.float32) * np.float32(0.01)), ims=np.full(n, np.nan, np.float32))
