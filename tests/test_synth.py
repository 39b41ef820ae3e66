"""The synthetic peptide tables (sage_b200/synth.py; benchmark / test data only): the native variable-modification expansion
(csrc/synth_expand.cpp) and the numpy one give the same table bit for bit, and the table has the shape Peptide::apply /
reorder_peptides give it (peptide.rs:258-305, database.rs:221-258)."""
import os

import numpy as np

from sage_b200 import synth
from sage_b200.build import build_synth_library


def make(native, **kw):
    os.environ["SAGE_B200_SYNTH_NUMPY"] = "0" if native else "1"
    synth._NATIVE = None
    try:
        return synth.make_peptides(**kw)
    finally:
        os.environ.pop("SAGE_B200_SYNTH_NUMPY", None)
        synth._NATIVE = None


def test_native_and_numpy_expansion_agree_and_follow_the_reference_order():
    build_synth_library()
    kw = dict(n_target=20000, seed=3, var_mods=(("M", 15.9949), ("STY", 79.9663)), max_variable_mods=2, static_c=True)
    a, b = make(True, **kw), make(False, **kw)
    for k in ("seq_off", "seq", "mods", "mono", "decoy", "missed"):
        x, y = getattr(a, k), getattr(b, k)
        assert x.shape == y.shape and x.tobytes() == y.tobytes(), k
    base = synth.make_peptides(20000, seed=3, static_c=True)
    assert len(a) > 5 * len(base)                       # every combination of <= 2 modified sites
    assert np.all(np.diff(a.mono) >= 0)                 # sorted by monoisotopic mass
    off = a.seq_off.astype(np.int64)
    # static C everywhere, variable masses only on their residues, at most 2 variable sites per row
    assert np.all(a.mods[a.seq == ord("C")] == np.float32(57.0216))
    var = (a.mods != 0) & (a.seq != ord("C"))
    assert set(np.unique(a.mods[var])) <= {np.float32(15.9949), np.float32(79.9663)}
    assert np.all(np.isin(a.seq[a.mods == np.float32(15.9949)], [ord("M")])) and np.all(np.isin(a.seq[a.mods == np.float32(79.9663)], list(b"STY")))
    nvar = np.add.reduceat(var.astype(np.int64), off[:-1])
    assert nvar.max() == 2 and (nvar == 0).sum() == len(base)
    # rows are unique, and isobaric positional isomers of one sequence are ordered by their modification vectors (peptide.rs:34-52)
    keys = [(a.seq[off[i]:off[i + 1]].tobytes(), a.mods[off[i]:off[i + 1]].tobytes()) for i in range(len(a))]
    assert len(set(keys)) == len(keys)
    same = np.nonzero((np.diff(a.mono) == 0))[0]
    checked = 0
    for i in same[:4000]:
        if keys[i][0] == keys[i + 1][0]:
            assert tuple(a.mods[off[i]:off[i + 1]]) < tuple(a.mods[off[i + 1]:off[i + 2]])
            checked += 1
    assert checked > 100
