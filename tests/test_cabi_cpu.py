"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/sage_b200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present. No compute calls."""
import os
import re

import numpy as np
import pytest

from sage_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sage_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sage_b200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(api.EXPORTED_SYMBOLS) == syms


def test_struct_layouts_match_header(tmp_path):
    # compile the header with the C compiler and compare struct sizes with the ctypes mirrors
    import ctypes as C
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(sage_b200_scorer_params),'
                   'sizeof(sage_b200_spectra), sizeof(sage_b200_counters), sizeof(sage_b200_feature), sizeof(sage_b200_db_info),'
                   'sizeof(sage_b200_peptides), sizeof(sage_b200_index));return 0;}\n' % os.path.join(ROOT, "include", "sage_b200.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(api.CScorerParams), C.sizeof(api.CSpectra), C.sizeof(api.CCounters), api.FEATURE_DTYPE.itemsize, C.sizeof(api.CDbInfo),
                     C.sizeof(api.CPeptides), C.sizeof(api.CIndex)]


@pytest.mark.skipif(api.device_count() > 0, reason="needs a box without GPUs")
def test_fails_loudly_without_gpu():
    from sage_b200 import IndexedDatabase, SageB200Error, synth
    pep = synth.make_peptides(500, seed=3)
    with pytest.raises(SageB200Error) as e:
        IndexedDatabase.build_from_peptides(pep)
    assert e.value.code == -2 and "no CPU fallback" in e.value.message


def test_synth_shapes():
    from sage_b200 import synth
    pep = synth.make_peptides(5000, seed=5)
    assert np.all(np.diff(pep.mono) >= 0) and pep.mono.min() >= 500 and pep.mono.max() <= 5000
    sp = synth.make_spectra(pep, 64, seed=6)
    assert len(sp) == 64 and np.all(np.diff(sp.peak_off) == 200)
    assert np.all(np.diff(sp.masses.reshape(64, 200), axis=1) >= 0)
