"""The CUDA path replaces the two constant divisions of Tolerance::bounds (mass.rs:21-35: /1e6 for ppm, /100 for pct) and fragment / 3
(scoring.rs:707, triply charged fragments) by a reciprocal multiply with one FMA correction (div_const_rn). That is only legal because it is bit-identical to IEEE division on the whole guarded range:
this test proves it exhaustively on the CPU (same IEEE-754 binary32 arithmetic, no contraction)."""
import os
import subprocess
import tempfile


def test_fma_corrected_reciprocal_equals_division_for_every_float_in_range():
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "div_const_check.c")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "div_const_check")
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", exe, src, "-lm"], env=env)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        lines = out.stdout.strip().splitlines()
        assert len(lines) == 3 and all(ln.split()[2] == "1393364419" and ln.endswith("bad 0") for ln in lines), out.stdout
