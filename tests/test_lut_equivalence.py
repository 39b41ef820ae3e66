"""The CUDA path replaces full binary searches by LUT-bracketed ones (spectrum LUT built in O(cells + peaks); precursor-mass LUT of
k_setup_queries). tests/lut_equivalence_check.c restates both on the CPU with the same binary32 operations and checks them against the plain
searches on random, clustered, duplicate-heavy and out-of-range inputs."""
import os
import subprocess
import tempfile


def test_lut_accelerated_searches_equal_plain_searches():
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lut_equivalence_check.c")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "lut_equivalence_check")
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], env=env)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "20000 arrays, 0 mismatching cells" in out.stdout and "2400000 queries, 0 mismatches" in out.stdout, out.stdout
