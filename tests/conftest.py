import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def config1():
    """Reference fixture inputs (tests/golden/config1_inputs.json, made by make_golden_inputs.py)."""
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_inputs.json")))
    d["mz"] = np.array(d["mz_bits"], dtype=np.uint32).view(np.float32)
    d["intensity"] = np.array(d["intensity_bits"], dtype=np.uint32).view(np.float32)
    return d
