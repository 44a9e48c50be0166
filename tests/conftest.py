import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/pyoracle.py, port built on demand."""
    from oracle import pyoracle
    pyoracle.build(ref=False)
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    here = os.path.join(ROOT, "tests", "golden")
    vec = np.load(os.path.join(here, "reference_vectors.npz"))
    with open(os.path.join(here, "reference_synthetic.json")) as f:
        meta = json.load(f)
    return vec, meta


@pytest.fixture(scope="session")
def gpu_ctx():
    import cpvs_b200
    from cpvs_b200 import build
    build.build()
    return cpvs_b200.default_context(0)
