import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden_v1.npz"))


@pytest.fixture(scope="session")
def golden_names(golden):
    return [str(n) for n in golden["names"]]


@pytest.fixture(scope="session")
def synth_hashes():
    import json
    with open(os.path.join(GOLDEN_DIR, "synth_hashes.json")) as f:
        return json.load(f)
