import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    p = entry.load_package()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


@pytest.fixture(scope="session")
def oracle():
    O, M = entry.load_oracle()
    O.orc()
    return O


@pytest.fixture(scope="session")
def asm_oracle():
    O, M = entry.load_oracle()
    return M


@pytest.fixture(scope="session")
def ref_tests():
    with open(os.path.join(ROOT, "tests", "golden", "reference_tests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_outputs():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs.npz")))


@pytest.fixture(scope="session")
def ref_face_outputs():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_face_outputs.npz")))


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()
