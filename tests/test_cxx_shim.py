"""The C++ mirror of the reference API (inmost-fem_b200/include/anifem_b200/*.hpp): builds on CPU, runs on the GPU."""
import os
import subprocess

import pytest

from conftest import ROOT

CXX_DIR = os.path.join(ROOT, "tests", "cxx")


def test_shim_builds(pkg):
    subprocess.check_call(["make", "-s", "-C", CXX_DIR])
    assert os.path.exists(os.path.join(CXX_DIR, "test_shim"))


@pytest.mark.gpu
def test_shim_runs_reference_style_tests(pkg):
    exe = os.path.join(CXX_DIR, "test_shim")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", CXX_DIR])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "all passed" in out.stdout
