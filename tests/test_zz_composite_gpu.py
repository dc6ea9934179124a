"""GPU leg of the composite-space and fem3DfaceN front ends (composite.hpp, face_normal.hpp): tests/cxx/test_composite.cpp built with
-DGPU_FRONT_END runs fem3Dtet<Operator<.., FemCom/FemVecT>>, the runtime ComplexFemSpace overload and fem3DfaceN through the
library (afb_fem3dtet_batched / afb_fem3dface_batched on the GPU) and compares with the reference build's own composite operators
and fem3DfaceN (oracle/_ref/libanifem_ref.so, the checker) on the same seeded data, tolerance 1e-12 of the matrix scale."""
import os
import subprocess

import pytest

from conftest import ROOT

CXX_DIR = os.path.join(ROOT, "tests", "cxx")


@pytest.mark.gpu
def test_composite_and_face_normal_front_ends_on_the_gpu(pkg):
    exe = os.path.join(CXX_DIR, "test_composite_gpu")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libanifem_ref.so")):
        pytest.skip("oracle/_ref (the reference build) is not present")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", CXX_DIR, "test_composite_gpu"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "all passed" in out.stdout
