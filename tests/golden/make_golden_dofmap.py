#!/usr/bin/env python3
"""Golden tables of the reference's local dof maps (fem/tetdofmap.h, compiled into oracle/_ref/libanifem_ref.so) ->
tests/golden/ref_dofmap.npz.  Run in the build container: python tests/golden/make_golden_dofmap.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dofmap_cases as dmc  # noqa: E402

if __name__ == "__main__":
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libanifem_ref.so"))
    out = dmc.collect(L, "ref")
    expected = np.array([e for _, _, e in dmc.EQUALITIES], dtype=np.int32)
    assert np.array_equal(out["equalities"], expected), (out["equalities"], expected)   # the reference agrees with its own test's claims
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_dofmap.npz"), **out)
    print("wrote %d arrays" % len(out))
