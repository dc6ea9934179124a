#!/usr/bin/env python3
"""Golden vectors of the reference's local Dirichlet helpers (fem/operations/dc_on_dof.h) -> tests/golden/ref_dirichlet_local.npz.
Run in the build container (needs oracle/_ref/libanifem_ref.so built from /root/reference): python tests/golden/make_golden_dc.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dc_cases import CASES, make_case, call  # noqa: E402


def main():
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libanifem_ref.so"))
    out = {}
    for k, case in enumerate(CASES):
        A, F, args = make_case(case)
        rc = call(L.ref_dirichlet_local, case, A, F, args)
        assert rc == 0, case
        out["A%d" % k], out["F%d" % k] = A, F
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_dirichlet_local.npz"), **out)
    print("wrote %d cases" % len(CASES))


if __name__ == "__main__":
    main()
