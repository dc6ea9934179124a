"""Seeded inputs for the local Dirichlet helpers (applyDir / applyVectorDir*, fem/operations/dc_on_dof.h), shared by the golden
generator (reference side) and the test (product side)."""
import ctypes

import numpy as np

# (what, n, d, ndc, use_dc_orth, seed): what as in oracle/ref_driver.cpp::ref_dirichlet_local
CASES = [
    (0, 30, 3, 1, False, 1), (0, 30, 3, 2, True, 2), (0, 30, 3, 3, False, 3), (0, 34, 3, 1, True, 4),   # P2^3 (x P1) element matrices
    (1, 30, 3, 2, False, 5), (1, 12, 3, 1, True, 6), (1, 8, 2, 1, False, 7),
    (2, 30, 3, 2, True, 8), (2, 12, 3, 3, False, 9),
    (3, 10, 1, 1, False, 10), (3, 20, 1, 1, False, 11),
    (4, 34, 3, 2, True, 12), (5, 34, 3, 1, False, 13), (5, 34, 3, 2, True, 14),
]


def make_case(case):
    what, n, d, ndc, use_orth, seed = case
    rng = np.random.default_rng(1000 + seed)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    F = rng.standard_normal(n)
    # dofs of one basis function in the d components (component-major local numbering: i, i + n/d', ...), shuffled start
    nb = n // max(d, 1) if what != 3 else n
    i0 = int(rng.integers(0, max(nb, 1)))
    dof_id = np.array([i0 + c * nb for c in range(d)], dtype=np.uint32) if what != 3 else np.array([i0], dtype=np.uint32)
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    Vorth = np.asfortranarray(Q)
    bc = rng.standard_normal(max(ndc, 1))
    dc_orth = rng.permutation(d)[:ndc].astype(np.uint32) if use_orth else None
    return A, F, (dof_id, Vorth, bc, dc_orth)


def call(fn, case, A, F, args):
    what, n, d, ndc, _, _ = case
    dof_id, Vorth, bc, dc_orth = args
    fn.restype = ctypes.c_int
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    return fn(ctypes.c_int(what), ctypes.c_int(n), p(A), p(F), ctypes.c_int(d), p(dof_id), p(Vorth), ctypes.c_int(ndc), p(bc), p(dc_orth))
