"""Descriptions of local dof maps (Ani::DofT, fem/tetdofmap.h) shared by the golden generator (reference build) and the test of
anifem_b200/dofmap.hpp.  Flat encoding: 1 n0..n5 = UniteDofMap | 2 dim <map> = VectorDofMap | 3 k <map>*k = merge (no
simplification) | 4 k <map>*k = merge_with_simplifications (operator*) | 5 k <map> = operator^."""
import ctypes

import numpy as np

U = lambda n: [1] + list(n)
V = lambda dim, m: [2, dim] + m
C = lambda *ms: [3, len(ms)] + [x for m in ms for x in m]
P = lambda *ms: [4, len(ms)] + [x for m in ms for x in m]
X = lambda k, m: [5, k] + m

ARR1, ARR2 = (3, 2, 1, 1, 3, 4), (1, 2, 0, 1, 0, 3)           # the arrays of tests/fem/tetdofmap_test.cpp
P1, P2, P3 = (1, 0, 0, 0, 0, 0), (1, 1, 0, 0, 0, 0), (1, 2, 0, 1, 0, 0)   # Lagrange spaces; P0 = one cell dof
P0 = (0, 0, 0, 0, 0, 1)
m1, m2 = U(ARR1), U(ARR2)
m3 = V(3, m1)

MAPS = {
    "m1": m1, "m2": m2, "m3_vector": m3, "m4_complex": C(m2, m3), "m5_merge": C(m3, m2),
    "p2": U(P2), "p3": U(P3), "taylor_hood": C(V(3, U(P2)), U(P1)), "p0_p3cubed": C(U(P0), V(3, U(P3))),
    "nested": C(V(2, C(U(P1), U(P2))), U(P3)), "prod_m2_m1m1m1": P(m2, m1, m1, m1), "prod_m4_m3": P(C(m2, m3), m3),
    "pow_of_vector": X(2, m3),
}
# selections: bits of (nodes, edges, faces, cell)
SELECTIONS = [(0, 0, 0b0010, 0), (0b0111, 0b001011, 0b0101, 0), (0, 0b000100, 0, 1), (0b1111, 0b111111, 0b1111, 1), (0b1000, 0, 0, 0)]
# pairs that must compare equal / unequal structurally (the assertions of tetdofmap_test.cpp:147-152 and a few more)
EQUALITIES = [
    (C(m2, m3), P(m2, X(3, m1)), 1), (P(m1, m1, m1), X(3, m1), 1), (P(m1, m1, m2), C(m1, m1, m2), 0), (P(m2, m1, m1, m1), P(m2, X(3, m1)), 1),
    (P(C(m2, m3), m3), P(m2, X(6, m1)), 1), (P(m3, C(m3, m2)), P(X(6, m1), m2), 1), (X(2, m3), V(6, m1), 1), (V(3, m1), V(3, m2), 0),
    (P(V(3, m1), V(2, m1)), C(V(3, m1), V(2, m1)), 1), (P(m1, V(2, m1)), V(3, m1), 1), (P(m1, m2), C(m1, m2), 1), (m1, m2, 0),
]
SPARSITY_OPS = [(3, 0, 1, -1, 0, 0), (2, 1, 1, -1, 0, 0), (2, 2, 1, 1, 5, 1), (2, 0, 0, -1, 0, 0), (1, 4, 1, -1, 0, 0), (3, 0, 1, 2, 3, 1),
                (3, 0, 1, 3, 0, 1), (2, 3, 1, 0, 0, 0), (1, 2, 1, 1, 2, 0), (0, 3, 0, -1, 0, 0)]

ia = lambda v: np.ascontiguousarray(v, dtype=np.int32)
ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def table(lib, prefix, spec, sel=None):
    fn = getattr(lib, prefix + "_dofmap_table")
    fn.restype = ctypes.c_int
    out, by = np.zeros(3 * 4096, dtype=np.int32), np.zeros(4097, dtype=np.int32)
    s = ia(spec)
    n = fn(ptr(s), ptr(out), 4096, ptr(ia(sel)) if sel is not None else None, ptr(by) if sel is not None else None)
    assert n >= 0, (prefix, n)
    return out[:3 * n].reshape(n, 3).copy(), (by[1:1 + by[0]].copy() if sel is not None else None)


def equal(lib, prefix, a, b):
    fn = getattr(lib, prefix + "_dofmap_equal")
    fn.restype = ctypes.c_int
    return fn(ptr(ia(a)), ptr(ia(b)))


def sparsity(lib, prefix, op):
    fn = getattr(lib, prefix + "_sparsity_ops")
    fn.restype = ctypes.c_int
    bits = np.zeros(4, dtype=np.int32)
    assert fn(*[ctypes.c_int(x) for x in op], ptr(bits)) == 0
    return bits


def collect(lib, prefix):
    out = {}
    for name, spec in MAPS.items():
        t, _ = table(lib, prefix, spec)
        out["table_" + name] = t
        for k, sel in enumerate(SELECTIONS):
            out["sel%d_%s" % (k, name)] = table(lib, prefix, spec, sel)[1]
    out["equalities"] = np.array([equal(lib, prefix, a, b) for a, b, _ in EQUALITIES], dtype=np.int32)
    out["sparsity"] = np.stack([sparsity(lib, prefix, op) for op in SPARSITY_OPS])
    return out
