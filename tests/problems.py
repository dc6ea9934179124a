"""Problem definitions of the BASELINE.json configs, shared by the GPU parity tests, smoke() and bench.py.
Each returns (variables, product forms, product rhs forms, oracle Problem) for a mesh with `ntet` tets."""
import numpy as np

import golden_cases as gc


def sym_K(xc):
    """C2 coefficient: SPD K(x) per tet = [[2+x, 1/2, 0],[1/2, 1, -1/4],[0, -1/4, 3]] at the centroid"""
    n = xc.shape[0]
    K = np.zeros((n, 3, 3))
    K[:, 0, 0] = 2 + xc[:, 0]; K[:, 1, 1] = 1; K[:, 2, 2] = 3
    K[:, 0, 1] = K[:, 1, 0] = 0.5
    K[:, 1, 2] = K[:, 2, 1] = -0.25
    return np.ascontiguousarray(K.reshape(n, 9))


def elasticity_C(lam=1.0, mu=3.0):
    """C4: C = lam dd + mu (dd + dd) as a 9x9 matrix on grad u (index 3*comp + d), lin_elast.cpp:101-117"""
    C = np.zeros((9, 9))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    C[3 * i + j, 3 * k + l] = lam * (i == j) * (k == l) + mu * ((i == k) * (j == l) + (i == l) * (j == k))
    return np.ascontiguousarray(C.reshape(1, 81))


def _mk(pkg, M, variables, mats, rhss):
    """mats: (trial, test, opA, opB, order, ttype, layout, D, alpha); rhss: (test, opB, order, ttype, layout, D, alpha)"""
    off, o = [], 0
    for fem, vec in variables:
        off.append(o)
        o += gc.NF[fem] * vec
    forms = [pkg.make_form(opA, variables[a][0], variables[a][1], opB, variables[b][0], variables[b][1], order, tt, lay, D, alpha, off[b], off[a])
             for (a, b, opA, opB, order, tt, lay, D, alpha) in mats]
    rhsf = [pkg.make_form(gc.IDEN, gc.P0, 1, opB, variables[b][0], variables[b][1], order, tt, lay, D, alpha, off[b], 0)
            for (b, opB, order, tt, lay, D, alpha) in rhss]
    prob = None
    if M is not None:
        prob = M.Problem(variables,
                         [dict(trial=a, test=b, opA=opA, opB=opB, order=order, ttype=tt, layout=lay, D=D, alpha=alpha)
                          for (a, b, opA, opB, order, tt, lay, D, alpha) in mats],
                         [dict(test=b, opB=opB, order=order, ttype=tt, layout=lay, D=D, alpha=alpha)
                          for (b, opB, order, tt, lay, D, alpha) in rhss])
    return variables, forms, rhsf, prob


def c1_p1_diffusion(pkg, M, coords, tets):
    """C1: P1, constant symmetric K = [[1,-1,0],[-1,1,0],[0,0,1]] stiffness + mass A=1 + load F=1, order 2
    (examples/Fem/Ani/diffusion.cpp:130-150,192-206)"""
    K = np.array([[1, -1, 0], [-1, 1, 0], [0, 0, 1.0]]).reshape(1, 9)
    return _mk(pkg, M, [(gc.P1, 1)],
               [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_CONST, K, 1.0), (0, 0, gc.IDEN, gc.IDEN, 2, gc.T_SCALAR, gc.L_CONST, np.ones((1, 1)), 1.0)],
               [(0, gc.IDEN, 2, gc.T_SCALAR, gc.L_CONST, np.ones((1, 1)), 1.0)])


def c2_p2_aniso(pkg, M, coords, tets):
    """C2: P2, full symmetric K(x) 3x3 per tet, order 2 (headline config)"""
    xc = coords[tets].mean(axis=1)
    return _mk(pkg, M, [(gc.P2, 1)], [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, sym_K(xc), 1.0)],
               [(0, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)])


def c3_p3_react_diff(pkg, M, coords, tets, xyg4=None, xyg6=None):
    """C3: P3, scalar K(x) stiffness at order 4 (q=14) + reaction A(x) mass at order 6 (q=24), per point
    (examples/tutorials/react_diff1.cpp:125-131 pattern with UFem = FEM_P3)"""
    Kx = np.ascontiguousarray(1 + xyg4[..., 0] ** 2)  # (ntet, q): one scalar per quadrature point
    Ax = np.ascontiguousarray(1 + xyg6[..., 1])
    return _mk(pkg, M, [(gc.P3, 1)],
               [(0, 0, gc.GRAD, gc.GRAD, 4, gc.T_SCALAR, gc.L_PER_POINT, Kx, 1.0), (0, 0, gc.IDEN, gc.IDEN, 6, gc.T_SCALAR, gc.L_PER_POINT, Ax, 1.0)],
               [(0, gc.IDEN, 4, gc.T_NULL, gc.L_CONST, None, 1.0)])


def c4_p2_elasticity(pkg, M, coords, tets):
    """C4: FemVec<3,P2>, constant 9x9 C (lam=1, mu=3), order 2, 30x30 blocks (examples/Fem/Ani/lin_elast.cpp:101-117)"""
    f = np.array([[0.0, 0.0, -1.0]])  # body force as a 3x1 GENERAL tensor for the rhs trick
    return _mk(pkg, M, [(gc.P2, 3)], [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_CONST, elasticity_C(), 1.0)],
               [(0, gc.IDEN, 2, gc.T_GENERAL, gc.L_CONST, f, 1.0)])


def c5_stokes(pkg, M, coords, tets):
    """C5: Taylor-Hood P2^3 x P1: <grad u, grad v> - <p, div v> - <div u, q>, order 2 (examples/Fem/Ani/stokes.cpp:150-157)"""
    return _mk(pkg, M, [(gc.P2, 3), (gc.P1, 1)],
               [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_NULL, gc.L_CONST, None, 1.0),
                (1, 0, gc.IDEN, gc.DIV, 2, gc.T_NULL, gc.L_CONST, None, -1.0),
                (0, 1, gc.DIV, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, -1.0)],
               [(0, gc.IDEN, 2, gc.T_GENERAL, gc.L_CONST, np.array([[0.0, 0.0, -1.0]]), 1.0)])
