#!/usr/bin/env python3
"""CPU check of the algebra behind "class merging" in k_rows_cl (profiles/r01e_rows_phase_costs.md, lever 4): for P2 and a symmetric
tensor the element matrix is linear in the six off-diagonal entries G[m][n] = |T| grad(lambda_m) . K grad(lambda_n), m < n, of the
barycentric Gram matrix (its diagonal follows from the zero row sums), and relabelling the vertices of the tet by a permutation rho
only permutes those six entries (the edges of the tet) and the ten local dofs.  Hence one table row per *kind* of local row is enough:
    A_e(i, j) = sum_c T[c][i0][pi_i(j)] * g_e[sigma_i(c)],   i0 = 0 for vertex rows, 4 for edge rows,
with per-row-class permutations (sigma_i, pi_i) induced by a vertex relabelling rho_i that moves local row i to position i0.
The visit classes of the gather then collapse from 10 to 2 (tests/analysis/plan_stats.py: 9.6 % fewer visit-steps on the cube).

TEST-SIDE ANALYSIS (uses the oracle for the element matrices); nothing here is product code.  Prints the tables and the error.
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import golden_cases as gc  # noqa: E402

EDGES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def offdiag_G(XY, K):
    """g[f, c] = |T| grad(lambda_m) . K grad(lambda_n) for the six edges (m, n) of every tet; XY (4, f, 3), K (f, 3, 3) symmetric"""
    f = XY.shape[1]
    g = np.zeros((f, 6))
    for r in range(f):
        P = XY[:, r, :]
        Mx = np.vstack([np.ones(4), P.T])            # lambda coefficients: Mx^T [a; b] = e_m
        C = np.linalg.inv(Mx)                        # row m = (a_m, grad lambda_m)
        grads = C[:, 1:]
        vol = abs(np.linalg.det(P[1:] - P[0])) / 6
        for c, (m, n) in enumerate(EDGES):
            g[r, c] = vol * grads[m] @ K[r] @ grads[n]
    return g


def induced(rho):
    """dof permutation pi (10) and edge permutation sigma (6) induced by the vertex relabelling rho (old vertex v -> new label rho[v])"""
    pi = list(rho)
    sigma_old_of_new = [0] * 6
    for c, (a, b) in enumerate(EDGES):
        na, nb = sorted((rho[a], rho[b]))
        cn = EDGES.index((na, nb))
        pi.append(4 + cn)
        sigma_old_of_new[cn] = c                     # new edge cn carries the value of old edge c
    return np.array(pi), np.array(sigma_old_of_new)


def main():
    O, M = entry.load_oracle()
    rng = np.random.default_rng(1)
    f = 64
    XY = gc.random_tets(rng, f)
    Kt = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, f, 4)
    form = (gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET)
    A = O.fem3dtet(form, XY, Kt)                     # (f, ia, ib)
    g = offdiag_G(XY, Kt.reshape(f, 3, 3))
    # T[c][i][j] by least squares: the model A = sum_c T[c] g[c] is exact, the residual measures that
    T, res, rank, _ = np.linalg.lstsq(g, A.reshape(f, 100), rcond=None)
    fit = np.abs(g @ T - A.reshape(f, 100)).max() / np.abs(A).max()
    print("linear model in the six off-diagonal G entries: rank %d, max residual %.2e (relative)" % (rank, fit))
    T = T.reshape(6, 10, 10)                         # [c][ia][ib]; symmetric in (ia, ib)
    # one relabelling per local row class
    worst = 0.0
    print("row i -> rho_i (old vertex -> new label), sigma_i (g index read for table component c), pi_i (table column of local dof j)")
    for i in range(10):
        target_vertices = (i,) if i < 4 else EDGES[i - 4]
        rho = None
        for cand in itertools.permutations(range(4)):
            if all(cand[v] == k for k, v in enumerate(target_vertices)):
                rho = cand
                break
        pi, sig = induced(rho)
        i0 = 0 if i < 4 else 4
        assert pi[i] == i0
        # A(i, j) = sum_c T[c][i0][pi(j)] * g[sigma(c)]
        pred = np.einsum("fc,cj->fj", g[:, sig], T[:, i0, :][:, pi])
        err = np.abs(pred - A[:, i, :]).max() / np.abs(A).max()
        worst = max(worst, err)
        print("  i=%d  rho=%s  sigma=%s  pi=%s  err %.1e" % (i, list(rho), [int(x) for x in sig], [int(x) for x in pi], err))
    print("max relative error of the merged-class evaluation: %.2e" % worst)
    assert fit < 1e-12 and worst < 1e-12
    return 0


if __name__ == "__main__":
    sys.exit(main())
