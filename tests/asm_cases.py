"""Seeded assembler-level cases shared by tests/golden/make_golden_asm.py (which records the outputs of the reference's own
Assembler) and tests/test_oracle_golden.py (which regenerates the same inputs for oracle/asm_oracle.py)."""
import numpy as np

import golden_cases as gc


def _scrambled(M, n, seed):
    co, te, _ = M.cube_mesh(*n)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(co.shape[0])
    co2 = np.empty_like(co)
    co2[perm] = co
    te2 = perm[te]
    te2 = te2[rng.permutation(te2.shape[0])]
    # keep every tet positively oriented (the assembler swaps nodes 2,3 otherwise; both sides do, but the local orders then differ)
    p = co2[te2]
    det = np.linalg.det(p[:, :3, :] - p[:, 3:4, :])
    te2[det < 0] = te2[det < 0][:, [0, 1, 3, 2]]
    return co2 + 0.03 * rng.standard_normal(co2.shape) / max(n), te2


def numbering_cases(M):
    """(name, coords, tets, variables)"""
    out = []
    plain = M.cube_mesh(3, 2, 2)[:2]
    scr = _scrambled(M, (3, 2, 2), 5)
    # re-orient after the jitter of the scrambled mesh
    for tag, (co, te) in (("cube", plain), ("scrambled", scr)):
        p = co[te]
        det = np.linalg.det(p[:, :3, :] - p[:, 3:4, :])
        te = te.copy()
        te[det < 0] = te[det < 0][:, [0, 1, 3, 2]]
        for vname, variables in (("p1", [(gc.P1, 1)]), ("p2", [(gc.P2, 1)]), ("p3", [(gc.P3, 1)]), ("th", [(gc.P2, 3), (gc.P1, 1)]),
                                 ("p0_p3v", [(gc.P0, 1), (gc.P3, 3)]), ("p1v_p2", [(gc.P1, 3), (gc.P2, 1)])):
            out.append(("%s_%s" % (tag, vname), co, te, variables))
    return out


def assembly_cases(M, O):
    """(name, coords, tets, variables, oracle Problem, options)"""
    out = []
    rng = np.random.default_rng(11)
    co, te, _ = M.cube_mesh(3, 3, 2)
    nt = te.shape[0]
    xc = co[te].mean(axis=1)
    K = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, nt, 4)
    c = gc.tensor(rng, gc.T_SCALAR, gc.L_PER_TET, 1, 1, nt, 4)
    Kc = np.array([[1, -1, 0], [-1, 1, 0], [0, 0, 1.0]]).reshape(1, 9)
    out.append(("c1_p1", co, te, [(gc.P1, 1)],
                M.Problem([(gc.P1, 1)],
                          [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=2, ttype=gc.T_SYMMETRIC, layout=gc.L_CONST, D=Kc),
                           dict(trial=0, test=0, opA=gc.IDEN, opB=gc.IDEN, order=2, ttype=gc.T_SCALAR, layout=gc.L_CONST, D=np.ones((1, 1)))],
                          [dict(test=0, opB=gc.IDEN, order=2, ttype=gc.T_SCALAR, layout=gc.L_CONST, D=np.ones((1, 1)))]), {}))
    out.append(("c2_p2", co, te, [(gc.P2, 1)],
                M.Problem([(gc.P2, 1)],
                          [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=2, ttype=gc.T_SYMMETRIC, layout=gc.L_PER_TET, D=K),
                           dict(trial=0, test=0, opA=gc.IDEN, opB=gc.IDEN, order=3, ttype=gc.T_SCALAR, layout=gc.L_PER_TET, D=c, alpha=0.5)],
                          [dict(test=0, opB=gc.IDEN, order=2, ttype=gc.T_SCALAR, layout=gc.L_PER_TET, D=c, alpha=2.0)]), {}))
    # drop_val above some entries: the reference's default scatter then omits them from the pattern (assembler.inl:416)
    out.append(("c2_p2_drop", co, te, [(gc.P2, 1)],
                M.Problem([(gc.P2, 1)], [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=2, ttype=gc.T_SYMMETRIC, layout=gc.L_PER_TET, D=K)], []),
                dict(drop_val=0.05)))
    cs, ts = _scrambled(M, (3, 2, 2), 9)
    Ks = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, ts.shape[0], 14)
    out.append(("p3_scrambled", cs, ts, [(gc.P3, 1)],
                M.Problem([(gc.P3, 1)], [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=4, ttype=gc.T_SYMMETRIC, layout=gc.L_PER_TET, D=Ks)],
                          [dict(test=0, opB=gc.IDEN, order=3, ttype=gc.T_NULL, layout=gc.L_CONST)]), {}))
    co5, te5, _ = M.cube_mesh(2, 2, 2)
    out.append(("c5_taylor_hood", co5, te5, [(gc.P2, 3), (gc.P1, 1)],
                M.Problem([(gc.P2, 3), (gc.P1, 1)],
                          [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=2, ttype=gc.T_NULL, layout=gc.L_CONST),
                           dict(trial=1, test=0, opA=gc.IDEN, opB=gc.DIV, order=2, ttype=gc.T_NULL, layout=gc.L_CONST, alpha=-1.0),
                           dict(trial=0, test=1, opA=gc.DIV, opB=gc.IDEN, order=2, ttype=gc.T_NULL, layout=gc.L_CONST, alpha=-1.0)],
                          [dict(test=0, opB=gc.IDEN, order=2, ttype=gc.T_GENERAL, layout=gc.L_CONST, D=np.array([[0.0, 0.0, -1.0]]))]), {}))
    return out
