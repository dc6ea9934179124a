"""GPU parity tests of the global-assembly path through the C ABI: mesh generator, NATURAL dof map,
CSR pattern (bit-exact against the restated AssembleTemplate), values and RHS (1e-12 relative),
status codes, accumulate semantics, explicit dof tables with ghost rows (multi-rank layout)."""
import os

import numpy as np
import pytest

import golden_cases as gc
import problems

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _setup(pkg, ctx, M, n, variables):
    nx, ny, nz = n
    ctx.mesh_cube(nx, ny, nz)
    coords, tets = ctx.mesh_get()
    co, te, _ = M.cube_mesh(nx, ny, nz)
    assert np.array_equal(te, tets), "cube connectivity differs from the restated generator"
    assert np.array_equal(co, coords), "cube coordinates differ bit-wise from the restated generator"
    ctx.dofmap_natural(variables)
    dm = M.DofMap(te, variables, nnode=co.shape[0])
    row, col = ctx.dofmap_get()
    assert np.array_equal(col, dm.elem2dof + 1), "NATURAL numbering differs from the restated enumerator"
    assert np.array_equal(row, col)
    nr, nc, rb, re, ng = ctx.dofmap_info()
    assert (nr, nc, rb, re, ng) == (dm.nloc, dm.nloc, 0, dm.nrows, dm.nrows)
    return co, te, dm


def _check(ctx, M, prob, forms, rhsf, co, te, dm, tag):
    nnz = ctx.pattern_build()
    rowptr, colind = ctx.pattern_get()
    rp, ci, v, r, st = M.assemble(prob, co, te, dm)
    assert np.array_equal(rowptr, rp), tag + ": rowptr not bit-exact"
    assert np.array_equal(colind, ci), tag + ": colind not bit-exact"
    assert st == 0
    # both product paths: the fused tensor-representation path (when it applies) and the generic staged path
    for env in ("", "1"):
        if env:
            os.environ["AFB_DISABLE_TENSOR_PATH"] = env
        else:
            os.environ.pop("AFB_DISABLE_TENSOR_PATH", None)
        val, rhs = np.full(nnz, np.nan), np.full(rp.size - 1, np.nan)
        status = ctx.assemble(forms, rhsf, val, rhs)
        os.environ.pop("AFB_DISABLE_TENSOR_PATH", None)
        assert status == 0
        _compare(val, rhs, v, r, rp, te, nnz, tag + (" [generic path]" if env else " [default path]"))
    return val, rhs, v, r


def _compare(val, rhs, v, r, rp, te, nnz, tag):
    # H7 tolerance: relative to the largest element contribution ~ largest |entry| of the row
    rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
    scale = np.repeat(rowmax, np.diff(rp))
    err = (np.abs(val - v) / scale).max()
    rerr = np.abs(rhs - r).max() / np.abs(r).max()
    print("%s: ntet=%d nnz=%d max rel err A %.2e rhs %.2e" % (tag, te.shape[0], nnz, err, rerr))
    assert err <= RTOL and rerr <= RTOL, (tag, err, rerr)


def test_c1_p1_diffusion(pkg, ctx, asm_oracle):
    co, te, dm = _setup(pkg, ctx, asm_oracle, (5, 4, 3), [(gc.P1, 1)])
    _, forms, rhsf, prob = problems.c1_p1_diffusion(pkg, asm_oracle, co, te)
    _check(ctx, asm_oracle, prob, forms, rhsf, co, te, dm, "C1")


def test_c2_p2_aniso(pkg, ctx, asm_oracle):
    co, te, dm = _setup(pkg, ctx, asm_oracle, (4, 5, 3), [(gc.P2, 1)])
    _, forms, rhsf, prob = problems.c2_p2_aniso(pkg, asm_oracle, co, te)
    _check(ctx, asm_oracle, prob, forms, rhsf, co, te, dm, "C2")


def test_c3_p3_reaction_diffusion(pkg, ctx, asm_oracle, oracle):
    co, te, dm = _setup(pkg, ctx, asm_oracle, (3, 3, 2), [(gc.P3, 1)])
    XY = co[te].transpose(1, 0, 2)
    _, forms, rhsf, prob = problems.c3_p3_react_diff(pkg, asm_oracle, co, te, oracle.quad_points(4, XY), oracle.quad_points(6, XY))
    _check(ctx, asm_oracle, prob, forms, rhsf, co, te, dm, "C3")


def test_c4_p2_elasticity(pkg, ctx, asm_oracle):
    co, te, dm = _setup(pkg, ctx, asm_oracle, (3, 2, 2), [(gc.P2, 3)])
    _, forms, rhsf, prob = problems.c4_p2_elasticity(pkg, asm_oracle, co, te)
    _check(ctx, asm_oracle, prob, forms, rhsf, co, te, dm, "C4")


def test_c5_taylor_hood_stokes(pkg, ctx, asm_oracle):
    co, te, dm = _setup(pkg, ctx, asm_oracle, (2, 3, 2), [(gc.P2, 3), (gc.P1, 1)])
    _, forms, rhsf, prob = problems.c5_stokes(pkg, asm_oracle, co, te)
    _check(ctx, asm_oracle, prob, forms, rhsf, co, te, dm, "C5")


def test_accumulate_and_determinism(pkg, ctx, asm_oracle):
    """Assemble adds into existing contents (assembler.inl:305-306); two runs are bit-identical"""
    co, te, dm = _setup(pkg, ctx, asm_oracle, (3, 3, 3), [(gc.P2, 1)])
    _, forms, rhsf, prob = problems.c2_p2_aniso(pkg, asm_oracle, co, te)
    nnz = ctx.pattern_build()
    nrows = dm.nrows
    a, fa = np.zeros(nnz), np.zeros(nrows)
    b, fb = np.zeros(nnz), np.zeros(nrows)
    assert ctx.assemble(forms, rhsf, a, fa) == 0
    assert ctx.assemble(forms, rhsf, b, fb) == 0
    assert np.array_equal(a, b) and np.array_equal(fa, fb), "assembly is not bit-reproducible"
    assert ctx.assemble(forms, rhsf, b, fb, accumulate=True) == 0
    assert np.array_equal(b, 2 * a) and np.array_equal(fb, 2 * fa)
    # matrix only / rhs only (AssembleMatrix / AssembleRHS)
    c = np.zeros(nnz)
    assert ctx.assemble(forms, [], c, None) == 0 and np.array_equal(c, a)
    # the load alone is assembled by the row gather, together with the matrix by the ring kernel: the same sums in another order
    fc = np.zeros(nrows)
    assert ctx.assemble([], rhsf, None, fc) == 0 and np.abs(fc - fa).max() <= 1e-14 * np.abs(fa).max()


def test_nan_status(pkg, ctx, asm_oracle):
    """a non-finite local value returns -1 (assembler.inl:419-424,475-479)"""
    co, te, dm = _setup(pkg, ctx, asm_oracle, (2, 2, 2), [(gc.P1, 1)])
    nnz = ctx.pattern_build()
    K = np.ones((te.shape[0], 1))
    K[7, 0] = np.nan
    forms = [pkg.make_form(gc.GRAD, gc.P1, 1, gc.GRAD, gc.P1, 1, 2, gc.T_SCALAR, gc.L_PER_TET, K)]
    assert ctx.assemble(forms, [], np.zeros(nnz), None) == -1
    K[7, 0] = np.inf
    assert ctx.assemble(forms, [], np.zeros(nnz), None) == -1
    K[7, 0] = 1.0
    assert ctx.assemble(forms, [], np.zeros(nnz), None) == 0


def test_call_order_errors(pkg, asm_oracle):
    """missing mesh / dof map / pattern -> error like the reference's runtime_error (assembler.inl:195-196,316-317)"""
    c = pkg.Context(0)
    with pytest.raises(pkg.AfbError) as e:
        c.dofmap_natural([(gc.P1, 1)])
    assert e.value.code == -6 and "Mesh was not specified" in str(e.value)
    c.mesh_cube(2, 2, 2)
    with pytest.raises(pkg.AfbError) as e:
        c.pattern_build()
    assert e.value.code == -6
    c.dofmap_natural([(gc.P1, 1)])
    with pytest.raises(pkg.AfbError) as e:
        c.assemble([pkg.make_form(gc.GRAD, gc.P1, 1, gc.GRAD, gc.P1, 1, 2, gc.T_NULL, gc.L_CONST)], [], np.zeros(10), None)
    assert e.value.code == -6
    c.close()


def test_explicit_dofmap_with_ghost_rows(pkg, ctx, asm_oracle):
    """two-rank layout: each rank gets explicit index codes with ghost rows = 0 (assembler.inl:174-181) and owns
    [BegInd,EndInd); the rank-local CSR equals the oracle's and the ranks tile the global matrix"""
    M = asm_oracle
    n = (4, 3, 3)
    co, te, cr = M.cube_mesh(*n, nranks=2)
    variables = [(gc.P2, 1)]
    dm = M.DofMap(te, variables, cr, 2, nnode=co.shape[0])
    _, forms0, rhsf0, prob = problems.c2_p2_aniso(pkg, M, co, te)
    for rank in range(2):
        rowcode, colcode = dm.codes(rank)
        sel = np.nonzero((rowcode != 0).any(axis=1))[0]  # cells touching an owned dof (halo recompute set)
        ctx.mesh_set(co, te[sel])
        ctx.dofmap_set(rowcode[sel], colcode[sel], int(dm.beg_ind[rank]), int(dm.end_ind[rank]), dm.nrows)
        nnz = ctx.pattern_build()
        rowptr, colind = ctx.pattern_get()
        rp, ci, v, r, st = M.assemble(prob, co, te, dm, rank=rank)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        xc = co[te[sel]].mean(axis=1)
        forms = [pkg.make_form(gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET, problems.sym_K(xc))]
        val, rhs = np.zeros(nnz), np.zeros(rp.size - 1)
        assert ctx.assemble(forms, rhsf0, val, rhs) == 0
        rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
        assert (np.abs(val - v) / np.repeat(rowmax, np.diff(rp))).max() <= RTOL
        assert np.abs(rhs - r).max() <= RTOL * np.abs(r).max()


def test_unstructured_mesh_and_unoriented_tets(pkg, ctx, asm_oracle):
    """generic input path (afb_mesh_set): jittered nodes, shuffled element order, P3 edge-pair orientation"""
    M = asm_oracle
    rng = np.random.default_rng(5)
    co, te, _ = M.cube_mesh(3, 3, 3)
    co = co + 0.05 * rng.standard_normal(co.shape) * (1.0 / 3)
    te = te[rng.permutation(te.shape[0])]
    # re-orient after the jitter exactly like the reference does per cell (ordering.inl:8-26)
    p = co[te]
    det = np.linalg.det(p[:, :3, :] - p[:, 3:4, :])
    te[det < 0] = te[det < 0][:, [0, 1, 3, 2]]
    ctx.mesh_set(co, te)
    ctx.mesh_orient()
    c2, t2 = ctx.mesh_get()
    assert np.array_equal(t2, te)
    variables = [(gc.P3, 1)]
    ctx.dofmap_natural(variables)
    dm = M.DofMap(te, variables, nnode=co.shape[0])
    row, col = ctx.dofmap_get()
    assert np.array_equal(col, dm.elem2dof + 1)
    K = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, te.shape[0], 14)
    _, forms, rhsf, prob = problems._mk(pkg, M, variables, [(0, 0, gc.GRAD, gc.GRAD, 4, gc.T_SYMMETRIC, gc.L_PER_TET, K, 1.0)],
                                        [(0, gc.IDEN, 3, gc.T_NULL, gc.L_CONST, None, 1.0)])
    _check(ctx, M, prob, forms, rhsf, co, te, dm, "unstructured P3")


def test_negative_sign_codes(pkg, ctx, asm_oracle, oracle):
    """oriented dofs: index codes sign*(id+1) with negative signs (assembler.inl:49-55, :155-159); the scatter multiplies every
    contribution by s_row * s_col and the load entry by s_row (assembler.inl:407, :416-418).  Signs are per (cell, local dof) like
    the reference's OrderTempl sign slot; same sign on the row and the column code of a local dof."""
    M, O = asm_oracle, oracle
    rng = np.random.default_rng(11)
    co, te, _ = M.cube_mesh(4, 3, 2)
    for variables, fem in (([(gc.P2, 1)], gc.P2), ([(gc.P1, 1)], gc.P1)):
        dm = M.DofMap(te, variables, nnode=co.shape[0])
        rowcode, colcode = dm.codes(None)
        sgn = np.where(rng.random(rowcode.shape) < 0.4, -1, 1).astype(np.int64)
        rowcode, colcode = rowcode * sgn, colcode * sgn
        K = problems.sym_K(co[te].mean(axis=1))
        _, forms, rhsf, prob = problems._mk(pkg, M, variables, [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K, 1.0)],
                                            [(0, gc.IDEN, 2, gc.T_SCALAR, gc.L_PER_TET, 1 + co[te].mean(axis=1)[:, :1].copy(), 1.0)])
        ctx.mesh_set(co, te)
        ctx.dofmap_set(rowcode, colcode, 0, dm.nrows, dm.nrows)
        nnz = ctx.pattern_build()
        rowptr, colind = ctx.pattern_get()
        rp, ci = M.template_pattern(rowcode, colcode, 0, dm.nrows)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        A, F = prob.element_matrices(co[te].transpose(1, 0, 2), idx=np.arange(te.shape[0]))
        v, r = np.zeros(ci.size), np.zeros(dm.nrows)
        assert O.scatter_csr(rowcode, colcode, A, F, 0, rp, ci, v, r) == 0
        val, rhs = np.full(nnz, np.nan), np.full(dm.nrows, np.nan)
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
        _compare(val, rhs, v, r, rp, te, nnz, "signed codes fem %d" % fem)
        # all signs positive gives a different matrix: the signs were honoured, not dropped
        v0 = np.zeros(ci.size)
        O.scatter_csr(np.abs(rowcode), np.abs(colcode), A, None, 0, rp, ci, v0, None)
        assert np.abs(v0 - v).max() > 1e-3 * np.abs(v).max()


def test_dofmap_rejects_out_of_range_codes(pkg, ctx, asm_oracle):
    """an index outside its interval aborts the reference (assembler.inl:399-412); here afb_dofmap_set fails with -7"""
    M = asm_oracle
    co, te, _ = M.cube_mesh(2, 2, 2)
    dm = M.DofMap(te, [(gc.P1, 1)], nnode=co.shape[0])
    rowcode, colcode = dm.codes(None)
    ctx.mesh_set(co, te)
    for which, bad in (("row", dm.nrows + 1), ("col", dm.nrows + 1), ("col", 0)):
        rc, cc = rowcode.copy(), colcode.copy()
        (rc if which == "row" else cc)[3, 1] = bad
        with pytest.raises(pkg.AfbError) as e:
            ctx.dofmap_set(rc, cc, 0, dm.nrows, dm.nrows)
        assert e.value.code == -7
    ctx.dofmap_set(rowcode, colcode, 0, dm.nrows, dm.nrows)   # the valid table is still accepted
