"""GPU parity tests of FE-function evaluation (afb_fem3dapply_batched / afb_eval_quadrature = Ani::fem3DapplyL,
fem/operations/eval.h:13-120, core.inl:369-404) against the oracle's operator tables (which are pinned by the reference's
predefined_spaces_test tables and by the reference build oracle/_ref), and of the nonlinear-coefficient loop it enables."""
import numpy as np
import pytest

import golden_cases as gc
import problems

pytestmark = pytest.mark.gpu

OPS = [(gc.IDEN, gc.P1, 1), (gc.GRAD, gc.P1, 1), (gc.IDEN, gc.P2, 1), (gc.GRAD, gc.P2, 1), (gc.IDEN, gc.P3, 1), (gc.GRAD, gc.P3, 1),
       (gc.IDEN, gc.P2, 3), (gc.GRAD, gc.P2, 3), (gc.DIV, gc.P2, 3), (gc.GRAD, gc.P1, 3), (gc.DIV, gc.P3, 3), (gc.IDEN, gc.P0, 1)]


@pytest.mark.parametrize("op,fem,vec", OPS)
def test_fem3dapply_batched(pkg, ctx, oracle, op, fem, vec):
    rng = np.random.default_rng(100 * op + 10 * fem + vec)
    f = 37
    XY = gc.random_tets(rng, f)
    nfa, dim = gc.op_dims(op, fem, vec)
    dofs = rng.standard_normal((f, nfa))
    for XYL in (pkg.tet_quadrature(4)[0], rng.dirichlet(np.ones(4), size=5)):
        got = ctx.fem3dapply(op, fem, vec, XYL, XY, dofs)
        U = oracle.operator_apply(op, fem, vec, XYL, XY)               # (f, nfa, q, dim)
        exp = np.einsum("rinK,ri->rnK", U, dofs)
        scale = np.abs(exp).max() + 1e-300
        assert np.abs(got - exp).max() <= 1e-12 * scale, (op, fem, vec)
        if oracle.have_ref():
            Ur = oracle.operator_apply(op, fem, vec, XYL, XY, impl="ref")
            assert np.abs(got - np.einsum("rinK,ri->rnK", Ur, dofs)).max() <= 1e-12 * scale


def test_eval_quadrature_on_mesh_taylor_hood(pkg, ctx, asm_oracle, oracle):
    M = asm_oracle
    variables = [(gc.P2, 3), (gc.P1, 1)]
    co, te, _ = M.cube_mesh(3, 2, 3)
    ctx.mesh_cube(3, 2, 3)
    ctx.dofmap_natural(variables)
    dm = M.DofMap(te, variables, nnode=co.shape[0])
    rng = np.random.default_rng(1)
    u = rng.standard_normal(dm.nrows)
    XY = co[te].transpose(1, 0, 2)
    for (op, fem, vec, off, order) in [(gc.GRAD, gc.P2, 3, 0, 2), (gc.DIV, gc.P2, 3, 0, 3), (gc.IDEN, gc.P1, 1, 30, 2), (gc.GRAD, gc.P1, 1, 30, 4)]:
        nfa, dim = gc.op_dims(op, fem, vec)
        got = ctx.eval_quadrature(op, fem, vec, off, order, u)
        XYL = pkg.tet_quadrature(order)[0]
        U = oracle.operator_apply(op, fem, vec, XYL, XY)
        dofs = u[dm.elem2dof[:, off:off + nfa]]
        exp = np.einsum("rinK,ri->rnK", U, dofs)
        assert got.shape == exp.shape
        assert np.abs(got - exp).max() <= 1e-12 * (np.abs(exp).max() + 1e-300)


def test_nonlinear_coefficient_loop_on_device(pkg, ctx, asm_oracle, oracle):
    """one Picard step of -div((1 + u^2) grad u) = 1: the coefficient is evaluated from u_h on the GPU (device buffers end to
    end) and fed to the assembly as a PER_POINT scalar; compared with the oracle assembling the same coefficient"""
    import torch
    M = asm_oracle
    variables = [(gc.P2, 1)]
    n = (3, 3, 3)
    co, te, _ = M.cube_mesh(*n)
    ctx.mesh_cube(*n)
    ctx.dofmap_natural(variables)
    dm = M.DofMap(te, variables, nnode=co.shape[0])
    nnz = ctx.pattern_build()
    rng = np.random.default_rng(2)
    u = rng.standard_normal(dm.nrows)
    u_d = torch.from_numpy(u).cuda()
    uq = ctx.eval_quadrature(gc.IDEN, gc.P2, 1, 0, 4, u_d)          # (ntet, 14, 1) on the device
    K_d = (1.0 + uq * uq).reshape(te.shape[0], -1).contiguous()
    forms = [pkg.make_form(gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 4, gc.T_SCALAR, gc.L_PER_POINT, K_d)]
    rhsf = [pkg.make_form(gc.IDEN, gc.P0, 1, gc.IDEN, gc.P2, 1, 2, gc.T_NULL, gc.L_CONST)]
    val = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    rhs = torch.zeros(dm.nrows, dtype=torch.float64, device="cuda")
    assert ctx.assemble(forms, rhsf, val, rhs) == 0
    # oracle: same coefficient from the oracle's own tables
    XY = co[te].transpose(1, 0, 2)
    U = oracle.operator_apply(gc.IDEN, gc.P2, 1, pkg.tet_quadrature(4)[0], XY)
    uq_o = np.einsum("rinK,ri->rnK", U, u[dm.elem2dof])[..., 0]
    K_o = np.ascontiguousarray(1.0 + uq_o * uq_o)
    _, _, _, prob = problems._mk(pkg, M, variables, [(0, 0, gc.GRAD, gc.GRAD, 4, gc.T_SCALAR, gc.L_PER_POINT, K_o, 1.0)],
                                 [(0, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)])
    rp, ci, v, r, st = M.assemble(prob, co, te, dm)
    rowmax = np.repeat(np.maximum.reduceat(np.abs(v), rp[:-1]), np.diff(rp))
    assert (np.abs(val.cpu().numpy() - v) / rowmax).max() <= 1e-12
    assert np.abs(rhs.cpu().numpy() - r).max() <= 1e-12 * np.abs(r).max()
