"""GPU parity tests of the element path: afb_fem3dtet_batched (C ABI) against the CPU oracle, the
committed reference outputs and the reference's own known-answer tables.  FP64 tolerance of
north_star: 1e-12 relative (to the largest entry of the element matrix)."""
import numpy as np
import pytest

import golden_cases as gc
from test_oracle_golden import FACE_TET, TET, face_normal_tensor, poly_tensor

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _form(pkg, form, D):
    opA, femA, vecA, opB, femB, vecB, order, tt, lay = form
    return pkg.make_form(opA, femA, vecA, opB, femB, vecB, order, tt, lay, D)


def test_seeded_cases_vs_oracle_and_reference(pkg, ctx, oracle, ref_outputs):
    worst = 0.0
    for name, form, XY, D in gc.cases():
        A = ctx.fem3dtet(_form(pkg, form, D), XY)
        ref = ref_outputs[name]
        orc = oracle.fem3dtet(form, XY, D)
        scale = np.abs(ref).reshape(ref.shape[0], -1).max(axis=1)[:, None, None] + 1e-300
        err = max((np.abs(A - ref) / scale).max(), (np.abs(A - orc) / scale).max())
        worst = max(worst, err)
        assert err <= RTOL, (name, err)
    print("worst relative error over seeded cases: %.2e" % worst)


def test_reference_known_answer_tables(pkg, ctx, ref_tests):
    eps = np.finfo(float).eps
    g = ref_tests["int_tet"]["grad_p3_x_grad_p1vec_general"]
    exp = np.array(g["table_rows_test_cols_trial"]) / g["coef"]
    XYG = ctx.quad_points(5, TET)
    D = poly_tensor(XYG, 9, 3)
    A = ctx.fem3dtet(pkg.make_form(gc.GRAD, gc.P3, 1, gc.GRAD, gc.P1, 3, 5, gc.T_GENERAL, gc.L_PER_POINT, D), TET)[0]
    assert np.linalg.norm(A.T - exp) <= 100 * (1 + np.linalg.norm(A)) * eps
    g = ref_tests["int_tet"]["grad_p1vec_sq_identity"]
    exp = np.array(g["table"]) / g["coef"]
    I9 = np.eye(9).reshape(1, 81)
    for tt, lay, D in [(gc.T_GENERAL, gc.L_CONST, I9), (gc.T_SYMMETRIC, gc.L_CONST, I9), (gc.T_SCALAR, gc.L_CONST, np.ones((1, 1))),
                       (gc.T_NULL, gc.L_CONST, None), (gc.T_GENERAL, gc.L_PER_TET, I9), (gc.T_SCALAR, gc.L_PER_POINT, np.ones((14, 1)))]:
        A = ctx.fem3dtet(pkg.make_form(gc.GRAD, gc.P1, 3, gc.GRAD, gc.P1, 3, 5, tt, lay, D), TET)[0]
        assert np.linalg.norm(A.T - exp) <= 100 * (1 + np.linalg.norm(A)) * eps
    g = ref_tests["int_tet"]["rhs_p0_x_iden_p2vec"]
    B = np.array(g["table"])
    A = ctx.fem3dtet(pkg.make_form(gc.IDEN, gc.P0, 1, gc.IDEN, gc.P2, 3, 2, gc.T_SCALAR, gc.L_CONST, np.full((1, 1), g["mu"])), TET)[0].ravel()
    assert np.linalg.norm(A - B) <= 100 * (1 + np.linalg.norm(A)) * eps


def test_face_seeded_cases_vs_oracle_and_reference(pkg, ctx, oracle, ref_face_outputs):
    """afb_fem3dface_batched (= Ani::fem3Dface) against the committed outputs of the reference and the oracle"""
    worst = 0.0
    for name, form, XY, face, D in gc.face_cases():
        A = ctx.fem3dface(_form(pkg, form, D), XY, face)
        ref = ref_face_outputs[name]
        orc = oracle.fem3dface(form, XY, face, D)
        scale = np.abs(ref).reshape(ref.shape[0], -1).max(axis=1)[:, None, None] + 1e-300
        err = max((np.abs(A - ref) / scale).max(), (np.abs(A - orc) / scale).max())
        worst = max(worst, err)
        assert err <= RTOL, (name, err)
    print("worst relative error over seeded face cases: %.2e" % worst)


def test_face_known_answer_table_and_properties(pkg, ctx, ref_tests):
    """the reference's own table for fem3Dface (int_face_test.cpp:78-95); triangle rules identical to the reference's; the face
    mass matrix of P2 sums to the face area on every face of 5000 random tets; wrong face index is an error (int_face.inl:28)"""
    g = ref_tests["int_face"]["grad_p2_x_iden_p1vec_face1"]
    exp = np.array(g["table_cols_trial_rows_test"])
    D = face_normal_tensor(pkg.tri_quadrature, g["face"])
    A = ctx.fem3dface(pkg.make_form(gc.GRAD, gc.P2, 1, gc.IDEN, gc.P1, 3, g["order"], gc.T_GENERAL, gc.L_PER_POINT, D), FACE_TET, [g["face"]])[0]
    assert np.linalg.norm(A - exp) <= 100 * (1 + np.linalg.norm(exp)) * np.finfo(float).eps
    rng = np.random.default_rng(7)
    f = 5000
    XY = gc.random_tets(rng, f)
    face = rng.integers(0, 4, f).astype(np.int32)
    M = ctx.fem3dface(pkg.make_form(gc.IDEN, gc.P2, 1, gc.IDEN, gc.P2, 1, 4, gc.T_NULL, gc.L_CONST), XY, face)
    idx = np.stack([(face + k) % 4 for k in range(3)], 0)
    P = XY[idx, np.arange(f)]                        # (3, f, 3)
    area = 0.5 * np.linalg.norm(np.cross(P[1] - P[0], P[2] - P[0]), axis=1)
    assert np.abs(M.sum(axis=(1, 2)) - area).max() <= 1e-12 * area.max()
    # dofs not on the face do not couple: the vertex opposite to the face and the three edges through it
    opp = (face + 3) % 4
    assert np.abs(M[np.arange(f), opp, :]).max() <= 1e-15 * area.max()
    with pytest.raises(pkg.AfbError) as e:
        ctx.fem3dface(pkg.make_form(gc.IDEN, gc.P1, 1, gc.IDEN, gc.P1, 1, 2, gc.T_NULL, gc.L_CONST), FACE_TET, [4])
    assert e.value.code == -7 and "face" in str(e.value)


def test_quad_points(pkg, ctx, oracle):
    rng = np.random.default_rng(1)
    XY = gc.random_tets(rng, 11)
    for order in (1, 2, 6, 13):
        a = ctx.quad_points(order, XY)
        b = oracle.quad_points(order, XY)
        assert np.abs(a - b).max() <= 1e-14 * (1 + np.abs(b).max())


def test_large_batch_and_ragged_sizes(pkg, ctx, oracle):
    rng = np.random.default_rng(2)
    for f in (1, 31, 32, 33, 1000, 4099):
        XY = gc.random_tets(rng, f)
        D = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, f, 4)
        form = (gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET)
        A = ctx.fem3dtet(_form(pkg, form, D), XY)
        B = oracle.fem3dtet(form, XY, D)
        scale = np.abs(B).reshape(f, -1).max(axis=1)[:, None, None]
        assert (np.abs(A - B) / scale).max() <= RTOL
    # empty batch is a no-op (int_tet.inl:7)
    assert ctx.fem3dtet(_form(pkg, form, D), np.zeros((4, 0, 3))).shape == (0, 10, 10)


def test_linearity_and_symmetry_properties(pkg, ctx):
    """size-independent properties at a large batch: A(K1+K2) = A(K1)+A(K2); symmetric K -> symmetric A;
    rows of a stiffness matrix sum to zero; mass matrix sums to |T|"""
    rng = np.random.default_rng(3)
    f = 20000
    XY = gc.random_tets(rng, f)
    K1 = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, f, 4)
    K2 = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, f, 4)
    mk = lambda K: pkg.make_form(gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K)
    A1, A2, A12 = ctx.fem3dtet(mk(K1), XY), ctx.fem3dtet(mk(K2), XY), ctx.fem3dtet(mk(K1 + K2), XY)
    s = np.abs(A12).max()
    assert np.abs(A12 - A1 - A2).max() <= 1e-12 * s
    assert np.abs(A1 - A1.transpose(0, 2, 1)).max() <= 1e-12 * s
    assert np.abs(A1.sum(axis=1)).max() <= 1e-11 * s
    M = ctx.fem3dtet(pkg.make_form(gc.IDEN, gc.P3, 1, gc.IDEN, gc.P3, 1, 6, gc.T_NULL, gc.L_CONST), XY)
    e1, e2, e3 = XY[1] - XY[0], XY[2] - XY[0], XY[3] - XY[0]
    vol = np.abs(np.einsum("ij,ij->i", e1, np.cross(e2, e3))) / 6
    assert np.abs(M.sum(axis=(1, 2)) - vol).max() <= 1e-12 * vol.max()


def test_error_codes(pkg, ctx):
    """error behaviour of the reference: incompatible identity tensor -> runtime_error (diff_tensor.h:315-317);
    unsupported space; bad quadrature order (quadrature_formulas.cpp:1498-1499)"""
    with pytest.raises(pkg.AfbError) as e:
        ctx.fem3dtet(pkg.make_form(gc.GRAD, gc.P1, 1, gc.IDEN, gc.P1, 1, 2, gc.T_NULL, gc.L_CONST), TET)
    assert e.value.code == -5 and "Identity tensor" in str(e.value)
    with pytest.raises(pkg.AfbError) as e:
        ctx.fem3dtet(pkg.make_form(gc.IDEN, 21, 1, gc.IDEN, gc.P1, 1, 2, gc.T_NULL, gc.L_CONST), TET)
    assert e.value.code == -3
    with pytest.raises(pkg.AfbError) as e:
        ctx.fem3dtet(pkg.make_form(gc.IDEN, gc.P1, 1, gc.IDEN, gc.P1, 1, 21, gc.T_NULL, gc.L_CONST), TET)
    assert e.value.code == -7
