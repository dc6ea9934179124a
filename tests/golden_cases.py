"""Seeded element-level cases shared by tests/golden/make_golden.py (which records the reference's
outputs in tests/golden/ref_outputs.npz) and by the parity tests (which regenerate the same inputs).
Covers every operator / space / tensor family of SURVEY.md section 8a: IDEN|GRAD on P0..P3,
FemVec<3,.>, DIV, mixed trial/test spaces, NULL/SCALAR/SYMMETRIC/GENERAL tensors in CONST / PER_TET /
PER_POINT layouts, the rhs trick, low and high quadrature orders, ragged batch sizes."""
import numpy as np

IDEN, GRAD, DIV = 1, 2, 3
P0, P1, P2, P3 = 1, 2, 3, 4
T_NULL, T_SCALAR, T_SYMMETRIC, T_GENERAL = 1, 2, 3, 4
L_CONST, L_PER_TET, L_PER_POINT = 0, 1, 2

NPTS = [1, 1, 4, 8, 14, 14, 24, 35, 46, 59, 81, 110, 168, 172, 204, 264, 304, 364, 436, 487, 552]
NF = {P0: 1, P1: 4, P2: 10, P3: 20}


def op_dims(op, fem, vec):
    if op == IDEN:
        return vec * NF[fem], vec
    if op == GRAD:
        return vec * NF[fem], 3 * vec
    return 3 * NF[fem], 1


def random_tets(rng, f, flip=True):
    """well-shaped random tets (4, f, 3); half of them negatively oriented"""
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    XY = base[:, None, :] * rng.uniform(0.5, 2.0, (1, f, 1)) + 0.2 * rng.standard_normal((4, f, 3)) + rng.uniform(-3, 3, (1, f, 3))
    if flip:
        XY[2, ::2], XY[3, ::2] = XY[3, ::2].copy(), XY[2, ::2].copy()
    return np.ascontiguousarray(XY)


def tensor(rng, ttype, layout, idim, jdim, f, q):
    n = {L_CONST: 1, L_PER_TET: f, L_PER_POINT: f * q}[layout]
    if ttype == T_NULL:
        return None
    if ttype == T_SCALAR:
        return np.ascontiguousarray(rng.uniform(0.5, 2.0, (n, 1)))
    D = rng.standard_normal((n, idim, jdim))  # memory order [j][k]: K(k,j) at k + jdim*j
    if ttype == T_SYMMETRIC:
        D = D + D.transpose(0, 2, 1) + 2 * idim * np.eye(idim)
    return np.ascontiguousarray(D.reshape(n, idim * jdim))


def cases():
    rng = np.random.default_rng(20261017)
    spec = [
        # name, (opA, femA, vecA), (opB, femB, vecB), order, ttype, layout, f
        ("p1_stiff_sym_const", (GRAD, P1, 1), (GRAD, P1, 1), 2, T_SYMMETRIC, L_CONST, 7),
        ("p1_stiff_scalar_pt", (GRAD, P1, 1), (GRAD, P1, 1), 2, T_SCALAR, L_PER_POINT, 5),
        ("p1_mass_null", (IDEN, P1, 1), (IDEN, P1, 1), 2, T_NULL, L_CONST, 3),
        ("p2_stiff_sym_tet", (GRAD, P2, 1), (GRAD, P2, 1), 2, T_SYMMETRIC, L_PER_TET, 33),
        ("p2_stiff_gen_pt", (GRAD, P2, 1), (GRAD, P2, 1), 3, T_GENERAL, L_PER_POINT, 6),
        ("p2_mass_scalar_tet", (IDEN, P2, 1), (IDEN, P2, 1), 4, T_SCALAR, L_PER_TET, 9),
        ("p2_conv_gen", (GRAD, P2, 1), (IDEN, P2, 1), 3, T_GENERAL, L_CONST, 4),
        ("p3_stiff_scalar_pt_q14", (GRAD, P3, 1), (GRAD, P3, 1), 4, T_SCALAR, L_PER_POINT, 5),
        ("p3_mass_scalar_pt_q24", (IDEN, P3, 1), (IDEN, P3, 1), 6, T_SCALAR, L_PER_POINT, 5),
        ("p3_stiff_sym_q46", (GRAD, P3, 1), (GRAD, P3, 1), 8, T_SYMMETRIC, L_PER_TET, 3),
        ("p3_stiff_null_q552", (GRAD, P3, 1), (GRAD, P3, 1), 20, T_NULL, L_CONST, 2),
        ("p2vec_elast_const", (GRAD, P2, 3), (GRAD, P2, 3), 2, T_SYMMETRIC, L_CONST, 5),
        ("p2vec_elast_gen_pt", (GRAD, P2, 3), (GRAD, P2, 3), 2, T_GENERAL, L_PER_POINT, 3),
        ("p2vec_mass", (IDEN, P2, 3), (IDEN, P2, 3), 4, T_SCALAR, L_CONST, 3),
        ("p1vec_stiff_null", (GRAD, P1, 3), (GRAD, P1, 3), 5, T_NULL, L_CONST, 4),
        ("stokes_p_divv", (IDEN, P1, 1), (DIV, P2, 3), 2, T_SCALAR, L_CONST, 5),
        ("stokes_divu_q", (DIV, P2, 3), (IDEN, P1, 1), 2, T_NULL, L_CONST, 5),
        ("div_div_p3", (DIV, P3, 3), (DIV, P3, 3), 4, T_SCALAR, L_PER_TET, 2),
        ("gradp3_x_gradp1vec_gen", (GRAD, P3, 1), (GRAD, P1, 3), 5, T_GENERAL, L_PER_POINT, 3),
        ("gradp1_x_gradp2", (GRAD, P1, 1), (GRAD, P2, 1), 2, T_SYMMETRIC, L_CONST, 3),
        ("idenp2vec_x_gradp1", (IDEN, P2, 3), (GRAD, P1, 1), 3, T_GENERAL, L_PER_TET, 3),
        ("rhs_p1_scalar_pt", (IDEN, P0, 1), (IDEN, P1, 1), 2, T_SCALAR, L_PER_POINT, 6),
        ("rhs_p2_null", (IDEN, P0, 1), (IDEN, P2, 1), 2, T_NULL, L_CONST, 4),
        ("rhs_p3_scalar_tet", (IDEN, P0, 1), (IDEN, P3, 1), 3, T_SCALAR, L_PER_TET, 4),
        ("rhs_p2vec_vector_pt", (IDEN, P0, 1), (IDEN, P2, 3), 2, T_GENERAL, L_PER_POINT, 4),
        ("rhs_p2vec_scalar", (IDEN, P0, 1), (IDEN, P2, 3), 2, T_SCALAR, L_CONST, 2),
        ("p0_mass", (IDEN, P0, 1), (IDEN, P0, 1), 1, T_NULL, L_CONST, 3),
        ("gradp0_gradp1", (GRAD, P0, 1), (GRAD, P1, 1), 1, T_NULL, L_CONST, 2),
        ("order0_p1_mass", (IDEN, P1, 1), (IDEN, P1, 1), 0, T_NULL, L_CONST, 1),
    ]
    out = []
    for name, A, B, order, tt, lay, f in spec:
        q = NPTS[order]
        XY = random_tets(rng, f)
        idim, jdim = op_dims(*A)[1], op_dims(*B)[1]
        D = tensor(rng, tt, lay, idim, jdim, f, q)
        out.append((name, A + B + (order, tt, lay), XY, D))
    return out


NPTS_TRI = [1, 1, 3, 6, 6, 7, 12, 15, 16, 19, 25, 28, 33, 37, 42, 49, 55, 60, 67, 73, 79]


def face_cases():
    """seeded surface-integral cases (fem3Dface, fem/operations/int_face.inl): Neumann / Robin style forms, every face number"""
    rng = np.random.default_rng(20261018)
    spec = [
        # name, (opA, femA, vecA), (opB, femB, vecB), order, ttype, layout, f
        ("robin_p1_scalar_tet", (IDEN, P1, 1), (IDEN, P1, 1), 2, T_SCALAR, L_PER_TET, 9),
        ("robin_p2_null", (IDEN, P2, 1), (IDEN, P2, 1), 4, T_NULL, L_CONST, 8),
        ("robin_p2_scalar_pt", (IDEN, P2, 1), (IDEN, P2, 1), 5, T_SCALAR, L_PER_POINT, 6),
        ("robin_p3_q12", (IDEN, P3, 1), (IDEN, P3, 1), 6, T_SCALAR, L_CONST, 5),
        ("neumann_p1_rhs", (IDEN, P0, 1), (IDEN, P1, 1), 2, T_SCALAR, L_PER_POINT, 7),
        ("neumann_p2_rhs_null", (IDEN, P0, 1), (IDEN, P2, 1), 3, T_NULL, L_CONST, 5),
        ("neumann_p2vec_traction", (IDEN, P0, 1), (IDEN, P2, 3), 3, T_GENERAL, L_PER_TET, 6),
        ("robin_p2vec_sym", (IDEN, P2, 3), (IDEN, P2, 3), 4, T_SYMMETRIC, L_CONST, 4),
        ("gradp2_x_idenp1vec_gen_pt", (GRAD, P2, 1), (IDEN, P1, 3), 3, T_GENERAL, L_PER_POINT, 5),
        ("gradp1_x_gradp1_face", (GRAD, P1, 1), (GRAD, P1, 1), 1, T_SYMMETRIC, L_PER_TET, 4),
        ("idenp1_x_divp2vec", (IDEN, P1, 1), (DIV, P2, 3), 2, T_SCALAR, L_CONST, 4),
        ("robin_p2_q79", (IDEN, P2, 1), (IDEN, P2, 1), 20, T_NULL, L_CONST, 2),
    ]
    out = []
    for name, A, B, order, tt, lay, f in spec:
        q = NPTS_TRI[order]
        XY = random_tets(rng, f)
        face = (np.arange(f) + rng.integers(0, 4)) % 4
        idim, jdim = op_dims(*A)[1], op_dims(*B)[1]
        D = tensor(rng, tt, lay, idim, jdim, f, q)
        out.append((name, A + B + (order, tt, lay), XY, face.astype(np.int32), D))
    return out
