"""CPU tests (no GPU): the oracle restatement against the reference's own known-answer tables, the
committed reference outputs, and -- when it was built here -- the reference itself."""
import os

import numpy as np
import pytest

import golden_cases as gc
from conftest import ROOT

TET = np.array([[0, 0, 0], [2, 1, 1], [1, 2, 1], [2, 1, 2]], float)[:, None, :]  # (4,1,3)


def poly_tensor(XYG, nrow, ncol):
    """the polynomial GENERAL tensor of tests/fem/operations/int_tet_test.cpp:213-225 in user layout"""
    f, q, _ = XYG.shape
    D = np.zeros((f, q, ncol, nrow))  # memory [j][k] : K(k,j) at k + nrow*j
    for i in range(nrow):
        for j in range(ncol):
            D[:, :, j, i] = (i * XYG[:, :, 0] + j * XYG[:, :, 1] + (i % 2) * XYG[:, :, 2]) * (i * ncol + j)
    return np.ascontiguousarray(D.reshape(f * q, nrow * ncol))


def test_int_tet_table_general(oracle, ref_tests):
    """GRAD(P3) x GRAD(P1^3), 12x20 table /720 (int_tet_test.cpp:230-244), tol 10(1+|A|)eps"""
    g = ref_tests["int_tet"]["grad_p3_x_grad_p1vec_general"]
    exp = np.array(g["table_rows_test_cols_trial"]) / g["coef"]  # [ib][ia]
    XYG = oracle.quad_points(5, TET)
    D = poly_tensor(XYG, 9, 3)
    form = (gc.GRAD, gc.P3, 1, gc.GRAD, gc.P1, 3, 5, gc.T_GENERAL, gc.L_PER_POINT)
    A = oracle.fem3dtet(form, TET, D)[0]  # [ia][ib]
    assert np.linalg.norm(A.T - exp) <= 10 * (1 + np.linalg.norm(A)) * np.finfo(float).eps * 4


def test_int_tet_table_identity_all_tensor_kinds(oracle, ref_tests):
    """GRAD(P1^3)^2 identity, 12x12 table /1440 for the trait variants (int_tet_test.cpp:332-389)"""
    g = ref_tests["int_tet"]["grad_p1vec_sq_identity"]
    exp = np.array(g["table"]) / g["coef"]
    I9 = np.eye(9).reshape(1, 81)
    for tt, lay, D in [(gc.T_GENERAL, gc.L_CONST, I9), (gc.T_SYMMETRIC, gc.L_CONST, I9), (gc.T_SCALAR, gc.L_CONST, np.ones((1, 1))),
                       (gc.T_NULL, gc.L_CONST, None), (gc.T_GENERAL, gc.L_PER_TET, I9), (gc.T_SCALAR, gc.L_PER_POINT, np.ones((14, 1))),
                       (gc.T_GENERAL, gc.L_PER_POINT, np.repeat(I9, 14, axis=0))]:
        form = (gc.GRAD, gc.P1, 3, gc.GRAD, gc.P1, 3, 5, tt, lay)
        A = oracle.fem3dtet(form, TET, D)[0]
        assert np.linalg.norm(A.T - exp) <= 10 * (1 + np.linalg.norm(A)) * np.finfo(float).eps


def test_rhs_trick(oracle, ref_tests):
    """int (mu,mu,mu).phi_i, OpA = IDEN(P0), P2^3: {-1 x4, 4 x6} x3 * (mu |T| / 20) pattern (int_tet_test.cpp:448-501)"""
    g = ref_tests["int_tet"]["rhs_p0_x_iden_p2vec"]
    B = np.array(g["table"])
    mu = g["mu"]
    for tt, D, scale in [(gc.T_SCALAR, np.full((1, 1), mu), 1.0), (gc.T_NULL, None, 1.0 / mu)]:
        form = (gc.IDEN, gc.P0, 1, gc.IDEN, gc.P2, 3, 2, tt, gc.L_CONST)
        A = oracle.fem3dtet(form, TET, D)[0].ravel()
        assert np.linalg.norm(A - scale * B) <= 10 * (1 + np.linalg.norm(A)) * np.finfo(float).eps


def test_space_tables(oracle, ref_tests):
    """U tables of IDEN/GRAD on P0..P3 at the 4-point rule on two fused tets (predefined_spaces_test.cpp:62-382)"""
    s = ref_tests["spaces"]
    XYL = np.array(s["XYL"])
    XYZ = np.array(s["XYZ"]).reshape(4, 2, 3)  # [l][r][k]
    fem = {"FEM_P0": gc.P0, "FEM_P1": gc.P1, "FEM_P2": gc.P2, "FEM_P3": gc.P3}
    for name, t in s["tables"].items():
        op = gc.IDEN if name.startswith("IDEN") else gc.GRAD
        U = oracle.operator_apply(op, fem[name.split("_", 1)[1]], 1, XYL, XYZ)
        exp = np.array(t["U"]).reshape(2, t["nfa"], 4, t["dim"])
        assert np.linalg.norm(U - exp) <= 100 * (1 + np.linalg.norm(exp)) * np.finfo(float).eps * 50, name


def test_quadrature_exactness(oracle):
    """every rule integrates x + 10 y^(ord-1) + 1000 z^ord exactly on the unit tet, point counts
    (quadrature_formulas_test.cpp:7-45)"""
    from math import factorial
    for order in range(1, 21):
        p, w = oracle.tet_quadrature(order)
        assert p.shape[0] == gc.NPTS[order]
        assert abs(w.sum() - 1) < 1e-14 and (p > 0).all()
        x, y, z = p[:, 1], p[:, 2], p[:, 3]
        num = (w * (x + 10 * y ** (order - 1) + 1000 * z ** order)).sum() / 6
        mono = lambda k: factorial(k) / factorial(k + 3)  # int_T z^k
        exact = mono(1) + 10 * mono(order - 1) + 1000 * mono(order)
        assert abs(num - exact) <= max(1e-10, 1e-15 * abs(exact))


def test_oracle_vs_committed_reference_outputs(oracle, ref_outputs):
    for name, form, XY, D in gc.cases():
        A = oracle.fem3dtet(form, XY, D)
        ref = ref_outputs[name]
        assert np.abs(A - ref).max() <= 1e-13 * (1 + np.abs(ref).max()), name


def test_oracle_vs_live_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    for name, form, XY, D in gc.cases():
        a = oracle.fem3dtet(form, XY, D)
        b = oracle.fem3dtet(form, XY, D, impl="ref", mode=1, fuse=3, nthreads=2)
        assert np.abs(a - b).max() <= 1e-13 * (1 + np.abs(b).max()), name


FACE_TET = np.array([[1, 1, 1], [2, 1, 1], [1, 2, 1], [1, 1, 2]], float)[:, None, :]  # int_face_test.cpp:16-19


def face_normal_tensor(tri_quadrature, face=1):
    """the tensor of tests/fem/operations/int_face_test.cpp:29-56 at the points of the order-3 triangle rule lifted to `face`:
    D(i,j) = sum_k (x[j%3] + (3i+k)/10) n_k, n = outward normal; user layout K(k,j) at k + 3*j"""
    P = FACE_TET[:, 0, :]
    p, _ = tri_quadrature(3)
    idx = [(face + k) % 4 for k in range(3)]
    X = p @ P[idx]                                    # (q,3) physical points on the face
    n = np.cross(P[idx[1]] - P[idx[0]], P[idx[2]] - P[idx[0]])
    if n @ (P[(face + 3) % 4] - P[idx[0]]) > 0:
        n = -n
    n /= np.linalg.norm(n)
    q = X.shape[0]
    D = np.zeros((q, 3, 3))                           # memory [j][i]
    for i in range(3):
        for j in range(3):
            D[:, j, i] = sum((X[:, j % 3] + (3 * i + k) / 10.0) * n[k] for k in range(3))
    return np.ascontiguousarray(D.reshape(q, 9))


def test_int_face_table(oracle, ref_tests):
    """fem3Dface GRAD(P2) x IDEN(P1^3) over face 1, 12x10 known-answer table (int_face_test.cpp:78-95), tol 100(1+|A|)eps"""
    g = ref_tests["int_face"]["grad_p2_x_iden_p1vec_face1"]
    exp = np.array(g["table_cols_trial_rows_test"])  # [ia][ib]
    D = face_normal_tensor(oracle.tri_quadrature, g["face"])
    form = (gc.GRAD, gc.P2, 1, gc.IDEN, gc.P1, 3, g["order"], gc.T_GENERAL, gc.L_PER_POINT)
    A = oracle.fem3dface(form, FACE_TET, [g["face"]], D)[0]
    assert np.linalg.norm(A - exp) <= 100 * (1 + np.linalg.norm(exp)) * np.finfo(float).eps


def test_triangle_rules(oracle):
    """weights sum to 1, interior points, exactness on monomials of the stated order (int_T x^a y^b = a! b! / (a+b+2)! on the unit triangle)"""
    from math import factorial
    for order in range(1, 21):
        p, w = oracle.tri_quadrature(order)
        assert p.shape[0] == gc.NPTS_TRI[order] and abs(w.sum() - 1) < 1e-14 and (p > 0).all()
        assert np.abs(p.sum(axis=1) - 1).max() < 1e-15
        x, y = p[:, 0], p[:, 1]
        for a in range(order + 1):
            b = order - a
            num = (w * x ** a * y ** b).sum() / 2
            assert abs(num - factorial(a) * factorial(b) / factorial(a + b + 2)) <= 1e-14, (order, a)


def test_face_oracle_vs_committed_reference_outputs(oracle, ref_face_outputs):
    for name, form, XY, face, D in gc.face_cases():
        A = oracle.fem3dface(form, XY, face, D)
        ref = ref_face_outputs[name]
        assert np.abs(A - ref).max() <= 1e-13 * (1 + np.abs(ref).max()), name


def test_face_oracle_vs_live_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    for name, form, XY, face, D in gc.face_cases():
        a = oracle.fem3dface(form, XY, face, D)
        b = oracle.fem3dface(form, XY, face, D, impl="ref")
        assert np.abs(a - b).max() <= 1e-13 * (1 + np.abs(b).max()), name


def test_apply_dir_reference_case(asm_oracle):
    """applyDir(A, F, dof, bc) exactly as tests/fem/operations/dc_on_dof_test.cpp:13-41 builds its expectation:
    A(i,j) = (i+1)(j+1), F(i) = i+1, dof 2, bc 1 -> row and column 2 zeroed, A(2,2) = 1, F -= A(:,2) bc, F(2) = bc"""
    N, dof, bc = 4, 2, 1.0
    A = np.array([[(i + 1) * (j + 1) for i in range(N)] for j in range(N)], float)[None]   # [e, col j, row i]
    F = np.arange(1, N + 1, dtype=float)[None]
    A_exp, F_exp = A.copy(), F.copy()
    A_exp[0, dof, :] = 0; A_exp[0, :, dof] = 0; A_exp[0, dof, dof] = 1
    F_exp[0] -= A[0, dof, :] * bc
    F_exp[0, dof] = bc
    flag = np.zeros(N, dtype=np.uint8); flag[dof] = 1
    value = np.zeros(N); value[dof] = bc
    colcode = np.arange(1, N + 1)[None]
    asm_oracle.apply_dir(A, F, colcode, flag, value)
    assert np.array_equal(A, A_exp) and np.array_equal(F, F_exp)


def test_identity_tensor_incompatible_dims(oracle):
    """TENSOR_NULL with Dim(OpA) != Dim(OpB) is an error (diff_tensor.h:315-317)"""
    form = (gc.GRAD, gc.P1, 1, gc.IDEN, gc.P1, 1, 2, gc.T_NULL, gc.L_CONST)
    with pytest.raises(RuntimeError):
        oracle.fem3dtet(form, TET, None)


# ---- assembler level: oracle/asm_oracle.py pinned to the reference's OWN Assembler -------------------------------------------
# (unmodified anifem++/inmost_interface/{assembler.inl, global_enumerator.cpp, ordering.inl, elemental_assembler.cpp} compiled on
# oracle/mock_inmost/inmost.h by `make -C oracle refasm`; its outputs for the cases of tests/asm_cases.py are committed as
# tests/golden/ref_assembler.npz by tests/golden/make_golden_asm.py)
import asm_cases  # noqa: E402


@pytest.fixture(scope="module")
def ref_asm_golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_assembler.npz")))


def _multi_dof_entities(M, variables):
    nd = np.zeros(4, dtype=int)
    for fem, vec in variables:
        nd += np.array(M.NDOF[fem]) * vec
    return (nd > 1).any()


@pytest.mark.parametrize("live", [False, True])
def test_numbering_vs_reference_enumerators(asm_oracle, oracle, ref_asm_golden, live):
    """all six GlobEnumeration types (global_enumerator.cpp:562-605, 671-777, 818-835) + fill_assemble_templates
    (assembler.inl:139-184): index codes bit-exact on a cube and on a mesh with scrambled node ids / element order.
    MINIBLOCKS: the reference's forward formula (:826) is not a bijection once an entity carries more than one dof (two strides are
    mixed) -- asserted here; the oracle and the product follow the layout its inverse map decodes (:866-871) in that case."""
    M = asm_oracle
    if live and not oracle.have_refasm():
        pytest.skip("oracle/_ref/libanifem_refasm.so not built (reference tree absent)")
    for name, co, te, variables in asm_cases.numbering_cases(M):
        for et in M.ENUM_TYPES:
            if live:
                R = oracle.RefAssembler(co, te, variables, et)
                codes, nrows = R.codesC, R.nrows
            else:
                codes, nrows = ref_asm_golden["num/%s/%s/codes" % (name, et)], int(ref_asm_golden["num/%s/%s/nrows" % (name, et)][0])
            exp, n = M.enumerate_dofs(te, variables, et, co.shape[0])
            assert n == nrows, (name, et)
            ref = np.abs(codes) - 1
            if et == "MINIBLOCKS" and _multi_dof_entities(M, variables):
                assert np.unique(ref).size < nrows, "the reference's MINIBLOCKS forward map became a bijection: re-pin"
                continue
            assert np.array_equal(ref, exp), (name, et)
            assert (codes > 0).all()   # Lagrange spaces: no oriented dofs, all signs +
            assert np.array_equal(M.DofMap(te, variables, nnode=co.shape[0], enum_type=et).elem2dof, exp)


@pytest.mark.parametrize("live", [False, True])
def test_assembly_vs_reference_assembler(asm_oracle, oracle, ref_asm_golden, live):
    """AssembleTemplate (assembler.inl:589-695): pattern bit-exact.  Assemble (:313-488) in its default mode (unsorted rows,
    find-or-append, drop_val decides what enters the pattern) and in the is_mtx_include_template + use_ordered_insert mode
    (:428-438): values / rhs within 1e-13 of the row scale; the default-mode pattern is the subset of the template whose
    omitted entries are exactly the ones every cell dropped."""
    M = asm_oracle
    if live and not oracle.have_refasm():
        pytest.skip("oracle/_ref/libanifem_refasm.so not built (reference tree absent)")
    for name, co, te, variables, prob, kw in asm_cases.assembly_cases(M, oracle):
        dm = M.DofMap(te, variables, nnode=co.shape[0])
        drop = kw.get("drop_val", 1e-100)
        rp, ci, v, r, st = M.assemble(prob, co, te, dm, drop_val=drop)
        nrows = rp.size - 1
        key = np.repeat(np.arange(nrows), np.diff(rp)) * nrows + ci
        R = oracle.RefAssembler(co, te, variables, "NATURAL") if live else None
        trp, tci = R.template() if live else (ref_asm_golden["asm/%s/template_rowptr" % name], ref_asm_golden["asm/%s/template_colind" % name])
        assert np.array_equal(trp, rp) and np.array_equal(tci, ci), name + ": template pattern"
        for mode, opts in (("plain", {}), ("templ", dict(include_template=True, ordered_insert=True))):
            if live:
                st2, rp2, ci2, v2, r2 = R.assemble(prob, drop_val=drop, **opts)
            else:
                g = lambda k: ref_asm_golden["asm/%s/%s/%s" % (name, mode, k)]
                st2, rp2, ci2, v2, r2 = int(g("status")[0]), g("rowptr"), g("colind"), g("val"), g("rhs")
            assert st2 == st == 0
            key2 = np.repeat(np.arange(nrows), np.diff(rp2)) * nrows + ci2
            assert np.isin(key2, key).all(), name + ": reference entry outside the template"
            if mode == "templ":
                assert np.array_equal(key2, key)
            pos = np.searchsorted(key, key2)
            rowmax = np.repeat(np.maximum.reduceat(np.abs(v), rp[:-1]), np.diff(rp))
            assert (np.abs(v[pos] - v2) <= 1e-13 * rowmax[pos]).all(), name
            miss = np.ones(key.size, bool)
            miss[pos] = False
            assert (v[miss] == 0).all(), name + ": an entry the reference never inserted is non-zero in the oracle"
            assert np.abs(r - r2).max() <= 1e-13 * max(np.abs(r).max(), 1e-300)


def test_reference_assembler_status_codes(asm_oracle, oracle):
    """non-finite local value -> -1 (assembler.inl:419-424, 475-479), same as the oracle"""
    if not oracle.have_refasm():
        pytest.skip("oracle/_ref/libanifem_refasm.so not built (reference tree absent)")
    M = asm_oracle
    co, te, _ = M.cube_mesh(2, 2, 1)
    variables = [(gc.P1, 1)]
    K = np.ones((te.shape[0], 1))
    K[3, 0] = np.nan
    prob = M.Problem(variables, [dict(trial=0, test=0, opA=gc.GRAD, opB=gc.GRAD, order=2, ttype=gc.T_SCALAR, layout=gc.L_PER_TET, D=K)], [])
    R = oracle.RefAssembler(co, te, variables, "NATURAL")
    assert R.assemble(prob)[0] == -1
    assert M.assemble(prob, co, te, M.DofMap(te, variables, nnode=co.shape[0]))[4] == -1
