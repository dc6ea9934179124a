"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle (oracle/liborc.so, C restatement) and,
when it was built, to the reference itself (oracle/_ref/libanifem_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (inmost-fem_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# enums shared with include/anifem_b200.h (values = the reference's Ani::OperatorType /
# Ani::FiniteElement / Ani::TensorType, anifem++/fem/operators.h:24-44, diff_tensor.h:17-22)
IDEN, GRAD, DIV = 1, 2, 3
P0, P1, P2, P3 = 1, 2, 3, 4
T_NULL, T_SCALAR, T_SYMMETRIC, T_GENERAL = 1, 2, 3, 4
L_CONST, L_PER_TET, L_PER_POINT = 0, 1, 2

_dp = ctypes.POINTER(ctypes.c_double)
_lp = ctypes.POINTER(ctypes.c_long)
_ip = ctypes.POINTER(ctypes.c_int)


def _P(a):
    return None if a is None else a.ctypes.data_as(_dp)


class _Form(ctypes.Structure):
    _fields_ = [("opA", ctypes.c_int), ("femA", ctypes.c_int), ("vecA", ctypes.c_int),
                ("opB", ctypes.c_int), ("femB", ctypes.c_int), ("vecB", ctypes.c_int),
                ("quad_order", ctypes.c_int), ("tensor_type", ctypes.c_int), ("tensor_layout", ctypes.c_int),
                ("D", _dp)]


def build(ref=True):
    """(Re)build liborc.so and, if the reference tree exists, _ref/libanifem_ref.so."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []))


_orc = None
_ref = None


def orc():
    global _orc
    if _orc is None:
        path = os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        L.orc_fem3dtet.restype = ctypes.c_int
        L.orc_fem3dtet.argtypes = [ctypes.POINTER(_Form), ctypes.c_long, _dp, _dp, _dp, _dp, _dp]
        L.orc_fem3dface.restype = ctypes.c_int
        L.orc_fem3dface.argtypes = [ctypes.POINTER(_Form), ctypes.c_long, _ip, _dp, _dp, _dp, _dp, _dp]
        L.orc_tri_quadrature.argtypes = [ctypes.c_int, ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
        L.orc_op_dims.argtypes = [ctypes.c_int] * 3 + [_ip, _ip]
        L.orc_quad_points.argtypes = [ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, _dp, _dp]
        L.orc_operator_apply.argtypes = [ctypes.c_int] * 4 + [_dp, ctypes.c_long, _dp, _dp, _dp, _dp, _dp]
        L.orc_tet_quadrature.argtypes = [ctypes.c_int, ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
        L.orc_scatter_csr.restype = ctypes.c_int
        L.orc_scatter_csr.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, _lp, _lp, _dp, _dp,
                                      ctypes.c_long, _lp, _ip, _dp, _dp, ctypes.c_double]
        _orc = L
    return _orc


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libanifem_ref.so"))


def ref():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(os.path.join(HERE, "_ref", "libanifem_ref.so"))
        L.ref_tet_quadrature.restype = ctypes.c_int
        L.ref_tet_quadrature.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int]
        L.ref_fem3dtet.restype = ctypes.c_int
        L.ref_fem3dtet.argtypes = [ctypes.c_int] * 9 + [_dp, ctypes.c_long] + [_dp] * 5 + [ctypes.c_int] * 3
        if hasattr(L, "ref_fem3dface"):
            L.ref_fem3dface.restype = ctypes.c_int
            L.ref_fem3dface.argtypes = [ctypes.c_int] * 9 + [_dp, ctypes.c_long, _ip] + [_dp] * 5
            L.ref_tri_quadrature.restype = ctypes.c_int
            L.ref_tri_quadrature.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int]
        L.ref_operator_apply.restype = ctypes.c_int
        L.ref_operator_apply.argtypes = [ctypes.c_int] * 4 + [_dp, _dp, ctypes.c_long] + [_dp] * 5
        L.ref_last_error.restype = ctypes.c_char_p
        _ref = L
    return _ref


def op_dims(op, fem, vec):
    nfa, dim = ctypes.c_int(), ctypes.c_int()
    rc = orc().orc_op_dims(op, fem, vec, ctypes.byref(nfa), ctypes.byref(dim))
    if rc:
        raise ValueError("unsupported operator/space (%d,%d,%d)" % (op, fem, vec))
    return nfa.value, dim.value


def tet_quadrature(order):
    p, w = _dp(), _dp()
    q = orc().orc_tet_quadrature(order, ctypes.byref(p), ctypes.byref(w))
    if q < 0:
        raise ValueError("quadrature order out of range")
    return (np.ctypeslib.as_array(p, shape=(4 * q,)).copy().reshape(q, 4),
            np.ctypeslib.as_array(w, shape=(q,)).copy())


def _xy(XY):
    """XY: (4, f, 3) array-like -> four contiguous 3 x f col-major buffers."""
    XY = np.ascontiguousarray(XY, dtype=np.float64)
    return [np.ascontiguousarray(XY[k]) for k in range(4)]


def quad_points(order, XY):
    xs = _xy(XY)
    f = xs[0].shape[0]
    q = len(tet_quadrature(order)[1])
    out = np.zeros((f, q, 3))
    orc().orc_quad_points(order, f, *[_P(x) for x in xs], _P(out))
    return out


def fem3dtet(form, XY, D=None, impl="oracle", mode=0, fuse=1, nthreads=1):
    """Element matrices of `form` = (opA, femA, vecA, opB, femB, vecB, order, ttype, layout) on the
    tets XY (4, f, 3).  Returns A (f, nfA, nfB): A[r, ia, ib] = reference A.data[ib + nfB*(ia + nfA*r)].
    impl: 'oracle' (C restatement) or 'ref' (the reference compiled in oracle/_ref)."""
    opA, femA, vecA, opB, femB, vecB, order, ttype, layout = form
    xs = _xy(XY)
    f = xs[0].shape[0]
    nfa, _ = op_dims(opA, femA, vecA)
    nfb, _ = op_dims(opB, femB, vecB)
    A = np.zeros((f, nfa, nfb))
    Dc = None if D is None else np.ascontiguousarray(D, dtype=np.float64)
    if impl == "oracle":
        fm = _Form(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, _P(Dc))
        rc = orc().orc_fem3dtet(ctypes.byref(fm), f, *[_P(x) for x in xs], _P(A))
        if rc:
            raise RuntimeError("orc_fem3dtet failed rc=%d" % rc)
    else:
        if Dc is None:
            Dc = np.zeros(1)
        rc = ref().ref_fem3dtet(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, _P(Dc), f,
                                *[_P(x) for x in xs], _P(A), mode, fuse, nthreads)
        if rc:
            raise RuntimeError("ref_fem3dtet failed rc=%d: %s" % (rc, ref().ref_last_error().decode()))
    return A


def tri_quadrature(order):
    p, w = _dp(), _dp()
    q = orc().orc_tri_quadrature(order, ctypes.byref(p), ctypes.byref(w))
    if q < 0:
        raise ValueError("quadrature order out of range")
    return (np.ctypeslib.as_array(p, shape=(3 * q,)).copy().reshape(q, 3),
            np.ctypeslib.as_array(w, shape=(q,)).copy())


def fem3dface(form, XY, face, D=None, impl="oracle"):
    """Surface-integral element matrices (fem3Dface, fem/operations/int_face.inl:160-199) of `form` over face face[r] of tet r
    (face k = vertices k, k+1, k+2 mod 4).  Layouts as fem3dtet; PER_POINT coefficients follow the triangle rule."""
    opA, femA, vecA, opB, femB, vecB, order, ttype, layout = form
    xs = _xy(XY)
    f = xs[0].shape[0]
    nfa, _ = op_dims(opA, femA, vecA)
    nfb, _ = op_dims(opB, femB, vecB)
    A = np.zeros((f, nfa, nfb))
    fc = np.ascontiguousarray(np.broadcast_to(np.asarray(face, dtype=np.int32), (f,)))
    Dc = None if D is None else np.ascontiguousarray(D, dtype=np.float64)
    if impl == "oracle":
        fm = _Form(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, _P(Dc))
        rc = orc().orc_fem3dface(ctypes.byref(fm), f, fc.ctypes.data_as(_ip), *[_P(x) for x in xs], _P(A))
        if rc:
            raise RuntimeError("orc_fem3dface failed rc=%d" % rc)
    else:
        if Dc is None:
            Dc = np.zeros(1)
        rc = ref().ref_fem3dface(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, _P(Dc), f, fc.ctypes.data_as(_ip),
                                 *[_P(x) for x in xs], _P(A))
        if rc:
            raise RuntimeError("ref_fem3dface failed rc=%d: %s" % (rc, ref().ref_last_error().decode()))
    return A


def operator_apply(op, fem, vec, XYL, XY, impl="oracle", W=None):
    """U table (f, nfa, q, dim): U[r,i,n,k] = reference U[k + dim*(n + q*(i + nfa*r))]."""
    xs = _xy(XY)
    f = xs[0].shape[0]
    XYL = np.ascontiguousarray(XYL, dtype=np.float64)
    q = XYL.size // 4
    nfa, dim = op_dims(op, fem, vec)
    U = np.zeros((f, nfa, q, dim))
    if impl == "oracle":
        rc = orc().orc_operator_apply(op, fem, vec, q, _P(XYL), f, *[_P(x) for x in xs], _P(U))
    else:
        W = np.full(q, 1.0 / q) if W is None else np.ascontiguousarray(W, dtype=np.float64)
        rc = ref().ref_operator_apply(op, fem, vec, q, _P(XYL), _P(W), f, *[_P(x) for x in xs], _P(U))
    if rc:
        raise RuntimeError("operator_apply failed rc=%d" % rc)
    return U


def scatter_csr(rowcode, colcode, A, F, row_begin, rowptr, colind, val, rhs, drop_val=1e-100):
    """Reference scatter rule into a pre-built sorted CSR; A (ne, ncol, nrow) i.e. col-major per element."""
    ne, nrow = rowcode.shape
    ncol = colcode.shape[1]
    rowcode = np.ascontiguousarray(rowcode, dtype=np.int64)
    colcode = np.ascontiguousarray(colcode, dtype=np.int64)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    return orc().orc_scatter_csr(
        ne, nrow, ncol, rowcode.ctypes.data_as(_lp), colcode.ctypes.data_as(_lp),
        _P(None if A is None else np.ascontiguousarray(A)), _P(None if F is None else np.ascontiguousarray(F)),
        row_begin, rowptr.ctypes.data_as(_lp), colind.ctypes.data_as(_ip), _P(val), _P(rhs), drop_val)


class _RefForm(ctypes.Structure):
    _fields_ = [("opA", ctypes.c_int), ("femA", ctypes.c_int), ("vecA", ctypes.c_int),
                ("opB", ctypes.c_int), ("femB", ctypes.c_int), ("vecB", ctypes.c_int),
                ("order", ctypes.c_int), ("ttype", ctypes.c_int), ("layout", ctypes.c_int), ("is_rhs", ctypes.c_int),
                ("D", _dp), ("alpha", ctypes.c_double), ("row_off", ctypes.c_int), ("col_off", ctypes.c_int)]


def ref_assemble_csr(problem, coords, tets, codes, row_begin, rowptr, colind, val, rhs, drop_val=1e-100, nthreads=1):
    """CPU baseline: reference element code (one cell per call) + restated scatter, threaded over cell ranges
    (oracle/ref_driver.cpp::ref_assemble_csr).  `problem` is an asm_oracle.Problem."""
    L = ref()
    L.ref_assemble_csr.restype = ctypes.c_int
    L.ref_assemble_csr.argtypes = [ctypes.c_int, ctypes.POINTER(_RefForm), ctypes.c_long, _dp, _lp, ctypes.c_int, _lp,
                                   ctypes.c_long, ctypes.c_long, _lp, _ip, _dp, _dp, ctypes.c_double, ctypes.c_int]
    forms, keep = [], []
    for fm in problem.mat_forms:
        fa, va = problem.vars[fm["trial"]]
        fb, vb = problem.vars[fm["test"]]
        D = None if fm.get("D") is None else np.ascontiguousarray(fm["D"], dtype=np.float64)
        keep.append(D)
        forms.append(_RefForm(fm["opA"], fa, va, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"], 0, _P(D),
                              fm.get("alpha", 1.0), problem.var_off[fm["test"]], problem.var_off[fm["trial"]]))
    for fm in problem.rhs_forms:
        fb, vb = problem.vars[fm["test"]]
        D = None if fm.get("D") is None else np.ascontiguousarray(fm["D"], dtype=np.float64)
        keep.append(D)
        forms.append(_RefForm(IDEN, P0, 1, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"], 1, _P(D),
                              fm.get("alpha", 1.0), problem.var_off[fm["test"]], 0))
    arr = (_RefForm * len(forms))(*forms)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    tets = np.ascontiguousarray(tets, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    rc = L.ref_assemble_csr(len(forms), arr, tets.shape[0], _P(coords), tets.ctypes.data_as(_lp), codes.shape[1],
                            codes.ctypes.data_as(_lp), row_begin, rowptr.size - 1, rowptr.ctypes.data_as(_lp),
                            colind.ctypes.data_as(_ip), _P(val), _P(rhs), drop_val, nthreads)
    if rc not in (0, -1):
        raise RuntimeError("ref_assemble_csr failed rc=%d: %s" % (rc, L.ref_last_error().decode()))
    return rc


# ---- the reference's own global assembler on the mock INMOST (oracle/ref_asm_driver.cpp, oracle/mock_inmost/inmost.h) ----------
_refasm = None


def have_refasm():
    return os.path.exists(os.path.join(HERE, "_ref", "libanifem_refasm.so"))


def refasm():
    global _refasm
    if _refasm is None:
        L = ctypes.CDLL(os.path.join(HERE, "_ref", "libanifem_refasm.so"))
        L.ref_last_error.restype = ctypes.c_char_p
        L.refasm_setup.restype = ctypes.c_int
        L.refasm_setup.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _ip, ctypes.c_long, _dp, ctypes.c_long, _lp, _lp, _lp, _lp]
        L.refasm_template.restype = ctypes.c_long
        L.refasm_assemble.restype = ctypes.c_int
        L.refasm_assemble.argtypes = [ctypes.c_int, ctypes.POINTER(_RefForm), ctypes.c_double, ctypes.c_int, _lp, _lp]
        L.refasm_get.argtypes = [_lp, _ip, _dp, _dp]
        _refasm = L
    return _refasm


def _ref_forms(problem):
    forms, keep = [], []
    for fm in problem.mat_forms:
        fa, va = problem.vars[fm["trial"]]
        fb, vb = problem.vars[fm["test"]]
        D = None if fm.get("D") is None else np.ascontiguousarray(fm["D"], dtype=np.float64)
        keep.append(D)
        forms.append(_RefForm(fm["opA"], fa, va, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"], 0, _P(D),
                              fm.get("alpha", 1.0), problem.var_off[fm["test"]], problem.var_off[fm["trial"]]))
    for fm in problem.rhs_forms:
        fb, vb = problem.vars[fm["test"]]
        D = None if fm.get("D") is None else np.ascontiguousarray(fm["D"], dtype=np.float64)
        keep.append(D)
        forms.append(_RefForm(IDEN, P0, 1, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"], 1, _P(D),
                              fm.get("alpha", 1.0), problem.var_off[fm["test"]], 0))
    return (_RefForm * max(1, len(forms)))(*forms), len(forms), keep


class RefAssembler:
    """The reference's Ani::Assembler (unmodified inmost_interface sources) on a tetrahedral mesh held by the mock INMOST:
    numbering of a GlobEnumeration type, index codes of fill_assemble_templates, AssembleTemplate, Assemble."""

    ENUM = ("ANITYPE", "MINIBLOCKS", "NATURAL", "DIMUNION", "BYELEMTYPE", "ETDIMBLOCKS")   # global_enumerator.h:393-401

    def __init__(self, coords, tets, variables, enum_type="NATURAL"):
        L = refasm()
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        tets = np.ascontiguousarray(tets, dtype=np.int64)
        nv = len(variables)
        fem = (ctypes.c_int * nv)(*[int(v[0]) for v in variables])
        vec = (ctypes.c_int * nv)(*[int(v[1]) for v in variables])
        nloc = sum(op_dims(IDEN, f, v)[0] for f, v in variables)
        self.codesC = np.zeros((tets.shape[0], nloc), dtype=np.int64)
        self.codesR = np.zeros((tets.shape[0], nloc), dtype=np.int64)
        nrows = ctypes.c_long()
        rc = L.refasm_setup(self.ENUM.index(enum_type), nv, fem, vec, coords.shape[0], _P(coords), tets.shape[0], tets.ctypes.data_as(_lp),
                            ctypes.byref(nrows), self.codesC.ctypes.data_as(_lp), self.codesR.ctypes.data_as(_lp))
        if rc != nloc:
            raise RuntimeError("refasm_setup failed rc=%d: %s" % (rc, L.ref_last_error().decode()))
        self.nrows, self.nloc = nrows.value, nloc

    def _get(self, nnz, with_rhs):
        rowptr = np.zeros(self.nrows + 1, dtype=np.int64)
        colind = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        rhs = np.zeros(self.nrows) if with_rhs else None
        refasm().refasm_get(rowptr.ctypes.data_as(_lp), colind.ctypes.data_as(_ip), _P(val), _P(rhs))
        return rowptr, colind, val, rhs

    def template(self):
        nnz = refasm().refasm_template()
        if nnz < 0:
            raise RuntimeError("AssembleTemplate failed: %s" % refasm().ref_last_error().decode())
        return self._get(nnz, False)[:2]

    def assemble(self, problem, drop_val=1e-100, include_template=False, ordered_insert=False):
        """Assemble(matrix, rhs, opts): returns (status, rowptr, colind, val, rhs) with the rows sorted by column"""
        arr, n, keep = _ref_forms(problem)
        nnz = ctypes.c_long()
        st = refasm().refasm_assemble(n, arr, drop_val, (1 if include_template else 0) | (2 if ordered_insert else 0), None, ctypes.byref(nnz))
        if st not in (0, -1):
            raise RuntimeError("Assemble failed rc=%d: %s" % (st, refasm().ref_last_error().decode()))
        return (st,) + self._get(nnz.value, True)
