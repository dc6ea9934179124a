"""inmost-fem_b200: B200-native element-matrix + global-assembly path behind the AniFem++ API.

This Python module is plumbing only: a ctypes binding of the C ABI in include/anifem_b200.h
(libanifem_b200.so, hand-written CUDA for sm_100a) used by tests/ and bench.py.  The product is
the shared library; the C++ mirror of the reference API lives in include/anifem_b200/*.hpp.
There is no CPU fallback: if the library is missing or no GPU is present, calls fail loudly.

Load it with `__graft_entry__.load_package()` (the directory name is not a Python identifier).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AFB_LIB", os.path.join(HERE, "libanifem_b200.so"))  # AFB_LIB: A/B runs of two builds in one session

HOST, DEVICE = 0, 1
IDEN, GRAD, DIV = 1, 2, 3
P0, P1, P2, P3 = 1, 2, 3, 4
TENSOR_NULL, TENSOR_SCALAR, TENSOR_SYMMETRIC, TENSOR_GENERAL = 1, 2, 3, 4
COEF_CONST, COEF_PER_TET, COEF_PER_POINT = 0, 1, 2

_dp = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class AfbForm(ctypes.Structure):
    """struct afb_form of include/anifem_b200.h"""
    _fields_ = [("opA", ctypes.c_int), ("femA", ctypes.c_int), ("vecA", ctypes.c_int),
                ("opB", ctypes.c_int), ("femB", ctypes.c_int), ("vecB", ctypes.c_int),
                ("quad_order", ctypes.c_int), ("tensor_type", ctypes.c_int), ("coef_layout", ctypes.c_int),
                ("coef_space", ctypes.c_int), ("D", ctypes.c_void_p), ("alpha", ctypes.c_double),
                ("row_off", ctypes.c_int), ("col_off", ctypes.c_int)]


EXPORTS = ["afb_ctx_create", "afb_ctx_destroy", "afb_last_error", "afb_sync", "afb_launch_count", "afb_stream_get",
           "afb_fem3dtet_batched", "afb_op_dims", "afb_tet_quadrature", "afb_quad_points",
           "afb_mesh_set", "afb_mesh_cube", "afb_mesh_orient", "afb_mesh_get",
           "afb_dofmap_set", "afb_dofmap_set_diag", "afb_dofmap_natural", "afb_dofmap_get",
           "afb_pattern_build", "afb_pattern_get", "afb_pattern_set", "afb_assemble", "afb_halo_add", "afb_last_times",
           "afb_dirichlet_set", "afb_fem3dapply_batched", "afb_eval_quadrature", "afb_priority_rows_set", "afb_assemble_phase",
           "afb_fields_set", "afb_fem3dface_batched", "afb_tri_quadrature",
           "afb_boundary_set", "afb_assemble_faces", "afb_assemble_elemental",
           "afb_comm_unique_id", "afb_comm_init", "afb_comm_set", "afb_halo_plan_set", "afb_halo_exchange_start",
           "afb_halo_exchange_finish", "afb_halo_exchange", "afb_assemble_distributed", "afb_last_kernels"]


def build(verbose=False):
    """compile libanifem_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)"""
    out = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libanifem_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libanifem_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, c64, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
        L.afb_ctx_create.argtypes = [ci, vp, ctypes.POINTER(vp)]
        L.afb_ctx_destroy.argtypes = [vp]
        L.afb_ctx_destroy.restype = None
        L.afb_last_error.argtypes = [vp]
        L.afb_last_error.restype = ctypes.c_char_p
        L.afb_sync.argtypes = [vp]
        L.afb_launch_count.argtypes = [vp, ci]
        L.afb_launch_count.restype = c64
        L.afb_stream_get.argtypes = [vp]
        L.afb_stream_get.restype = vp
        L.afb_fem3dtet_batched.argtypes = [vp, ctypes.POINTER(AfbForm), c64, vp, vp, vp, vp, vp, ci]
        L.afb_fem3dface_batched.argtypes = [vp, ctypes.POINTER(AfbForm), c64, vp, vp, vp, vp, vp, vp, ci]
        L.afb_tri_quadrature.argtypes = [ci, vp, vp, ci]
        L.afb_boundary_set.argtypes = [vp, c64, vp, vp, ci]
        L.afb_assemble_faces.argtypes = [vp, ci, ctypes.POINTER(AfbForm), ci, ctypes.POINTER(AfbForm), vp, vp, cd, ci]
        L.afb_op_dims.argtypes = [ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.afb_tet_quadrature.argtypes = [ci, vp, vp, ci]
        L.afb_quad_points.argtypes = [vp, ci, c64, vp, vp, vp, vp, vp, ci]
        L.afb_mesh_set.argtypes = [vp, c64, vp, vp, vp, c64, vp, vp, vp, vp, ci]
        L.afb_mesh_cube.argtypes = [vp, ci, ci, ci, cd, ci, ci, ci, ci, ci, ci]
        L.afb_mesh_orient.argtypes = [vp]
        L.afb_mesh_get.argtypes = [vp, _i64p, _i64p, vp, vp, ci]
        L.afb_dofmap_set.argtypes = [vp, ci, ci, vp, vp, c64, c64, c64, ci]
        L.afb_dofmap_natural.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.afb_dofmap_get.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), _i64p, _i64p, _i64p, vp, vp, ci]
        L.afb_dofmap_set_diag.argtypes = [vp, vp, ci]
        L.afb_fields_set.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci), ci, _i64p, ci, _i64p]
        L.afb_pattern_set.argtypes = [vp, vp, vp, c64, ci]
        L.afb_halo_add.argtypes = [vp, c64, vp, vp, vp]
        L.afb_pattern_build.argtypes = [vp, _i64p]
        L.afb_pattern_get.argtypes = [vp, vp, vp, ci]
        L.afb_assemble.argtypes = [vp, ci, ctypes.POINTER(AfbForm), ci, ctypes.POINTER(AfbForm), vp, vp, ci, cd, ci]
        L.afb_last_times.argtypes = [vp, _dp]
        L.afb_last_kernels.argtypes = [vp, ctypes.c_char_p, ci]
        L.afb_assemble_elemental.argtypes = [vp, c64, c64, vp, vp, ci, vp, vp, cd, ci]
        L.afb_comm_unique_id.argtypes = [vp]
        L.afb_comm_init.argtypes = [vp, vp, ci, ci]
        L.afb_comm_set.argtypes = [vp, vp, ci, ci]
        L.afb_halo_plan_set.argtypes = [vp, ci, c64, c64, vp, vp, vp, vp, vp, vp, ci]
        L.afb_halo_exchange_start.argtypes = [vp, vp, vp]
        L.afb_halo_exchange_finish.argtypes = [vp, vp, vp]
        L.afb_halo_exchange.argtypes = [vp, vp, vp]
        L.afb_assemble_distributed.argtypes = [vp, ci, ctypes.POINTER(AfbForm), ci, ctypes.POINTER(AfbForm), vp, vp, cd]
        L.afb_dirichlet_set.argtypes = [vp, vp, vp, ci]
        L.afb_priority_rows_set.argtypes = [vp, c64]
        L.afb_assemble_phase.argtypes = [vp, ci, ctypes.POINTER(AfbForm), ci, ctypes.POINTER(AfbForm), vp, vp, cd, ci]
        L.afb_fem3dapply_batched.argtypes = [vp, ci, ci, ci, ci, vp, c64, vp, vp, vp, vp, vp, vp, ci]
        L.afb_eval_quadrature.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, ci]
        _lib = L
    return _lib


class AfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("anifem_b200 error %d: %s" % (code, msg))
        self.code = code


def _ptr(a):
    """(address, mem_space) of a numpy array (host) or a torch tensor (host or cuda)"""
    if a is None:
        return None, HOST
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, HOST
    # torch tensor
    assert a.is_contiguous()
    return a.data_ptr(), (DEVICE if a.is_cuda else HOST)


def op_dims(op, fem, vec):
    nfa, dim = ctypes.c_int(), ctypes.c_int()
    rc = lib().afb_op_dims(op, fem, vec, ctypes.byref(nfa), ctypes.byref(dim))
    if rc:
        raise AfbError(rc, "unsupported operator/space")
    return nfa.value, dim.value


def tet_quadrature(order):
    q = lib().afb_tet_quadrature(order, None, None, 0)
    if q < 0:
        raise AfbError(q, "quadrature order must be in 0..20")
    p, w = np.zeros(4 * q), np.zeros(q)
    lib().afb_tet_quadrature(order, p.ctypes.data, w.ctypes.data, q)
    return p.reshape(q, 4), w


def tri_quadrature(order):
    q = lib().afb_tri_quadrature(order, None, None, 0)
    if q < 0:
        raise AfbError(q, "quadrature order must be in 0..20")
    p, w = np.zeros(3 * q), np.zeros(q)
    lib().afb_tri_quadrature(order, p.ctypes.data, w.ctypes.data, q)
    return p.reshape(q, 3), w


def make_form(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, D=None, alpha=1.0, row_off=0, col_off=0):
    """afb_form + a reference to D so that it outlives the call (the reference takes the tensor
    functor by const-ref with the same lifetime rule, fem/diff_tensor.h:67-70)"""
    addr, space = _ptr(D)
    f = AfbForm(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, space, addr, alpha, row_off, col_off)
    f._keep = D
    return f


class Context:
    """One assembly context = one GPU + one stream (afb_ctx)."""

    def __init__(self, device=0, stream=None):
        self._h = ctypes.c_void_p()
        rc = lib().afb_ctx_create(device, stream, ctypes.byref(self._h))
        if rc:
            raise AfbError(rc, lib().afb_last_error(None).decode())

    def close(self):
        if self._h:
            lib().afb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow=()):
        if rc < 0 and rc not in allow:
            raise AfbError(rc, lib().afb_last_error(self._h).decode())
        return rc

    def sync(self):
        self._ck(lib().afb_sync(self._h))

    def stream_handle(self):
        """cudaStream_t (integer) every kernel of this context is issued on"""
        return int(lib().afb_stream_get(self._h) or 0)

    def launch_count(self, reset=False):
        return int(lib().afb_launch_count(self._h, 1 if reset else 0))

    # ---- element level -------------------------------------------------------------------------
    def fem3dtet(self, form, XY, out=None):
        """Batched Ani::fem3Dtet: XY (4, f, 3) numpy -> A (f, nfA, nfB) numpy with
        A[r, ia, ib] = reference A.data[ib + nfB*(ia + nfA*r)]"""
        XY = np.ascontiguousarray(XY, dtype=np.float64)
        f = XY.shape[1]
        nfa, _ = op_dims(form.opA, form.femA, form.vecA)
        nfb, _ = op_dims(form.opB, form.femB, form.vecB)
        A = np.zeros((f, nfa, nfb)) if out is None else out
        xs = [np.ascontiguousarray(XY[k]) for k in range(4)]
        self._ck(lib().afb_fem3dtet_batched(self._h, ctypes.byref(form), f, xs[0].ctypes.data, xs[1].ctypes.data,
                                            xs[2].ctypes.data, xs[3].ctypes.data, A.ctypes.data, HOST))
        return A

    def fem3dface(self, form, XY, face, out=None):
        """Batched Ani::fem3Dface: XY (4, f, 3), face (f,) in 0..3 -> A (f, nfA, nfB) like fem3dtet"""
        XY = np.ascontiguousarray(XY, dtype=np.float64)
        f = XY.shape[1]
        nfa, _ = op_dims(form.opA, form.femA, form.vecA)
        nfb, _ = op_dims(form.opB, form.femB, form.vecB)
        A = np.zeros((f, nfa, nfb)) if out is None else out
        fc = np.ascontiguousarray(np.broadcast_to(np.asarray(face, dtype=np.int32), (f,)))
        xs = [np.ascontiguousarray(XY[k]) for k in range(4)]
        self._ck(lib().afb_fem3dface_batched(self._h, ctypes.byref(form), f, fc.ctypes.data, xs[0].ctypes.data, xs[1].ctypes.data,
                                             xs[2].ctypes.data, xs[3].ctypes.data, A.ctypes.data, HOST))
        return A

    def quad_points(self, order, XY):
        XY = np.ascontiguousarray(XY, dtype=np.float64)
        f = XY.shape[1]
        q = tet_quadrature(order)[1].size
        out = np.zeros((f, q, 3))
        xs = [np.ascontiguousarray(XY[k]) for k in range(4)]
        self._ck(lib().afb_quad_points(self._h, order, f, xs[0].ctypes.data, xs[1].ctypes.data, xs[2].ctypes.data,
                                       xs[3].ctypes.data, out.ctypes.data, HOST))
        return out

    def fem3dapply(self, op, fem, vec, XYL, XY, dofs):
        """Batched Ani::fem3DapplyL: XYL (q,4) barycentric points, XY (4,f,3), dofs (f,nfa) -> (f,q,dim)"""
        XYL = np.ascontiguousarray(XYL, dtype=np.float64)
        XY = np.ascontiguousarray(XY, dtype=np.float64)
        dofs = np.ascontiguousarray(dofs, dtype=np.float64)
        f, q = XY.shape[1], XYL.shape[0]
        _, dim = op_dims(op, fem, vec)
        out = np.zeros((f, q, dim))
        xs = [np.ascontiguousarray(XY[k]) for k in range(4)]
        self._ck(lib().afb_fem3dapply_batched(self._h, op, fem, vec, q, XYL.ctypes.data, f, xs[0].ctypes.data, xs[1].ctypes.data,
                                              xs[2].ctypes.data, xs[3].ctypes.data, dofs.ctypes.data, out.ctypes.data, HOST))
        return out

    def eval_quadrature(self, op, fem, vec, col_off, order, u, out=None):
        """Op(u_h) at the quadrature points of `order` on the context's mesh: (ntet, q, dim); u numpy or torch cuda"""
        q = tet_quadrature(order)[1].size
        _, dim = op_dims(op, fem, vec)
        _, nt = self.mesh_sizes()
        pu, su = _ptr(u)
        if out is None:
            if su == DEVICE:
                import torch
                out = torch.empty((nt, q, dim), dtype=torch.float64, device=u.device)
            else:
                out = np.zeros((nt, q, dim))
        po, so = _ptr(out)
        assert su == so
        self._ck(lib().afb_eval_quadrature(self._h, op, fem, vec, col_off, order, pu, po, su))
        return out

    # ---- mesh ----------------------------------------------------------------------------------
    def mesh_set(self, coords, tets):
        """coords (nnode,3) float64, tets (ntet,4) integer (numpy, host)"""
        coords = np.asarray(coords, dtype=np.float64)
        xs = [np.ascontiguousarray(coords[:, k]) for k in range(3)]
        vs = [np.ascontiguousarray(np.asarray(tets)[:, k], dtype=np.int32) for k in range(4)]
        self._ck(lib().afb_mesh_set(self._h, coords.shape[0], xs[0].ctypes.data, xs[1].ctypes.data, xs[2].ctypes.data,
                                    len(vs[0]), vs[0].ctypes.data, vs[1].ctypes.data, vs[2].ctypes.data, vs[3].ctypes.data, HOST))

    def mesh_cube(self, nx, ny, nz, size=1.0, block=None):
        bx, by, bz, lx, ly, lz = (0, 0, 0, nx, ny, nz) if block is None else block
        self._ck(lib().afb_mesh_cube(self._h, nx, ny, nz, size, bx, by, bz, lx, ly, lz))

    def mesh_orient(self):
        self._ck(lib().afb_mesh_orient(self._h))

    def mesh_sizes(self):
        nn, nt = ctypes.c_int64(), ctypes.c_int64()
        self._ck(lib().afb_mesh_get(self._h, ctypes.byref(nn), ctypes.byref(nt), None, None, HOST))
        return nn.value, nt.value

    def mesh_get(self):
        nn, nt = self.mesh_sizes()
        xyz = np.zeros((3, nn))
        v = np.zeros((4, nt), dtype=np.int32)
        self._ck(lib().afb_mesh_get(self._h, None, None, xyz.ctypes.data, v.ctypes.data, HOST))
        return np.ascontiguousarray(xyz.T), np.ascontiguousarray(v.T)

    def mesh_get_torch(self):
        """coords (nnode,3) float64 and tets (ntet,4) int32 as torch tensors on this context's GPU"""
        import torch
        nn, nt = self.mesh_sizes()
        xyz = torch.empty((3, nn), dtype=torch.float64, device="cuda")
        v = torch.empty((4, nt), dtype=torch.int32, device="cuda")
        self._ck(lib().afb_mesh_get(self._h, None, None, xyz.data_ptr(), v.data_ptr(), DEVICE))
        return xyz.t().contiguous(), v.t().contiguous()

    def pattern_get_torch(self):
        import torch
        _, _, rb, re, _ = self.dofmap_info()
        rowptr = torch.empty(re - rb + 1, dtype=torch.int64, device="cuda")
        colind = torch.empty(max(self.nnz, 1), dtype=torch.int32, device="cuda")
        self._ck(lib().afb_pattern_get(self._h, rowptr.data_ptr(), colind.data_ptr(), DEVICE))
        return rowptr, colind[:self.nnz]

    # ---- dof map -------------------------------------------------------------------------------
    def dofmap_set(self, rowcode, colcode, row_begin, row_end, ncols_global):
        rowcode = np.ascontiguousarray(rowcode, dtype=np.int64)
        colcode = np.ascontiguousarray(colcode, dtype=np.int64)
        self._ck(lib().afb_dofmap_set(self._h, rowcode.shape[1], colcode.shape[1], rowcode.ctypes.data, colcode.ctypes.data,
                                      row_begin, row_end, ncols_global, HOST))

    def dofmap_set_any(self, rowcode, colcode, row_begin, row_end, ncols_global, diag_col=None):
        """like dofmap_set but accepts numpy arrays or torch tensors (host or cuda); optional forced-diagonal columns"""
        pr, sr = _ptr(rowcode)
        pc, sc = _ptr(colcode)
        assert sr == sc
        self._ck(lib().afb_dofmap_set(self._h, rowcode.shape[1], colcode.shape[1], pr, pc, row_begin, row_end, ncols_global, sr))
        if diag_col is not None:
            pd, sd = _ptr(diag_col)
            self._ck(lib().afb_dofmap_set_diag(self._h, pd, sd))

    def fields_set(self, fields):
        """scalar fields of an explicit dof map: [(fem, loff, [(row_first, row_count), ...], [(col_first, col_count), ...]), ...]
        (afb_fields_set: row intervals relative to row_begin, column intervals global); [] clears"""
        n = len(fields)
        if n == 0:
            self._ck(lib().afb_fields_set(self._h, 0, None, None, 1, None, 1, None))
            return
        nsr = max(1, max(len(f[2]) for f in fields))
        nsc = max(1, max(len(f[3]) for f in fields))
        rs = np.zeros((n, nsr, 2), dtype=np.int64)
        cs = np.zeros((n, nsc, 2), dtype=np.int64)
        for k, f in enumerate(fields):
            for j, (a, b) in enumerate(f[2]):
                rs[k, j] = (a, b)
            for j, (a, b) in enumerate(f[3]):
                cs[k, j] = (a, b)
        fem = (ctypes.c_int * n)(*[int(f[0]) for f in fields])
        loff = (ctypes.c_int * n)(*[int(f[1]) for f in fields])
        self._ck(lib().afb_fields_set(self._h, n, fem, loff, nsr, rs.ctypes.data_as(_i64p), nsc, cs.ctypes.data_as(_i64p)))

    def pattern_set(self, rowptr, colind):
        pr, sr = _ptr(rowptr)
        pc, sc = _ptr(colind)
        assert sr == sc
        self.nnz = int(colind.shape[0])
        self._ck(lib().afb_pattern_set(self._h, pr, pc, self.nnz, sr))

    # ---- multi-GPU exchange inside the library (afb_comm.cu)
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId (rank 0 creates them, every rank passes them to comm_init)"""
        buf = ctypes.create_string_buffer(128)
        rc = lib().afb_comm_unique_id(buf)
        if rc:
            raise AfbError(rc, lib().afb_last_error(None).decode())
        return buf.raw

    def comm_init(self, id128, rank, nranks):
        self._ck(lib().afb_comm_init(self._h, ctypes.c_char_p(bytes(id128)), int(rank), int(nranks)))

    def halo_plan_set(self, n_own, nnz_own, send_val, send_rhs, recv_val, recv_rhs, val_slots, rhs_slots):
        """counts: python lists per rank; val_slots / rhs_slots: int64 torch cuda tensors (concatenated in rank order)"""
        n = len(send_val)
        arr = lambda v: (ctypes.c_int64 * n)(*[int(x) for x in v])
        sv, sr, rv, rr = arr(send_val), arr(send_rhs), arr(recv_val), arr(recv_rhs)
        pv = val_slots.data_ptr() if val_slots is not None and val_slots.numel() else None
        pr = rhs_slots.data_ptr() if rhs_slots is not None and rhs_slots.numel() else None
        self._ck(lib().afb_halo_plan_set(self._h, n, int(n_own), int(nnz_own), sv, sr, rv, rr, pv, pr, DEVICE))

    def halo_exchange(self, val, rhs):
        pv, _ = _ptr(val)
        pr, _ = _ptr(rhs)
        self._ck(lib().afb_halo_exchange(self._h, pv, pr))

    def assemble_distributed(self, forms, rhs_forms, val, rhs, drop_val=1e-100):
        """afb_assemble_distributed on torch cuda tensors (extended arrays): phased assembly + NCCL exchange + additions"""
        fa = (AfbForm * max(1, len(forms)))(*forms)
        fr = (AfbForm * max(1, len(rhs_forms)))(*rhs_forms)
        pv, sv = _ptr(val)
        pr, sr = _ptr(rhs)
        assert (val is None or sv == DEVICE) and (rhs is None or sr == DEVICE)
        return self._ck(lib().afb_assemble_distributed(self._h, len(forms), fa, len(rhs_forms), fr, pv, pr, drop_val), allow=(-1,))

    def halo_add(self, slot, contrib, dst):
        """dst[slot] += contrib for the contributions of one peer (torch cuda tensors; slot int64, distinct)"""
        self._ck(lib().afb_halo_add(self._h, int(slot.shape[0]), slot.data_ptr(), contrib.data_ptr(), dst.data_ptr()))

    def dofmap_natural(self, variables):
        n = len(variables)
        fem = (ctypes.c_int * n)(*[v[0] for v in variables])
        vec = (ctypes.c_int * n)(*[v[1] for v in variables])
        self._ck(lib().afb_dofmap_natural(self._h, n, fem, vec))

    def dofmap_info(self):
        nr, nc = ctypes.c_int(), ctypes.c_int()
        rb, re, ng = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._ck(lib().afb_dofmap_get(self._h, ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(rb), ctypes.byref(re),
                                      ctypes.byref(ng), None, None, HOST))
        return nr.value, nc.value, rb.value, re.value, ng.value

    def dofmap_get(self):
        nr, nc, rb, re, ng = self.dofmap_info()
        _, nt = self.mesh_sizes()
        row = np.zeros((nt, nr), dtype=np.int64)
        col = np.zeros((nt, nc), dtype=np.int64)
        self._ck(lib().afb_dofmap_get(self._h, None, None, None, None, None, row.ctypes.data, col.ctypes.data, HOST))
        return row, col

    def dirichlet_set(self, is_dirichlet, value):
        """essential BCs per global dof (numpy uint8 / float64 or torch tensors); None clears (afb_dirichlet_set)"""
        if is_dirichlet is None:
            self._ck(lib().afb_dirichlet_set(self._h, None, None, HOST))
            return
        if isinstance(is_dirichlet, np.ndarray):
            is_dirichlet = np.ascontiguousarray(is_dirichlet, dtype=np.uint8)
            value = np.ascontiguousarray(value, dtype=np.float64)
        pf, sf = _ptr(is_dirichlet)
        pv, sv = _ptr(value)
        assert sf == sv
        self._keep_dir = (is_dirichlet, value)
        self._ck(lib().afb_dirichlet_set(self._h, pf, pv, sf))

    # ---- pattern -------------------------------------------------------------------------------
    def pattern_build(self):
        nnz = ctypes.c_int64()
        self._ck(lib().afb_pattern_build(self._h, ctypes.byref(nnz)))
        self.nnz = nnz.value
        return nnz.value

    def pattern_get(self):
        _, _, rb, re, _ = self.dofmap_info()
        rowptr = np.zeros(re - rb + 1, dtype=np.int64)
        colind = np.zeros(max(self.nnz, 1), dtype=np.int32)
        self._ck(lib().afb_pattern_get(self._h, rowptr.ctypes.data, colind.ctypes.data, HOST))
        return rowptr, colind[:self.nnz]

    # ---- assembly ------------------------------------------------------------------------------
    def assemble(self, forms, rhs_forms, val=None, rhs=None, accumulate=False, drop_val=1e-100):
        """val / rhs: numpy arrays (host) or torch cuda tensors (device), both in the same space.
        Returns the reference's status code (0 ok, -1 non-finite local value)."""
        fa = (AfbForm * max(1, len(forms)))(*forms)
        fr = (AfbForm * max(1, len(rhs_forms)))(*rhs_forms)
        pv, sv = _ptr(val)
        pr, sr = _ptr(rhs)
        space = sv if val is not None else sr
        if val is not None and rhs is not None:
            assert sv == sr
        return self._ck(lib().afb_assemble(self._h, len(forms), fa, len(rhs_forms), fr, pv, pr, 1 if accumulate else 0,
                                           drop_val, space), allow=(-1,))

    def assemble_elemental(self, e_lo, A_elem, F_elem, val=None, rhs=None, drop_val=1e-100):
        """afb_assemble_elemental: adds caller-evaluated local matrices A_elem (nel, ncol_loc, nrow_loc) = the reference's
        column-major m_A per cell, and local right-hand sides F_elem (nel, nrow_loc), of the cells e_lo.. into val / rhs"""
        nel = (A_elem if A_elem is not None else F_elem).shape[0]
        pa, sa = _ptr(A_elem)
        pf, sf = _ptr(F_elem)
        pv, sv = _ptr(val)
        pr, sr = _ptr(rhs)
        espace = sa if A_elem is not None else sf
        space = sv if val is not None else sr
        return self._ck(lib().afb_assemble_elemental(self._h, int(e_lo), int(nel), pa, pf, espace, pv, pr, drop_val, space), allow=(-1,))

    def boundary_set(self, face_tet, face_num):
        """boundary faces carrying surface terms: face face_num[b] of element face_tet[b] (numpy int arrays)"""
        ft = np.ascontiguousarray(face_tet, dtype=np.int32)
        fn = np.ascontiguousarray(face_num, dtype=np.int32)
        assert ft.shape == fn.shape
        self._ck(lib().afb_boundary_set(self._h, ft.shape[0], ft.ctypes.data, fn.ctypes.data, HOST))

    def assemble_faces(self, forms, rhs_forms, val=None, rhs=None, drop_val=1e-100):
        """adds the surface forms over the faces of boundary_set into val / rhs (afb_assemble_faces)"""
        fa = (AfbForm * max(1, len(forms)))(*forms)
        fr = (AfbForm * max(1, len(rhs_forms)))(*rhs_forms)
        pv, sv = _ptr(val)
        pr, sr = _ptr(rhs)
        space = sv if val is not None else sr
        if val is not None and rhs is not None:
            assert sv == sr
        return self._ck(lib().afb_assemble_faces(self._h, len(forms), fa, len(rhs_forms), fr, pv, pr, drop_val, space), allow=(-1,))

    def priority_rows_set(self, first_priority_row):
        self._ck(lib().afb_priority_rows_set(self._h, int(first_priority_row)))

    def assemble_phase(self, forms, rhs_forms, val, rhs, phase, drop_val=1e-100):
        """afb_assemble_phase on torch cuda tensors: phase 1 returns without synchronising, phase 2 returns the status"""
        fa = (AfbForm * max(1, len(forms)))(*forms)
        fr = (AfbForm * max(1, len(rhs_forms)))(*rhs_forms)
        pv, sv = _ptr(val)
        pr, sr = _ptr(rhs)
        assert (val is None or sv == DEVICE) and (rhs is None or sr == DEVICE)
        return self._ck(lib().afb_assemble_phase(self._h, len(forms), fa, len(rhs_forms), fr, pv, pr, drop_val, phase), allow=(-1,))

    def last_times(self):
        t = (ctypes.c_double * 4)()
        lib().afb_last_times(self._h, t)
        kern = {0: ("k_element_generic", "k_gather"), 1: ("k_geom", "k_gather_tensor"), 2: ("k_geom", "k_rows_cl"), 3: ("k_geom", "k_rings")}[int(t[3])]
        if int(t[3]) == 0:   # generic staged path: the library names the kernels that ran
            buf = ctypes.create_string_buffer(128)
            lib().afb_last_kernels(self._h, buf, 128)
            names = buf.value.decode().split("|")
            if len(names) == 2 and names[0] and names[1]:
                kern = (names[0], names[1])
        return {"element_ms": t[0], "gather_ms": t[1], "copy_ms": t[2], "fused_path": bool(t[3]), "element_kernel": kern[0], "gather_kernel": kern[1]}
