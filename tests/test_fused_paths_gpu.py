"""GPU parity tests aimed at the fused assembly paths added after the first slice:
  * cluster-tiled gather k_rows_cl with several clusters per mesh, long / short row regions, ragged slices, padding visits;
  * block-decomposed assembly of vector / mixed spaces (afb_blocks.cu) incl. per-tet tensors, accumulate, matrix-only, rhs-only;
  * register-tiled element kernel k_element_sq (all tensor kinds and layouts) through the generic staged path;
  * size-independent properties at a larger size: symmetry of the matrix, constant null space of the stiffness rows,
    sum of the load vector = volume, bit-reproducibility.
All comparisons go through the C ABI against the CPU oracle (CSR pattern bit-exact, values within 1e-12 of the row scale)."""
import os

import numpy as np
import pytest

import golden_cases as gc
import problems

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, tag, env=None):
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        nnz = ctx.pattern_build()
        rowptr, colind = ctx.pattern_get()
        val, rhs = np.full(nnz, np.nan), np.full(rowptr.size - 1, np.nan)
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
        path = ctx.last_times()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    rp, ci, v, r, st = M.assemble(prob, co, te, dm)
    assert st == 0
    assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci), tag + ": pattern not bit-exact"
    rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
    rowmax[rowmax == 0] = 1.0
    err = (np.abs(val - v) / np.repeat(rowmax, np.diff(rp))).max()
    rs = np.abs(r).max()
    rerr = np.abs(rhs - r).max() / (rs if rs > 0 else 1.0)
    print("%s [%s + %s]: ntet=%d nnz=%d err A %.2e rhs %.2e" % (tag, path["element_kernel"], path["gather_kernel"], te.shape[0], nnz, err, rerr))
    assert err <= RTOL and rerr <= RTOL, (tag, err, rerr)
    return val, rhs, path


def _mesh(pkg, ctx, M, n, variables, jitter=0.0, seed=0):
    co, te, _ = M.cube_mesh(*n)
    if jitter:
        rng = np.random.default_rng(seed)
        co = co + jitter * rng.standard_normal(co.shape) / max(n)
        p = co[te]
        det = np.linalg.det(p[:, :3, :] - p[:, 3:4, :])
        te[det < 0] = te[det < 0][:, [0, 1, 3, 2]]
        ctx.mesh_set(co, te)
    else:
        ctx.mesh_cube(*n)
    ctx.dofmap_natural(variables)
    dm = M.DofMap(te, variables, nnode=co.shape[0])
    return co, te, dm


@pytest.mark.parametrize("chunk", [32, 96, 512])
@pytest.mark.parametrize("space", ["p1", "p2", "p3"])
def test_cluster_gather_many_clusters(pkg, ctx, asm_oracle, chunk, space):
    """small chunks force many clusters, partially filled slices and halo elements on a mesh the oracle finishes quickly"""
    M = asm_oracle
    fem = {"p1": gc.P1, "p2": gc.P2, "p3": gc.P3}[space]
    n = {"p1": (7, 6, 5), "p2": (5, 4, 4), "p3": (3, 3, 3)}[space]
    co, te, dm = _mesh(pkg, ctx, M, n, [(fem, 1)], jitter=0.08, seed=chunk)
    rng = np.random.default_rng(chunk + fem)
    K = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, te.shape[0], 4)
    c = gc.tensor(rng, gc.T_SCALAR, gc.L_PER_TET, 1, 1, te.shape[0], 4)
    _, forms, rhsf, prob = problems._mk(pkg, M, [(fem, 1)],
                                        [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K, 1.0),
                                         (0, 0, gc.IDEN, gc.IDEN, 3, gc.T_SCALAR, gc.L_PER_TET, c, 0.5)],
                                        [(0, gc.IDEN, 2, gc.T_SCALAR, gc.L_PER_TET, c, 2.0)])
    # the ring kernel (afb_rings.cu, tests/test_rings_gpu.py) would take the P2 case: this test is about the row gather
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "%s chunk %d" % (space, chunk),
                                 {"AFB_ROWS_CHUNK": str(chunk), "AFB_DISABLE_RING_KERNEL": "1"})
    assert path["gather_kernel"] == "k_rows_cl"


def test_cluster_gather_general_tensor_and_drop(pkg, ctx, asm_oracle):
    """GENERAL (non-symmetric) 3x3 tensor = 9 components; drop_val larger than some entries changes values, not the pattern"""
    M = asm_oracle
    co, te, dm = _mesh(pkg, ctx, M, (4, 4, 3), [(gc.P2, 1)])
    rng = np.random.default_rng(3)
    K = gc.tensor(rng, gc.T_GENERAL, gc.L_PER_TET, 3, 3, te.shape[0], 4)
    _, forms, rhsf, prob = problems._mk(pkg, M, [(gc.P2, 1)], [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_GENERAL, gc.L_PER_TET, K, 1.0)], [])
    val, _, path = _oracle_compare(ctx, M, prob, forms, [], co, te, dm, "general tensor")
    assert path["gather_kernel"] == "k_rows_cl"
    # drop_val: contributions with |A_e(i,j)| <= drop are skipped (assembler.inl:416); compare with the oracle run with the same drop
    drop = 0.02 * np.abs(val).max()
    nnz = ctx.pattern_build()
    v2 = np.zeros(nnz)
    assert ctx.assemble(forms, [], v2, None, drop_val=drop) == 0
    rp, ci, v, r, st = M.assemble(prob, co, te, dm, drop_val=drop)
    rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
    rowmax[rowmax == 0] = 1.0
    # entries within rounding of the threshold may flip: allow a handful of element contributions of size drop
    bad = np.abs(v2 - v) / np.repeat(rowmax, np.diff(rp)) > RTOL
    assert bad.mean() < 1e-3 and (np.abs(v2 - v)[bad] <= 1.001 * drop * 4).all()
    assert not np.array_equal(v2, val)


@pytest.mark.parametrize("problem", ["c4", "c5"])
def test_block_path_accumulate_and_partial(pkg, ctx, asm_oracle, problem):
    M = asm_oracle
    variables = [(gc.P2, 3)] if problem == "c4" else [(gc.P2, 3), (gc.P1, 1)]
    co, te, dm = _mesh(pkg, ctx, M, (3, 3, 2), variables)
    _, forms, rhsf, prob = (problems.c4_p2_elasticity if problem == "c4" else problems.c5_stokes)(pkg, M, co, te)
    a, fa, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, problem + " blocks")
    assert path["gather_kernel"] == "k_rows_cl", "the block path did not run"
    nnz, nrows = a.size, fa.size
    b, fb = a.copy(), fa.copy()
    assert ctx.assemble(forms, rhsf, b, fb, accumulate=True) == 0
    assert np.array_equal(b, 2 * a) and np.array_equal(fb, 2 * fa), "Assemble must add into the existing contents"
    c = np.full(nnz, np.nan)
    assert ctx.assemble(forms, [], c, None) == 0 and np.array_equal(c, a)       # AssembleMatrix
    fc = np.full(nrows, np.nan)
    assert ctx.assemble([], rhsf, None, fc) == 0 and np.array_equal(fc, fa)     # AssembleRHS
    # same numbers as the generic staged path up to rounding
    os.environ["AFB_DISABLE_TENSOR_PATH"] = "1"
    try:
        d, fd = np.zeros(nnz), np.zeros(nrows)
        assert ctx.assemble(forms, rhsf, d, fd) == 0
        lt = ctx.last_times()
        assert not lt["fused_path"] and lt["gather_kernel"].startswith("k_gather")
    finally:
        os.environ.pop("AFB_DISABLE_TENSOR_PATH")
    assert np.abs(d - a).max() <= 1e-12 * np.abs(a).max() and np.abs(fd - fa).max() <= 1e-12 * np.abs(fa).max()


def test_block_path_per_tet_elasticity_tensor(pkg, ctx, asm_oracle):
    """FemVec<3,P2>: symmetric 9x9 tensor varying per tet + vector mass with a 3x3 tensor + body force per tet"""
    M = asm_oracle
    variables = [(gc.P2, 3)]
    co, te, dm = _mesh(pkg, ctx, M, (3, 2, 3), variables, jitter=0.05, seed=11)
    rng = np.random.default_rng(12)
    nt = te.shape[0]
    C = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 9, 9, nt, 4)
    Mm = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, nt, 4)
    f = np.ascontiguousarray(rng.standard_normal((nt, 3)))
    _, forms, rhsf, prob = problems._mk(pkg, M, variables,
                                        [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, C, 1.0),
                                         (0, 0, gc.IDEN, gc.IDEN, 2, gc.T_SYMMETRIC, gc.L_PER_TET, Mm, 0.3)],
                                        [(0, gc.IDEN, 2, gc.T_GENERAL, gc.L_PER_TET, f, 1.0)])
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "per-tet elasticity")
    assert path["gather_kernel"] == "k_rows_cl"


def test_block_path_mixed_p2_p1_scalars(pkg, ctx, asm_oracle):
    """two scalar variables P2 and P1 coupled by mass-like and convection-like blocks (rectangular pair plans)"""
    M = asm_oracle
    variables = [(gc.P2, 1), (gc.P1, 1)]
    co, te, dm = _mesh(pkg, ctx, M, (3, 3, 3), variables)
    rng = np.random.default_rng(5)
    nt = te.shape[0]
    b = np.ascontiguousarray(rng.standard_normal((nt, 3)))  # GRAD(A) x IDEN(B): K is 1 x 3
    c = gc.tensor(rng, gc.T_SCALAR, gc.L_PER_TET, 1, 1, nt, 4)
    _, forms, rhsf, prob = problems._mk(pkg, M, variables,
                                        [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_NULL, gc.L_CONST, None, 1.0),
                                         (1, 1, gc.GRAD, gc.GRAD, 2, gc.T_SCALAR, gc.L_PER_TET, c, 1.0),
                                         (1, 0, gc.IDEN, gc.IDEN, 2, gc.T_SCALAR, gc.L_PER_TET, c, -2.0),
                                         (0, 1, gc.GRAD, gc.IDEN, 2, gc.T_GENERAL, gc.L_PER_TET, b, 1.5)],
                                        [(1, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)])
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "mixed scalars")
    assert path["gather_kernel"] == "k_rows_cl"


def _segmented_fields(M, dm, variables, world):
    """fields of a per-rank NATURAL numbering held by ONE context (rows = global ids): one interval per owner rank"""
    fields, loff = [], 0
    num = [np.bincount(dm.owner[d], minlength=world) for d in range(4)]
    for v, (fem, vec) in enumerate(variables):
        ds = [d for d in range(4) if M.NDOF[fem][d]]
        for c in range(vec):
            segs = [(int(dm.beg_ind[p] + dm.grp_off[(v, c, ds[0])][p]), int(sum(num[d][p] * M.NDOF[fem][d] for d in ds))) for p in range(world)]
            fields.append((fem, loff, segs, segs))
            loff += gc.NF[fem]
    return fields


@pytest.mark.parametrize("problem,world", [("c5", 2), ("c4", 3), ("c5", 8)])
def test_block_path_segmented_numbering(pkg, ctx, asm_oracle, problem, world):
    """the per-rank NATURAL numbering of a partitioned mesh (fields contiguous inside every rank's interval only,
    global_enumerator.cpp:594-604) on one context: afb_fields_set keeps the problem on the block path; blocks of a row
    are then non-contiguous runs found by search.  Also with a superset pattern (columns no local cell touches, as the
    interface rows of a multi-GPU run have): the extra entries stay zero."""
    M = asm_oracle
    variables = [(gc.P2, 3)] if problem == "c4" else [(gc.P2, 3), (gc.P1, 1)]
    n = (4, 3, 2) if world < 8 else (4, 4, 2)
    co, te, cr = M.cube_mesh(*n, nranks=world)
    ctx.mesh_set(co, te)
    dm = M.DofMap(te, variables, cr, world, nnode=co.shape[0])
    rowcode, colcode = dm.codes(None)
    ctx.dofmap_set(rowcode, colcode, 0, dm.nrows, dm.nrows)
    ctx.fields_set(_segmented_fields(M, dm, variables, world))
    _, forms, rhsf, prob = (problems.c4_p2_elasticity if problem == "c4" else problems.c5_stokes)(pkg, M, co, te)
    a, fa, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "%s segmented x%d" % (problem, world))
    assert path["gather_kernel"] == "k_rows_cl", "the block path did not run"
    b, fb = a.copy(), fa.copy()
    assert ctx.assemble(forms, rhsf, b, fb, accumulate=True) == 0
    assert np.array_equal(b, 2 * a) and np.array_equal(fb, 2 * fa)
    # superset pattern: one extra column in every third row
    rowptr, colind = ctx.pattern_get()
    nrows = rowptr.size - 1
    rows = np.repeat(np.arange(nrows), np.diff(rowptr))
    key = rows.astype(np.int64) * dm.nrows + colind
    extra_r = np.arange(0, nrows, 3)
    extra = extra_r.astype(np.int64) * dm.nrows + (extra_r * 7 + 3) % dm.nrows
    key2 = np.union1d(key, extra)
    rp2 = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(rp2, key2 // dm.nrows + 1, 1)
    rp2 = np.cumsum(rp2)
    ci2 = (key2 % dm.nrows).astype(np.int32)
    ctx.pattern_set(rp2, ci2)
    v2, f2 = np.full(ci2.size, np.nan), np.full(nrows, np.nan)
    assert ctx.assemble(forms, rhsf, v2, f2) == 0
    assert ctx.last_times()["gather_kernel"] == "k_rows_cl"
    old = np.isin(key2, key)
    assert np.array_equal(v2[old], a) and np.all(v2[~old] == 0.0) and np.array_equal(f2, fa)


@pytest.mark.parametrize("ttype", [gc.T_NULL, gc.T_SCALAR, gc.T_SYMMETRIC, gc.T_GENERAL])
@pytest.mark.parametrize("layout", [gc.L_CONST, gc.L_PER_TET, gc.L_PER_POINT])
def test_tiled_element_kernel_all_tensor_kinds(pkg, ctx, asm_oracle, ttype, layout):
    """k_element_sq through the generic staged path: P3 stiffness (order 4) + P3 mass (order 6)"""
    if ttype == gc.T_NULL and layout != gc.L_CONST:
        pytest.skip("no data for TENSOR_NULL")
    M = asm_oracle
    variables = [(gc.P3, 1)]
    co, te, dm = _mesh(pkg, ctx, M, (2, 2, 2), variables, jitter=0.05, seed=7)
    rng = np.random.default_rng(100 * ttype + layout)
    nt = te.shape[0]
    K = gc.tensor(rng, ttype, layout, 3, 3, nt, 14)
    mt = gc.T_NULL if ttype == gc.T_NULL else gc.T_SCALAR
    c = gc.tensor(rng, mt, layout, 1, 1, nt, 24)
    if layout == gc.L_PER_POINT:  # one record per element: (ntet, q * dlen), the FusiveTensor layout D[dlen*(n + q*r)]
        K, c = K.reshape(nt, -1), c.reshape(nt, -1)
    _, forms, rhsf, prob = problems._mk(pkg, M, variables,
                                        [(0, 0, gc.GRAD, gc.GRAD, 4, ttype, layout, K, 1.0), (0, 0, gc.IDEN, gc.IDEN, 6, mt, layout, c, 0.7)],
                                        [(0, gc.IDEN, 3, gc.T_NULL, gc.L_CONST, None, 1.0)])
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "tiled ttype %d layout %d" % (ttype, layout), {"AFB_DISABLE_TENSOR_PATH": "1"})
    # generic staged path: the square P3 forms in one launch of the FP64 tensor-core kernel (k_element_generic only for the rhs)
    assert not path["fused_path"] and path["element_kernel"] == "k_element_mma" and path["gather_kernel"].startswith("k_gather_cols")


@pytest.mark.parametrize("case", ["p2_fused", "p2_generic", "p1", "th_blocks"])
def test_dirichlet_conditions(pkg, ctx, asm_oracle, case):
    """afb_dirichlet_set = applyDir on every Dirichlet dof of every cell (dc_on_dof.h:27-45), all product paths; accumulate
    and rhs-only semantics with constraints"""
    M = asm_oracle
    variables = {"p1": [(gc.P1, 1)], "th_blocks": [(gc.P2, 3), (gc.P1, 1)]}.get(case, [(gc.P2, 1)])
    co, te, dm = _mesh(pkg, ctx, M, (3, 3, 2) if case == "th_blocks" else (4, 3, 3), variables)
    if case == "th_blocks":
        _, forms, rhsf, prob = problems.c5_stokes(pkg, M, co, te)
    elif case == "p1":
        _, forms, rhsf, prob = problems.c1_p1_diffusion(pkg, M, co, te)
    else:
        _, forms, rhsf, prob = problems.c2_p2_aniso(pkg, M, co, te)
    rng = np.random.default_rng(9)
    ndof = dm.nrows
    flag = (rng.random(ndof) < 0.3).astype(np.uint8)
    if case == "th_blocks":
        flag[-(co.shape[0]):] = 0   # pressure dofs stay free (velocity Dirichlet only, stokes.cpp)
    value = rng.standard_normal(ndof)
    env = {"AFB_DISABLE_TENSOR_PATH": "1"} if case == "p2_generic" else {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        nnz = ctx.pattern_build()
        rowptr, colind = ctx.pattern_get()
        ctx.dirichlet_set(flag, value)
        val, rhs = np.full(nnz, np.nan), np.full(ndof, np.nan)
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
        rp, ci, v, r, st = M.assemble(prob, co, te, dm, dirichlet=(flag, value))
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
        rowmax[rowmax == 0] = 1.0
        # the rhs of a free row collects A_rc bc_c: scale with |A| |bc|
        rscale = max(np.abs(r).max(), 1.0)
        assert (np.abs(val - v) / np.repeat(rowmax, np.diff(rp))).max() <= RTOL
        assert np.abs(rhs - r).max() <= RTOL * rscale * 10
        # Dirichlet rows: deg on the diagonal, deg * bc in the rhs
        rows = np.repeat(np.arange(ndof), np.diff(rp))
        drow = flag[rows].astype(bool)
        assert (val[drow & (rows != ci)] == 0).all() and (val[drow & (rows == ci)] >= 1).all()
        # accumulate adds the constrained contribution once more; rhs-only needs no matrix output
        v2, r2 = val.copy(), rhs.copy()
        assert ctx.assemble(forms, rhsf, v2, r2, accumulate=True) == 0
        assert np.abs(v2 - 2 * val).max() <= 1e-14 * np.abs(val).max() and np.abs(r2 - 2 * rhs).max() <= 1e-14 * rscale
        r3 = np.full(ndof, np.nan)
        assert ctx.assemble(forms, rhsf, None, r3) == 0
        assert np.abs(r3 - rhs).max() <= 1e-14 * rscale
        # cleared again: unconstrained system
        ctx.dirichlet_set(None, None)
        v4 = np.zeros(nnz)
        assert ctx.assemble(forms, [], v4, None) == 0
        rp, ci, v0, r0, st = M.assemble(prob, co, te, dm)
        assert (np.abs(v4 - v0) / np.repeat(np.maximum(np.maximum.reduceat(np.abs(v0), rp[:-1]), 1e-300), np.diff(rp))).max() <= RTOL
    finally:
        ctx.dirichlet_set(None, None)
        for k, vv in old.items():
            if vv is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = vv


def _boundary_faces(co, te, sides):
    """(tet, face number) of every face whose three vertices lie on one of the `sides` (predicates on coordinates);
    face k = vertices k, k+1, k+2 mod 4"""
    ft, fn = [], []
    for k in range(4):
        idx = [(k + m) % 4 for m in range(3)]
        hit = np.zeros(te.shape[0], dtype=bool)
        for side in sides:
            hit |= side(co[te[:, idx[0]]]) & side(co[te[:, idx[1]]]) & side(co[te[:, idx[2]]])
        sel = np.nonzero(hit)[0]
        ft.append(sel); fn.append(np.full(sel.size, k))
    ft, fn = np.concatenate(ft), np.concatenate(fn)
    order = np.lexsort((fn, ft))
    return ft[order].astype(np.int32), fn[order].astype(np.int32)


@pytest.mark.parametrize("case", ["p2_robin_neumann", "p1_with_dirichlet", "p2vec_traction", "p3_generic"])
def test_boundary_face_terms(pkg, ctx, asm_oracle, case):
    """afb_boundary_set + afb_assemble_faces = the fem3Dface terms the reference's local assemblers add on labelled boundary
    faces (examples/Fem/Ani/diffusion.cpp:215-245): Robin matrix + Neumann load on two sides of the cube, with and without
    essential conditions on another side, scalar and vector spaces, on top of the fused and the generic volume paths"""
    M = asm_oracle
    variables = {"p1_with_dirichlet": [(gc.P1, 1)], "p2vec_traction": [(gc.P2, 3)], "p3_generic": [(gc.P3, 1)]}.get(case, [(gc.P2, 1)])
    fem, vec = variables[0]
    n = (3, 2, 2) if case in ("p2vec_traction", "p3_generic") else (4, 3, 3)
    co, te, dm = _mesh(pkg, ctx, M, n, variables, jitter=0.0)
    if case == "p2vec_traction":
        _, forms, rhsf, prob = problems.c4_p2_elasticity(pkg, M, co, te)
    elif case == "p1_with_dirichlet":
        _, forms, rhsf, prob = problems.c1_p1_diffusion(pkg, M, co, te)
    else:
        rng0 = np.random.default_rng(3)
        K = gc.tensor(rng0, gc.T_SYMMETRIC, gc.L_PER_TET, 3, 3, te.shape[0], 4)
        _, forms, rhsf, prob = problems._mk(pkg, M, variables, [(0, 0, gc.GRAD, gc.GRAD, 2 if fem != gc.P3 else 4, gc.T_SYMMETRIC, gc.L_PER_TET, K, 1.0)],
                                            [(0, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)])
    ft, fn = _boundary_faces(co, te, [lambda X: np.abs(X[:, 0] - 1.0) < 1e-12, lambda X: np.abs(X[:, 2]) < 1e-12])
    nbf = ft.size
    assert nbf == 2 * (n[1] * n[2] + n[0] * n[1])
    rng = np.random.default_rng(17)
    order = 4
    q = gc.NPTS_TRI[order]
    if vec == 1:
        alpha = gc.tensor(rng, gc.T_SCALAR, gc.L_PER_TET, 1, 1, nbf, q)        # Robin coefficient per face
        gN = gc.tensor(rng, gc.T_SCALAR, gc.L_PER_POINT, 1, 1, nbf, q).reshape(nbf, q)   # Neumann flux per quadrature point of every face
        fmats = [(0, 0, gc.IDEN, gc.IDEN, order, gc.T_SCALAR, gc.L_PER_TET, alpha, 1.0)]
        frhss = [(0, gc.IDEN, order, gc.T_SCALAR, gc.L_PER_POINT, gN, 1.0)]
    else:
        R = gc.tensor(rng, gc.T_SYMMETRIC, gc.L_CONST, 3, 3, nbf, q)             # elastic foundation
        t = np.ascontiguousarray(rng.standard_normal((nbf, 3)))                  # traction per face
        fmats = [(0, 0, gc.IDEN, gc.IDEN, order, gc.T_SYMMETRIC, gc.L_CONST, R, 0.5)]
        frhss = [(0, gc.IDEN, order, gc.T_GENERAL, gc.L_PER_TET, t, 1.0)]
    _, fforms, frhsf, fprob = problems._mk(pkg, M, variables, fmats, frhss)
    dirichlet = None
    if case == "p1_with_dirichlet":
        flag = np.zeros(dm.nrows, dtype=np.uint8)
        flag[np.nonzero(np.abs(co[:, 1]) < 1e-12)[0]] = 1      # P1: dof = node (NATURAL numbering); y = 0 side, shares edges with x = 1 and z = 0
        value = rng.standard_normal(dm.nrows)
        dirichlet = (flag, value)
    env = {"AFB_DISABLE_TENSOR_PATH": "1"} if case == "p3_generic" else {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        nnz = ctx.pattern_build()
        rowptr, colind = ctx.pattern_get()
        if dirichlet is not None:
            ctx.dirichlet_set(*dirichlet)
        val, rhs = np.full(nnz, np.nan), np.full(dm.nrows, np.nan)
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
        ctx.boundary_set(ft, fn)
        assert ctx.assemble_faces(fforms, frhsf, val, rhs) == 0
        rp, ci, v, r, st = M.assemble(prob, co, te, dm, dirichlet=dirichlet, faces=(ft, fn, fprob))
        assert st == 0 and np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
        rowmax[rowmax == 0] = 1.0
        err = (np.abs(val - v) / np.repeat(rowmax, np.diff(rp))).max()
        rerr = np.abs(rhs - r).max() / max(np.abs(r).max(), 1.0)
        print("%s: %d boundary faces, err A %.2e rhs %.2e" % (case, nbf, err, rerr))
        assert err <= RTOL and rerr <= 10 * RTOL
        # the face terms really are in there: without them the matrix differs
        v0 = M.assemble(prob, co, te, dm, dirichlet=dirichlet)[2]
        assert np.abs(v - v0).max() > 1e-3 * np.abs(v).max()
        # bit-reproducible, and device buffers give the same numbers as host buffers
        val2, rhs2 = np.full(nnz, np.nan), np.full(dm.nrows, np.nan)
        assert ctx.assemble(forms, rhsf, val2, rhs2) == 0 and ctx.assemble_faces(fforms, frhsf, val2, rhs2) == 0
        assert np.array_equal(val, val2) and np.array_equal(rhs, rhs2)
    finally:
        ctx.dirichlet_set(None, None)
        ctx.boundary_set(np.zeros(0, np.int32), np.zeros(0, np.int32))
        for k, vv in old.items():
            if vv is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = vv


@pytest.mark.parametrize("problem", ["c3", "c5"])
def test_generic_path_in_element_chunks(pkg, ctx, asm_oracle, oracle, problem):
    """the generic staged path bounds its staging memory (AFB_STAGE_BYTES): many element chunks give the same matrix up to the
    re-association of the per-chunk partial sums"""
    M = asm_oracle
    variables = [(gc.P3, 1)] if problem == "c3" else [(gc.P2, 3), (gc.P1, 1)]
    co, te, dm = _mesh(pkg, ctx, M, (3, 2, 2), variables)
    if problem == "c3":
        XY = co[te].transpose(1, 0, 2)
        _, forms, rhsf, prob = problems.c3_p3_react_diff(pkg, M, co, te, oracle.quad_points(4, XY), oracle.quad_points(6, XY))
    else:
        _, forms, rhsf, prob = problems.c5_stokes(pkg, M, co, te)
    a, fa, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, problem + " one chunk", {"AFB_DISABLE_TENSOR_PATH": "1"})
    assert not path["fused_path"] and path["gather_kernel"].startswith("k_gather")
    b, fb, _ = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, problem + " chunks of ~7 elements",
                               {"AFB_DISABLE_TENSOR_PATH": "1", "AFB_STAGE_BYTES": str(7 * dm.nloc * (dm.nloc + 1) * 8)})
    assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max() and np.abs(fa - fb).max() <= 1e-14 * max(np.abs(fa).max(), 1.0)
    assert not np.isnan(b).any()


def test_properties_at_scale(pkg, asm_oracle):
    """C2 at 48^3 x 6 = 663,552 tets (too large for the numpy oracle): properties that do not need one"""
    c = pkg.Context(0)
    n = 48
    c.mesh_cube(n, n, n)
    c.dofmap_natural([(gc.P2, 1)])
    nnz = c.pattern_build()
    rowptr, colind = c.pattern_get()
    coords, tets = c.mesh_get()
    _, forms, rhsf, _ = problems.c2_p2_aniso(pkg, None, coords, tets)
    nrows = rowptr.size - 1
    a, fa = np.zeros(nnz), np.zeros(nrows)
    os.environ["AFB_DISABLE_RING_KERNEL"] = "1"   # this test is about the row gather; tests/test_rings_gpu.py has the ring kernel's twin
    try:
        assert c.assemble(forms, rhsf, a, fa) == 0
    finally:
        os.environ.pop("AFB_DISABLE_RING_KERNEL")
    assert c.last_times()["gather_kernel"] == "k_rows_cl"
    os.environ["AFB_DISABLE_RING_KERNEL"] = "1"
    b, fb = np.zeros(nnz), np.zeros(nrows)
    try:
        assert c.assemble(forms, rhsf, b, fb) == 0
    finally:
        os.environ.pop("AFB_DISABLE_RING_KERNEL")
    assert np.array_equal(a, b) and np.array_equal(fa, fb), "not bit-reproducible"
    rows = np.repeat(np.arange(nrows), np.diff(rowptr))
    scale = np.abs(a).max()
    # stiffness rows annihilate constants; the load vector integrates 1 over the unit cube
    assert np.abs(np.bincount(rows, weights=a, minlength=nrows)).max() <= 1e-12 * scale * 70
    assert abs(fa.sum() - 1.0) <= 1e-12
    # symmetric tensor => symmetric matrix: compare with the transpose through a sort of (col,row) keys
    key = colind.astype(np.int64) * nrows + rows
    perm = np.argsort(key, kind="stable")
    assert np.array_equal(rows[perm], colind) and np.array_equal(colind[perm], rows), "pattern not structurally symmetric"
    assert np.abs(a[perm] - a).max() <= 1e-12 * scale
    # columns ascending inside every row, diagonal present
    d = np.diff(colind.astype(np.int64))
    starts = rowptr[1:-1] - 1
    d[starts] = 1
    assert (d > 0).all()
    assert np.isin(np.arange(nrows, dtype=np.int64) * nrows + np.arange(nrows), rows * np.int64(nrows) + colind).all()
    c.close()
