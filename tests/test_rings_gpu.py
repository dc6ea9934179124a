"""GPU parity tests of the ring-traversal kernel k_rings (inmost-fem_b200/csrc/afb_rings.cu): square P2 problems with one
symmetric stiffness form (+ mass, + load) through the C ABI against the CPU oracle -- CSR pattern bit-exact, values / rhs within
1e-12 of the row scale.  Cases: several clusters per mesh (AFB_RING_EDGES), jittered and element-shuffled meshes (open and closed
rings, every frame), accumulate, drop_val, NaN status, bit-reproducibility, phased assembly, and the size-independent properties
of the stiffness matrix at a larger size.  The plan logic itself is covered on the CPU by tests/cxx/test_ring_plan.cpp."""
import os

import numpy as np
import pytest

import golden_cases as gc
import problems
from test_fused_paths_gpu import _mesh, _oracle_compare

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _forms(pkg, M, te, rng, mass=True, load=True, const=False):
    nt = te.shape[0]
    lay = gc.L_CONST if const else gc.L_PER_TET
    K = gc.tensor(rng, gc.T_SYMMETRIC, lay, 3, 3, nt, 4)
    c = gc.tensor(rng, gc.T_SCALAR, lay, 1, 1, nt, 4)
    mats = [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, lay, K, 1.0)]
    if mass:
        mats.append((0, 0, gc.IDEN, gc.IDEN, 3, gc.T_SCALAR, lay, c, 0.5))
    rhss = [(0, gc.IDEN, 2, gc.T_SCALAR, lay, c, 2.0)] if load else []
    return problems._mk(pkg, M, [(gc.P2, 1)], mats, rhss)


@pytest.mark.parametrize("edges", [32, 64, 256])
@pytest.mark.parametrize("variant", ["stiff", "stiff+mass+load", "stiff+load"])
def test_rings_many_clusters(pkg, ctx, asm_oracle, edges, variant):
    M = asm_oracle
    co, te, dm = _mesh(pkg, ctx, M, (5, 4, 4), [(gc.P2, 1)], jitter=0.08, seed=edges)
    rng = np.random.default_rng(edges)
    _, forms, rhsf, prob = _forms(pkg, M, te, rng, mass="mass" in variant, load="load" in variant)
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "rings %s, %d edges per cluster" % (variant, edges), {"AFB_RING_EDGES": str(edges)})
    assert path["gather_kernel"] == "k_rings"


def test_rings_shuffled_unstructured_mesh(pkg, ctx, asm_oracle):
    """element order shuffled, nodes jittered, tets re-oriented by the library: rings start anywhere, every local frame occurs"""
    M = asm_oracle
    rng = np.random.default_rng(7)
    co, te, _ = M.cube_mesh(4, 3, 3)
    co = co + 0.06 * rng.standard_normal(co.shape) / 4
    te = te[rng.permutation(te.shape[0])]
    te = np.stack([rng.permutation(row) for row in te])      # arbitrary local vertex order ...
    ctx.mesh_set(co, te)
    ctx.mesh_orient()                                         # ... made positive like ordering.inl:8-26
    _, te = ctx.mesh_get()
    ctx.dofmap_natural([(gc.P2, 1)])
    dm = M.DofMap(te, [(gc.P2, 1)], nnode=co.shape[0])
    _, forms, rhsf, prob = _forms(pkg, M, te, rng)
    _, _, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "rings shuffled mesh", {"AFB_RING_EDGES": "64"})
    assert path["gather_kernel"] == "k_rings"


def test_rings_accumulate_drop_nan_determinism(pkg, ctx, asm_oracle):
    M = asm_oracle
    co, te, dm = _mesh(pkg, ctx, M, (4, 4, 3), [(gc.P2, 1)])
    rng = np.random.default_rng(1)
    _, forms, rhsf, prob = _forms(pkg, M, te, rng)
    val, rhs, path = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "rings base")
    assert path["gather_kernel"] == "k_rings"
    nnz = val.size
    # accumulate: the assembly ADDS into the matrix (assembler.inl:305-306)
    v2, r2 = val.copy(), rhs.copy()
    assert ctx.assemble(forms, rhsf, v2, r2, accumulate=True) == 0
    assert np.array_equal(v2, 2 * val) and np.array_equal(r2, 2 * rhs)
    # bit-reproducible
    v3, r3 = np.zeros(nnz), np.zeros(rhs.size)
    assert ctx.assemble(forms, rhsf, v3, r3) == 0
    assert np.array_equal(v3, val) and np.array_equal(r3, rhs)
    # drop_val (assembler.inl:416): same rule as the oracle, entries at the threshold may flip
    drop = 0.02 * np.abs(val).max()
    v4 = np.zeros(nnz)
    assert ctx.assemble(forms, [], v4, None, drop_val=drop) == 0
    assert ctx.last_times()["gather_kernel"] == "k_rings"
    rp, ci, v, r, st = M.assemble(M.Problem(prob.vars, prob.mat_forms, []), co, te, dm, drop_val=drop)
    rowmax = np.maximum.reduceat(np.abs(v), rp[:-1])
    bad = np.abs(v4 - v) / np.repeat(rowmax, np.diff(rp)) > RTOL
    assert bad.mean() < 1e-3 and (np.abs(v4 - v)[bad] <= 1.001 * drop * 4).all()
    assert not np.array_equal(v4, val)
    # non-finite coefficient -> status -1 (assembler.inl:419-424)
    K = forms[0]._keep
    old = K[5, 4]
    K[5, 4] = np.nan
    assert ctx.assemble(forms, rhsf, np.zeros(nnz), np.zeros(rhs.size)) == -1
    K[5, 4] = np.inf
    assert ctx.assemble(forms, rhsf, np.zeros(nnz), np.zeros(rhs.size)) == -1
    K[5, 4] = old
    assert ctx.assemble(forms, rhsf, v3, r3) == 0
    assert np.array_equal(v3, val)


def test_rings_equal_row_gather(pkg, ctx, asm_oracle):
    """the two fused kernels (k_rings, k_rows_cl) agree to rounding on a constant-coefficient problem"""
    M = asm_oracle
    co, te, dm = _mesh(pkg, ctx, M, (6, 5, 4), [(gc.P2, 1)])
    rng = np.random.default_rng(2)
    _, forms, rhsf, prob = _forms(pkg, M, te, rng, const=True)
    a, ra, pa = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "const rings")
    b, rb, pb = _oracle_compare(ctx, M, prob, forms, rhsf, co, te, dm, "const rows", {"AFB_DISABLE_RING_KERNEL": "1"})
    assert pa["gather_kernel"] == "k_rings" and pb["gather_kernel"] == "k_rows_cl"
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max() and np.abs(ra - rb).max() <= 1e-13 * np.abs(ra).max()


def test_rings_phased_assembly(pkg, asm_oracle):
    """afb_assemble_phase (the clusters writing to priority rows first, then the rest) gives the bits of the one-shot assembly"""
    import torch
    M = asm_oracle
    ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    ctx.mesh_cube(6, 5, 5)
    ctx.dofmap_natural([(gc.P2, 1)])
    nnz = ctx.pattern_build()
    co, te = ctx.mesh_get()
    rng = np.random.default_rng(3)
    _, forms, rhsf, prob = _forms(pkg, M, te, rng)
    nrows = ctx.dofmap_info()[3]
    ref_v = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    ref_r = torch.zeros(nrows, dtype=torch.float64, device="cuda")
    os.environ["AFB_RING_EDGES"] = "64"   # the plan is built by the first assembly that uses it
    try:
        assert ctx.assemble(forms, rhsf, ref_v, ref_r) == 0
    finally:
        os.environ.pop("AFB_RING_EDGES")
    assert ctx.last_times()["gather_kernel"] == "k_rings"
    for first in (nrows - 37, nrows // 2, 5):
        ctx.priority_rows_set(first)
        v = torch.full((nnz,), float("nan"), dtype=torch.float64, device="cuda")
        r = torch.full((nrows,), float("nan"), dtype=torch.float64, device="cuda")
        ctx.assemble_phase(forms, rhsf, v, r, 1)
        torch.cuda.synchronize()
        rp, _ = ctx.pattern_get()
        lo = int(rp[first])
        assert torch.equal(v[lo:], ref_v[lo:]) and torch.equal(r[first:], ref_r[first:]), "priority rows incomplete after phase 1"
        assert ctx.assemble_phase(forms, rhsf, v, r, 2) == 0
        torch.cuda.synchronize()
        assert torch.equal(v, ref_v) and torch.equal(r, ref_r)
    ctx.close()


def test_rings_properties_at_scale(pkg, asm_oracle):
    """40^3 x 6 = 384,000 tets: symmetric matrix, constants in the null space of the stiffness rows, load sums to the volume"""
    ctx = pkg.Context(0)
    n = 40
    ctx.mesh_cube(n, n, n)
    ctx.dofmap_natural([(gc.P2, 1)])
    nnz = ctx.pattern_build()
    rowptr, colind = ctx.pattern_get()
    co, te = ctx.mesh_get()
    K = problems.sym_K(co[te].mean(axis=1))
    forms = [pkg.make_form(gc.GRAD, gc.P2, 1, gc.GRAD, gc.P2, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K)]
    rhsf = [pkg.make_form(gc.IDEN, gc.P0, 1, gc.IDEN, gc.P2, 1, 2, gc.T_NULL, gc.L_CONST)]
    val, rhs = np.zeros(nnz), np.zeros(rowptr.size - 1)
    assert ctx.assemble(forms, rhsf, val, rhs) == 0
    assert ctx.last_times()["gather_kernel"] == "k_rings"
    import scipy.sparse as sp
    A = sp.csr_matrix((val, colind, rowptr), shape=(rowptr.size - 1,) * 2)
    scale = np.abs(val).max()
    assert abs(A - A.T).max() <= 1e-12 * scale
    assert np.abs(A @ np.ones(A.shape[0])).max() <= 1e-11 * scale
    assert abs(rhs.sum() - 1.0) <= 1e-12
    # against the row gather on the same data
    os.environ["AFB_DISABLE_RING_KERNEL"] = "1"
    try:
        v2, r2 = np.zeros(nnz), np.zeros(rhs.size)
        assert ctx.assemble(forms, rhsf, v2, r2) == 0
        assert ctx.last_times()["gather_kernel"] == "k_rows_cl"
    finally:
        os.environ.pop("AFB_DISABLE_RING_KERNEL")
    rowmax = np.maximum.reduceat(np.abs(v2), rowptr[:-1])
    assert (np.abs(val - v2) / np.repeat(rowmax, np.diff(rowptr))).max() <= RTOL
    assert np.abs(rhs - r2).max() <= RTOL * np.abs(r2).max()
    ctx.close()


def test_rings_superset_pattern(pkg, ctx, asm_oracle):
    """a user pattern with columns no local element contributes to (what the union pattern of a partitioned mesh looks like,
    afb_pattern_set): those entries come out as zeros, with accumulate they stay untouched"""
    M = asm_oracle
    co, te, dm = _mesh(pkg, ctx, M, (5, 4, 3), [(gc.P2, 1)], jitter=0.05, seed=4)
    rng = np.random.default_rng(4)
    _, forms, rhsf, prob = _forms(pkg, M, te, rng)
    rp, ci, v, r, st = M.assemble(prob, co, te, dm)
    nrows = rp.size - 1
    rows = np.repeat(np.arange(nrows), np.diff(rp))
    extra_r = rng.integers(0, nrows, 4000)
    extra_c = rng.integers(0, nrows, 4000)
    key = np.unique(np.concatenate([rows * nrows + ci.astype(np.int64), extra_r * nrows + extra_c]))
    rp2 = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(rp2, key // nrows + 1, 1)
    rp2 = np.cumsum(rp2)
    ci2 = (key % nrows).astype(np.int32)
    exp = np.zeros(key.size)
    exp[np.searchsorted(key, rows * nrows + ci.astype(np.int64))] = v
    ctx.pattern_build()
    ctx.pattern_set(rp2, ci2)
    os.environ["AFB_RING_EDGES"] = "64"
    try:
        val, rhs = np.full(key.size, np.nan), np.full(nrows, np.nan)
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
    finally:
        os.environ.pop("AFB_RING_EDGES")
    assert ctx.last_times()["gather_kernel"] == "k_rings"
    scale = np.abs(v).max()
    assert np.abs(val - exp).max() <= RTOL * scale and np.abs(rhs - r).max() <= RTOL * np.abs(r).max()
    assert (val[exp == 0] == 0).all()
    val2 = np.full(key.size, 0.25)
    assert ctx.assemble(forms, [], val2, None, accumulate=True) == 0
    assert np.abs(val2 - 0.25 - exp).max() <= RTOL * scale and (val2[exp == 0] == 0.25).all()
