"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's global-assembly path
(anifem++/inmost_interface + the cube generator of anifem++/utils).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs -- never by the product.

PARITY STATUS: **pinned** on one rank since round 2, against the reference's OWN assembler: the unmodified
inmost_interface/{global_enumerator.cpp, elemental_assembler.cpp, assembler.inl, ordering.inl} compile on top of
oracle/mock_inmost/inmost.h (the bounded INMOST surface they use; INMOST itself is an un-vendored external pinned at
INMOST-DEV/INMOST@f3392cef4cbb4d05b91c5cc94922040637bea4ee, cmake/Downloadinmost.cmake:4) into
oracle/_ref/libanifem_refasm.so (oracle/Makefile target refasm, driver oracle/ref_asm_driver.cpp).  Its outputs on cubes and a
scrambled mesh -- numbering of all six enumerators, AssembleTemplate pattern, Assemble values / drop rule / status -- are
committed as tests/golden/ref_assembler.npz (generator tests/golden/make_golden_asm.py) and tests/test_oracle_golden.py checks
this module against them (indices bit-exact, values 1e-13) and against the live build when present.
Still ours and NOT pinned (INMOST's parallel mesh cannot be built here): what INMOST decides on SEVERAL ranks -- entity
ownership and the rank-major GlobalIDs -- follows the documented conventions below; everything AniFem++ itself decides is
restated from its sources:

  mesh          utils/mesh_utils.cpp:20-48 (6 tets per hex), :110-145 (node/hex loops)
  orientation   inmost_interface/ordering.inl:8-26      (swap nodes 2,3 if det<0)
  local edges   01,02,03,12,13,23; local face i = nodes (i,i+1,i+2)%4   (ordering.inl:84-116,
                fem/fem_space.h:27-69)
  local dofs    vertices, edges, faces, cell; vector = component-major; variables in order
                (fem/tetdofmap.cpp:327-339,433-440,606-643)
  P3 edge pair  slot flips with the global order of the edge's endpoints (tetdofmap.inl:98-104,
                assembler.inl:160-164)
  numbering     NATURAL = lexicographic (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID), per-rank contiguous
                interval [BegInd,EndInd)  (global_enumerator.cpp:562-605, :702-777, :866-890)
  index codes   sign*(id+1), 0 = row skipped (ghost)           (assembler.inl:49-55, :139-184)
  pattern       AssembleTemplate: sorted rows, forced diagonal  (assembler.inl:114-136, :589-695)
  values        matrix[r][c] += s_r s_c A(i,j) if |A|>drop_val; rhs[r] += s_r F[i]   (:397-425)

Conventions replacing INMOST (ours, documented in DESIGN.md):
  * cell->node order = order of first appearance in the face lists of CreateNWTetElements;
  * canonical entity ids: nodes in (i,j,k) creation order; edges / faces numbered in
    lexicographic order of their sorted node tuples;
  * with R ranks: a cell belongs to the box block of mesh_utils.cpp:67-108; an entity is owned by
    the lowest rank among its adjacent cells; GlobalID = BegElemID[owner] + position among the
    owner's entities in canonical order.
"""
import numpy as np

from . import oracle as O

# nodes of the six tets of a hex, first-appearance order of the face lists (mesh_utils.cpp:22-40)
HEX_TETS = np.array([[0, 1, 5, 3], [0, 3, 5, 7], [0, 7, 5, 4], [0, 3, 7, 2], [0, 7, 4, 2], [4, 6, 2, 7]])
LOCAL_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])
LOCAL_FACES = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 0], [3, 0, 1]])

# dofs per (node, edge, face, cell) of the scalar spaces (spaces/poly_*.h Dof<>::Map())
NDOF = {O.P0: (0, 0, 0, 1), O.P1: (1, 0, 0, 0), O.P2: (1, 1, 0, 0), O.P3: (1, 2, 1, 0)}


def proc_grid(nranks, sizes):
    """process grid of GenerateParallelepiped (mesh_utils.cpp:67-86)"""
    divs, d = [], nranks
    while d > 1:
        for k in range(2, d + 1):
            if d % k == 0:
                divs.append(k)
                d //= k
                break
    ppa = [1, 1, 1]
    epp = list(sizes)
    for k in reversed(divs):
        m = int(np.argmax(epp))  # std::max_element returns the first maximum
        ppa[m] *= k
        epp[m] //= k
    return ppa


def cube_mesh(nx, ny, nz, size=1.0, nranks=1):
    """coords (nnode,3), tets (ntet,4) positively oriented, cell_rank (ntet,)"""
    ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    coords = np.stack([ii.ravel() * (size - 0.0) / nx + 0.0, jj.ravel() * (size - 0.0) / ny + 0.0,
                       kk.ravel() * (size - 0.0) / nz + 0.0], axis=1)
    vid = lambda i, j, k: (i * (ny + 1) + j) * (nz + 1) + k
    hi, hj, hk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    # hex vertex v: bit0 -> +x, bit1 -> +y, bit2 -> +z  (mesh_utils.cpp:133-140)
    hv = np.stack([vid(hi + (v & 1), hj + ((v >> 1) & 1), hk + ((v >> 2) & 1)) for v in range(8)], axis=1)
    tets = hv[:, HEX_TETS].reshape(-1, 4)
    # ordering.inl:8-26
    p = coords[tets]
    m = p[:, :3, :] - p[:, 3:4, :]
    det = np.linalg.det(m)
    neg = det < 0
    tets[neg, 2], tets[neg, 3] = tets[neg, 3].copy(), tets[neg, 2].copy()
    # box blocks (mesh_utils.cpp:88-108)
    ppa = proc_grid(nranks, (nx, ny, nz))
    avg = [int(np.ceil(s / p_)) for s, p_ in zip((nx, ny, nz), ppa)]
    pc = [np.minimum(h // a, p_ - 1) for h, a, p_ in zip((hi, hj, hk), avg, ppa)]
    hrank = pc[2] * ppa[0] * ppa[1] + pc[1] * ppa[0] + pc[0]
    return coords, tets.astype(np.int64), np.repeat(hrank, 6)


def _unique_rows(keys):
    """ids of rows in lexicographic order of the rows"""
    u, inv = np.unique(keys, axis=0, return_inverse=True)
    return u, inv.reshape(-1)


def connectivity(tets):
    """edges (nedge,2), faces (nface,3) as sorted node tuples in canonical order,
    tet_edges (ntet,6), tet_faces (ntet,4)"""
    e = np.sort(tets[:, LOCAL_EDGES], axis=2).reshape(-1, 2)
    edges, inv = _unique_rows(e)
    f = np.sort(tets[:, LOCAL_FACES], axis=2).reshape(-1, 3)
    faces, finv = _unique_rows(f)
    return edges, faces, inv.reshape(-1, 6), finv.reshape(-1, 4)


class DofMap:
    """elem->dof tables for a list of variables [(fem, vecdim), ...]; NATURAL numbering unless enum_type names another
    GlobEnumeration type (then the index inside every rank's interval is the rank of the dof's tuple among the dofs the rank
    owns, restated from the definitions like enumerate_dofs; ELEM_ID = GlobalID - BegElemID of the rank)."""

    def __init__(self, tets, variables, cell_rank=None, nranks=1, nnode=None, enum_type="NATURAL"):
        self.tets = tets
        self.vars = list(variables)
        ntet = tets.shape[0]
        nnode = int(tets.max()) + 1 if nnode is None else nnode
        cell_rank = np.zeros(ntet, dtype=np.int64) if cell_rank is None else cell_rank
        self.cell_rank, self.nranks = cell_rank, nranks
        edges, faces, te, tf = connectivity(tets)
        ent_of_tet = [tets, te, tf, np.arange(ntet)[:, None]]
        nent = [nnode, edges.shape[0], faces.shape[0], ntet]
        # ownership: lowest rank among adjacent cells; GlobalID contiguous per owner
        owner, gid, beg, num = [], [], [], []
        for d in range(4):
            ow = np.full(nent[d], nranks, dtype=np.int64)
            np.minimum.at(ow, ent_of_tet[d].ravel(), np.repeat(cell_rank, ent_of_tet[d].shape[1]))
            ow[ow == nranks] = 0  # isolated entity (cannot happen on the cube)
            order = np.argsort(ow, kind="stable")  # canonical order inside each owner
            g = np.empty(nent[d], dtype=np.int64)
            g[order] = np.arange(nent[d])
            cnt = np.bincount(ow, minlength=nranks)
            owner.append(ow); gid.append(g); num.append(cnt); beg.append(np.concatenate([[0], np.cumsum(cnt)[:-1]]))
        self.owner, self.gid, self.ent_of_tet, self.nent = owner, gid, ent_of_tet, nent
        ndof_ent = np.zeros(4, dtype=np.int64)  # dofs per entity summed over vars and components
        for fem, vec in self.vars:
            ndof_ent += np.array(NDOF[fem]) * vec
        # per-rank interval (global_enumerator.cpp:594-604)
        self.beg_ind = np.array([sum(beg[d][r] * ndof_ent[d] for d in range(4)) for r in range(nranks)])
        self.end_ind = np.array([self.beg_ind[r] + sum(num[d][r] * ndof_ent[d] for d in range(4)) for r in range(nranks)])
        self.nrows = int(sum(nent[d] * ndof_ent[d] for d in range(4)))
        # NATURAL: offsets of the (var, dim, etype) groups inside each rank's interval
        cols, signs_unused = [], None
        grp_off = {}
        off = np.zeros(nranks, dtype=np.int64)
        for v, (fem, vec) in enumerate(self.vars):
            for c in range(vec):
                for d in range(4):
                    nd = NDOF[fem][d]
                    if nd == 0:
                        continue
                    grp_off[(v, c, d)] = off.copy()
                    off = off + num[d] * nd
        self.grp_off = grp_off
        # element -> global dof, local order: var-major, component-major, vertices/edges/faces/cell
        gnode = gid[0][tets]  # global node ids (for the P3 edge-pair orientation)
        for v, (fem, vec) in enumerate(self.vars):
            for c in range(vec):
                for d in range(4):
                    nd = NDOF[fem][d]
                    if nd == 0:
                        continue
                    ents = ent_of_tet[d]  # (ntet, nloc_ent)
                    ow = owner[d][ents]
                    base = self.beg_ind[ow] + grp_off[(v, c, d)][ow] + (gid[d][ents] - beg[d][ow]) * nd
                    for le in range(ents.shape[1]):
                        if d == 1 and nd == 2:  # S2 pair on an edge
                            a, b = LOCAL_EDGES[le]
                            flip = (gnode[:, a] > gnode[:, b]).astype(np.int64)
                            cols.append(base[:, le] + flip)
                            cols.append(base[:, le] + 1 - flip)
                        else:
                            for k in range(nd):
                                cols.append(base[:, le] + k)
        self.elem2dof = np.stack(cols, axis=1)  # (ntet, nloc) global ids
        self.nloc = self.elem2dof.shape[1]
        if enum_type != "NATURAL":
            self._renumber(enum_type, owner, gid, beg, num, nent)
        # owner rank of every local dof's entity (for row codes)
        owc = []
        for v, (fem, vec) in enumerate(self.vars):
            for c in range(vec):
                for d in range(4):
                    nd = NDOF[fem][d]
                    for le in range(ent_of_tet[d].shape[1] if nd else 0):
                        for k in range(nd):
                            owc.append(owner[d][ent_of_tet[d][:, le]])
        self.dof_owner = np.stack(owc, axis=1)

    def _renumber(self, enum_type, owner, gid, beg, num, nent):
        """replace the NATURAL ids by those of enum_type: old NATURAL id -> new id, rank by rank, from the definitions"""
        new_of_old = np.full(self.nrows, -1, dtype=np.int64)
        col = {"VAR": 0, "DIM": 1, "ELEM_TYPE": 2, "ELEM_ID": 3, "DOF_ID": 4}
        for r in range(self.nranks):
            recs, old = [], []
            for v, (fem, vec) in enumerate(self.vars):
                for c in range(vec):
                    for d in range(4):
                        nd = NDOF[fem][d]
                        ents = np.nonzero(owner[d] == r)[0]
                        if nd == 0 or ents.size == 0:
                            continue
                        g = gid[d][ents] - beg[d][r]
                        for k in range(nd):
                            recs.append(np.stack([np.full(g.size, v), np.full(g.size, c), np.full(g.size, d), g, np.full(g.size, k)], 1))
                            old.append(self.beg_ind[r] + self.grp_off[(v, c, d)][r] + g * nd + k)
            if not recs:
                continue
            recs, old = np.concatenate(recs, 0), np.concatenate(old)
            n = recs.shape[0]
            if enum_type in _ARRANGEMENT:
                keys = [recs[:, col[name]] for name in _ARRANGEMENT[enum_type]]
                order = np.lexsort(tuple(reversed(keys)))
                loc = np.empty(n, dtype=np.int64)
                loc[order] = np.arange(n)
            else:
                i_nd = [sum(NDOF[fem][d] * vec for fem, vec in self.vars) for d in range(4)]
                cnt = [int(num[d][r]) for d in range(4)]
                init = np.concatenate([[0], np.cumsum([i_nd[d] * cnt[d] for d in range(4)])])
                shift = {}
                for d in range(4):
                    o = 0
                    for v, (fem, vec) in enumerate(self.vars):
                        shift[(v, d)] = o
                        o += NDOF[fem][d] * vec
                vecs = np.array([vec for _, vec in self.vars])
                # position of a dof among all dofs of its entity = GetElemDofId (global_enumerator.h:76, .cpp:262-274): variables in
                # order, inside a vector variable component-major (dim_dof_shift = nd * dim_id, then the dof of the component)
                ns_rec = np.array([NDOF[self.vars[int(a)][0]][int(b)] for a, b in recs[:, [0, 2]]])
                iodf = np.array([shift[(int(a), int(b))] for a, b in recs[:, [0, 2]]]) + recs[:, 1] * ns_rec + recs[:, 4]
                d = recs[:, 2]
                if enum_type == "ANITYPE":
                    loc = init[d] + recs[:, 3] + iodf * np.array(cnt)[d]
                elif enum_type == "MINIBLOCKS":
                    loc = init[d] + recs[:, 3] * np.array(i_nd)[d] + iodf
                else:
                    raise ValueError("unknown enumeration type " + str(enum_type))
            assert np.array_equal(np.sort(loc), np.arange(n))
            new_of_old[old] = self.beg_ind[r] + loc
        assert (new_of_old >= 0).all()
        self.elem2dof = new_of_old[self.elem2dof]

    def codes(self, rank=None):
        """(rowcode, colcode) as assemble_index_encode: sign*(id+1); rows of entities not owned by
        `rank` are CODE_UNDEF=0 (assembler.inl:174-181). rank=None: every row active."""
        col = self.elem2dof + 1
        row = col.copy()
        if rank is not None:
            row[self.dof_owner != rank] = 0
        return row, col


# GlobEnumeration types (inmost_interface/global_enumerator.h:393-401) and the arrangement of the OrderedEnumerator tuple
# (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID) each of them sets (global_enumerator.cpp:818-835)
ENUM_TYPES = ("ANITYPE", "MINIBLOCKS", "NATURAL", "DIMUNION", "BYELEMTYPE", "ETDIMBLOCKS")
_ARRANGEMENT = {"NATURAL": ("VAR", "DIM", "ELEM_TYPE", "ELEM_ID", "DOF_ID"),
                "DIMUNION": ("VAR", "ELEM_TYPE", "ELEM_ID", "DOF_ID", "DIM"),
                "BYELEMTYPE": ("ELEM_TYPE", "VAR", "DIM", "ELEM_ID", "DOF_ID"),
                "ETDIMBLOCKS": ("ELEM_TYPE", "VAR", "ELEM_ID", "DOF_ID", "DIM")}


def enumerate_dofs(tets, variables, enum_type="NATURAL", nnode=None):
    """elem -> global dof table (ntet, nloc) and the number of dofs on ONE rank for every GlobEnumeration type, restated from
    the definitions, not from closed forms:
      * OrderedEnumerator (NATURAL, DIMUNION, BYELEMTYPE, ETDIMBLOCKS): the index of a dof is the rank of its tuple
        (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID) in the lexicographic order of the type's arrangement
        (global_enumerator.cpp:702-777: loc_emap over the tuples; DIM = di % ndim, DOF_ID = di / ndim for the di-th dof of a
        vector variable on an entity, :727-731);
      * SimpleEnumerator ANITYPE: InitElemIndex[edim] + (gid - BegElemID) + iodf * NumElem[edim] (:823-825), iodf = position of
        the dof among all dofs on the entity: variables in order, component fastest inside a vector variable;
      * SimpleEnumerator MINIBLOCKS: the layout its own inverse map decodes (:866-871): InitElemIndex[edim] + gid * nd[edim] +
        iodf.  (The forward formula at :826 multiplies iodf by NumElem[edim], which leaves the valid range when an entity
        carries more than one dof; the inverse and the name say dofs of one entity are adjacent.)
    Entity ids as everywhere in this oracle: nodes by id, edges / faces in lexicographic order of their sorted node tuples.
    Local order on the tet as in DofMap (variable, component, vertices, edges [P3 pair oriented by node ids], faces, cell)."""
    ntet = tets.shape[0]
    nnode = int(tets.max()) + 1 if nnode is None else nnode
    edges, faces, te, tf = connectivity(tets)
    ent_of_tet = [tets, te, tf, np.arange(ntet)[:, None]]
    nent = [nnode, edges.shape[0], faces.shape[0], ntet]
    # every dof as a record (var, comp, dim-type d, entity, k)
    recs = []
    for v, (fem, vec) in enumerate(variables):
        for c in range(vec):
            for d in range(4):
                for k in range(NDOF[fem][d]):
                    g = np.arange(nent[d])
                    recs.append(np.stack([np.full(nent[d], v), np.full(nent[d], c), np.full(nent[d], d), g, np.full(nent[d], k)], 1))
    recs = np.concatenate(recs, 0)
    ndofs = recs.shape[0]
    if enum_type in _ARRANGEMENT:
        col = {"VAR": 0, "DIM": 1, "ELEM_TYPE": 2, "ELEM_ID": 3, "DOF_ID": 4}
        keys = [recs[:, col[name]] for name in _ARRANGEMENT[enum_type]]
        order = np.lexsort(tuple(reversed(keys)))          # lexsort: last key is the primary one
        ids = np.empty(ndofs, dtype=np.int64)
        ids[order] = np.arange(ndofs)
    elif enum_type in ("ANITYPE", "MINIBLOCKS"):
        i_nd = [sum(NDOF[fem][d] * vec for fem, vec in variables) for d in range(4)]    # dofs per entity of dimension d
        init = np.concatenate([[0], np.cumsum([i_nd[d] * nent[d] for d in range(4)])])
        # iodf = GetElemDofId (global_enumerator.h:76, .cpp:262-274): variables in order; inside a vector variable component c,
        # dof k of the component -> c * nd + k (dim_dof_shift = nd * dim_id)
        shift = {}
        for d in range(4):
            o = 0
            for v, (fem, vec) in enumerate(variables):
                shift[(v, d)] = o
                o += NDOF[fem][d] * vec
        vecs = np.array([vec for _, vec in variables])
        ns_rec = np.array([NDOF[variables[int(r[0])][0]][int(r[2])] for r in recs[:, :3]])
        iodf = np.array([shift[(int(r[0]), int(r[2]))] for r in recs[:, :3]]) + recs[:, 1] * ns_rec + recs[:, 4]
        d = recs[:, 2]
        if enum_type == "ANITYPE":
            ids = init[d] + recs[:, 3] + iodf * np.array(nent)[d]
        else:
            ids = init[d] + recs[:, 3] * np.array(i_nd)[d] + iodf
    else:
        raise ValueError("unknown enumeration type " + str(enum_type))
    assert np.array_equal(np.sort(ids), np.arange(ndofs)), "the enumeration is not a bijection"
    lookup = {}
    for r, i in zip(recs, ids):
        lookup[(int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]))] = int(i)
    cols = []
    for v, (fem, vec) in enumerate(variables):
        for c in range(vec):
            for d in range(4):
                nd = NDOF[fem][d]
                if nd == 0:
                    continue
                ents = ent_of_tet[d]
                for le in range(ents.shape[1]):
                    ks = [np.full(ntet, k) for k in range(nd)]
                    if d == 1 and nd == 2:   # S2 pair on an edge, oriented by the node ids (tetdofmap.inl:98-104)
                        a, b = LOCAL_EDGES[le]
                        flip = (tets[:, a] > tets[:, b]).astype(np.int64)
                        ks = [flip, 1 - flip]
                    for kk in ks:
                        cols.append(np.array([lookup[(v, c, d, int(g), int(k))] for g, k in zip(ents[:, le], kk)], dtype=np.int64))
    return np.stack(cols, 1), ndofs


def template_pattern(rowcode, colcode, row_begin, row_end):
    """AssembleTemplate: sorted CSR rows over [row_begin,row_end) incl. forced diagonal"""
    ne, nrow = rowcode.shape
    ncol = colcode.shape[1]
    r = np.repeat(np.abs(rowcode) - 1, ncol, axis=1).ravel()
    c = np.tile(np.abs(colcode) - 1, (1, nrow)).ravel()
    act = np.repeat(rowcode != 0, ncol, axis=1).ravel()
    r, c = r[act], c[act]
    diag = np.arange(row_begin, row_end)
    r = np.concatenate([r, diag]); c = np.concatenate([c, diag])
    key = np.unique(r * (int(c.max()) + 1 if c.size else 1) + c)
    m = int(c.max()) + 1 if c.size else 1
    rr, cc = key // m, key % m
    rowptr = np.zeros(row_end - row_begin + 1, dtype=np.int64)
    np.add.at(rowptr, rr - row_begin + 1, 1)
    return np.cumsum(rowptr), cc.astype(np.int32)


class Problem:
    """FemExprDescr-like description: variables + volume forms.
    mat_forms: dicts(trial=var idx, test=var idx, opA, opB, order, ttype, layout, D, alpha)
    rhs_forms: dicts(test=var idx, opB, order, ttype, layout, D, alpha)  -- IDEN(P0) trick (int_tet_test.cpp:448-501)"""

    def __init__(self, variables, mat_forms=(), rhs_forms=()):
        self.vars = list(variables)
        self.mat_forms, self.rhs_forms = list(mat_forms), list(rhs_forms)
        self.var_off, off = [], 0
        for fem, vec in self.vars:
            self.var_off.append(off)
            off += O.op_dims(O.IDEN, fem, vec)[0]
        self.nloc = off

    @staticmethod
    def _D(fm, idx):
        D = fm.get("D")
        if D is None or fm["layout"] == O.L_CONST or idx is None:
            return D
        return np.ascontiguousarray(np.asarray(D)[idx])  # per-tet / per-point data follow the element index

    def element_matrices(self, XY, idx=None, impl="oracle", nthreads=1):
        """A (f, nloc_col, nloc_row) [col-major per element like the reference's m_A], F (f, nloc).
        idx: global element indices of the columns of XY (selects per-tet / per-point tensor data)."""
        f = XY.shape[1]
        A = np.zeros((f, self.nloc, self.nloc))
        F = np.zeros((f, self.nloc))
        for fm in self.mat_forms:
            fa, va = self.vars[fm["trial"]]
            fb, vb = self.vars[fm["test"]]
            form = (fm["opA"], fa, va, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"])
            Ae = O.fem3dtet(form, XY, self._D(fm, idx), impl=impl, mode=1 if impl == "ref" else 0, nthreads=nthreads)
            ca, rb = self.var_off[fm["trial"]], self.var_off[fm["test"]]
            A[:, ca:ca + Ae.shape[1], rb:rb + Ae.shape[2]] += fm.get("alpha", 1.0) * Ae
        for fm in self.rhs_forms:
            fb, vb = self.vars[fm["test"]]
            form = (O.IDEN, O.P0, 1, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"])
            Fe = O.fem3dtet(form, XY, self._D(fm, idx), impl=impl, mode=1 if impl == "ref" else 0, nthreads=nthreads)
            rb = self.var_off[fm["test"]]
            F[:, rb:rb + Fe.shape[2]] += fm.get("alpha", 1.0) * Fe[:, 0, :]
        return A, F


    def face_matrices(self, XY, face, idx=None, impl="oracle"):
        """The same for surface integrals (fem3Dface over face face[r] of the tet XY[:, r]; forms and data of THIS Problem are
        the face forms, per-face data indexed by idx = boundary-face indices): A (f, nloc, nloc), F (f, nloc)."""
        f = XY.shape[1]
        A = np.zeros((f, self.nloc, self.nloc))
        F = np.zeros((f, self.nloc))
        for fm in self.mat_forms:
            fa, va = self.vars[fm["trial"]]
            fb, vb = self.vars[fm["test"]]
            form = (fm["opA"], fa, va, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"])
            Ae = O.fem3dface(form, XY, face, self._D(fm, idx), impl=impl)
            ca, rb = self.var_off[fm["trial"]], self.var_off[fm["test"]]
            A[:, ca:ca + Ae.shape[1], rb:rb + Ae.shape[2]] += fm.get("alpha", 1.0) * Ae
        for fm in self.rhs_forms:
            fb, vb = self.vars[fm["test"]]
            form = (O.IDEN, O.P0, 1, fm["opB"], fb, vb, fm["order"], fm["ttype"], fm["layout"])
            Fe = O.fem3dface(form, XY, face, self._D(fm, idx), impl=impl)
            rb = self.var_off[fm["test"]]
            F[:, rb:rb + Fe.shape[2]] += fm.get("alpha", 1.0) * Fe[:, 0, :]
        return A, F


def apply_dir(A, F, colcode, flag, value):
    """applyDir(A, F, k, bc) of fem/operations/dc_on_dof.h:27-45 for every Dirichlet local dof k of every element, as the
    reference's local assemblers do (examples/tutorials/ex1.cpp:96-105): F(i) -= A(i,k) bc; F(k) = bc; row and column k
    of A zeroed; A(k,k) = 1.  A is (f, ncol, nrow) with A[e, k, i] = A_e(i, k).  The result does not depend on the order of k."""
    g = np.abs(colcode) - 1
    d = flag[g].astype(bool)
    bcv = np.where(d, value[g], 0.0)
    F -= np.einsum("eki,ek->ei", A, bcv)
    F[d] = bcv[d]
    A[d, :] = 0.0                                    # column k
    A[np.broadcast_to(d[:, None, :], A.shape)] = 0.0  # row k
    e, k = np.nonzero(d)
    A[e, k, k] = 1.0


def assemble(problem, coords, tets, dofmap, rank=None, impl="oracle", drop_val=1e-100, chunk=200000, dirichlet=None, faces=None):
    """Full reference-style assembly on `rank`'s row interval: returns rowptr, colind, val, rhs.
    dirichlet = (flag[ndof], value[ndof]): essential BCs applied to every element matrix like the reference's examples.
    faces = (face_tet[nbf], face_num[nbf], Problem of the surface forms): Neumann / Robin terms added to the cell matrix by the
    local assembler before the essential BCs (examples/Fem/Ani/diffusion.cpp:215-245)."""
    r = 0 if rank is None else rank
    row_begin, row_end = (0, dofmap.nrows) if rank is None else (int(dofmap.beg_ind[r]), int(dofmap.end_ind[r]))
    rowcode, colcode = dofmap.codes(rank)
    sel = np.nonzero((rowcode != 0).any(axis=1))[0]  # has_active (assembler.inl:367)
    rowcode, colcode = rowcode[sel], colcode[sel]
    rowptr, colind = template_pattern(rowcode, colcode, row_begin, row_end)
    val = np.zeros(colind.shape[0])
    rhs = np.zeros(row_end - row_begin)
    status = 0
    for s in range(0, sel.shape[0], chunk):
        idx = sel[s:s + chunk]
        XY = coords[tets[idx]].transpose(1, 0, 2)  # (4, f, 3)
        A, F = problem.element_matrices(XY, idx=idx, impl=impl)
        if faces is not None:
            ft, fn, fprob = faces
            ft = np.asarray(ft)
            loc = np.full(tets.shape[0], -1, dtype=np.int64)
            loc[idx] = np.arange(idx.shape[0])
            fsel = np.nonzero(loc[ft] >= 0)[0]           # boundary faces of the cells of this chunk, ascending face index
            if fsel.size:
                Af, Ff = fprob.face_matrices(coords[tets[ft[fsel]]].transpose(1, 0, 2), np.asarray(fn)[fsel], idx=fsel, impl=impl)
                np.add.at(A, loc[ft[fsel]], Af)
                np.add.at(F, loc[ft[fsel]], Ff)
        if dirichlet is not None:
            apply_dir(A, F, colcode[s:s + chunk], np.asarray(dirichlet[0]), np.asarray(dirichlet[1], dtype=float))
        st = O.scatter_csr(rowcode[s:s + chunk], colcode[s:s + chunk], A, F, row_begin, rowptr, colind, val, rhs, drop_val)
        status = min(status, st)
    return rowptr, colind, val, rhs, status
