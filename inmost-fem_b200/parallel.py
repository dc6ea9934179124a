"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing): element partition by box
blocks, owned row intervals, interface-row exchange plan.

Reference behaviour being replaced (AniFem++ / INMOST):
  * box-block partition of the cube over ranks              utils/mesh_utils.cpp:67-108
  * per-rank contiguous row interval [BegInd, EndInd)       inmost_interface/global_enumerator.cpp:562-605
  * index tags of shared entities exchanged at setup        global_enumerator.cpp:698,754 (mesh->ExchangeData)
  * assembly itself: the reference recomputes ghost cells and exchanges no values (assembler.inl:162-183);
    BASELINE.json's north_star asks for owner-computes instead: every GPU assembles its own elements, and the
    contributions to interface rows owned by another rank are exchanged (NCCL all-to-all) and added in rank order.

Conventions standing in for INMOST (same as oracle/asm_oracle.py, documented in DESIGN.md): an entity shared by
several blocks is owned by the lowest rank; inside a rank, owned nodes are ordered by global node id and owned
edges lexicographically by their (min,max) global node ids; NATURAL numbering inside the rank's interval.

Everything here is device-agnostic torch code (runs on CPU/gloo in the tests, on CUDA/NCCL in bench.py); the
numerics stay in the CUDA library.  Scope: variables living on nodes, edges and faces (P1, P2, P3, vectors of them,
Taylor-Hood); P0 (cell dofs) is not partitioned.  P3 is covered by the CPU (gloo) tests of the numbering / exchange plan only.
"""
import os

import torch
import torch.distributed as dist

NDOF = {2: (1, 0, 0), 3: (1, 1, 0), 4: (1, 2, 1)}  # fem (P1, P2, P3) -> dofs per (node, edge, face); P0 (cell dofs) is not partitioned
LOCAL_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
LOCAL_FACES = ((0, 1, 2), (1, 2, 3), (2, 3, 0), (3, 0, 1))


def proc_grid(nranks, sizes):
    """process grid of GenerateParallelepiped (utils/mesh_utils.cpp:67-86)"""
    divs, d = [], nranks
    while d > 1:
        for k in range(2, d + 1):
            if d % k == 0:
                divs.append(k)
                d //= k
                break
    ppa, epp = [1, 1, 1], list(sizes)
    for k in reversed(divs):
        m = max(range(3), key=lambda a: (epp[a], -a))  # first maximum, like std::max_element
        ppa[m] *= k
        epp[m] //= k
    return ppa


def block_of_rank(rank, nranks, sizes):
    """(bx,by,bz,lx,ly,lz) of `rank` (utils/mesh_utils.cpp:88-108)"""
    ppa = proc_grid(nranks, sizes)
    pc = (rank % ppa[0], rank // ppa[0] % ppa[1], rank // (ppa[0] * ppa[1]))
    out_b, out_l = [], []
    for a in range(3):
        avg = -(-sizes[a] // ppa[a])
        start = avg * pc[a]
        size = sizes[a] - avg * (ppa[a] - 1) if pc[a] == ppa[a] - 1 else avg
        out_b.append(start)
        out_l.append(size)
    return tuple(out_b) + tuple(out_l)


def _all_gather_var(t, group=None):
    """all_gather of 1-D tensors of different lengths -> list of tensors"""
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=t.dtype, device=t.device)
    buf[:t.numel()] = t
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def _all_to_all_var(parts, group=None):
    """parts[p] goes to rank p; returns the list received from every rank (variable sizes).  Uses
    all_to_all_single (NCCL); falls back to an all_gather emulation where the backend lacks it (gloo)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev, dt = parts[0].device, parts[0].dtype
    send_sizes = torch.tensor([p.numel() for p in parts], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(send_sizes) for _ in range(world)]
    dist.all_gather(all_sizes, send_sizes, group=group)
    recv_sizes = [int(all_sizes[p][rank].item()) for p in range(world)]
    send = torch.cat(parts) if sum(p.numel() for p in parts) else torch.zeros(0, dtype=dt, device=dev)
    try:
        recv = torch.zeros(sum(recv_sizes), dtype=dt, device=dev)
        dist.all_to_all_single(recv, send, recv_sizes, [p.numel() for p in parts], group=group)
        return list(torch.split(recv, recv_sizes))
    except (RuntimeError, NotImplementedError):
        everything = _all_gather_var(send, group)
        out = []
        for p in range(world):
            off = int(all_sizes[p][:rank].sum().item())
            out.append(everything[p][off:off + recv_sizes[p]].clone())
        return out


class Numbering:
    """Global NATURAL numbering of the local block + owned interval, built with one metadata exchange."""

    ENUM_TYPES = ("ANITYPE", "MINIBLOCKS", "NATURAL", "DIMUNION", "BYELEMTYPE", "ETDIMBLOCKS")   # global_enumerator.h:393-401

    def __init__(self, tets, gnode, iface_node, variables, nn_global, group=None, enum_type="NATURAL"):
        """tets (ntet,4) local node ids; gnode (nnode,) global node id of each local node; iface_node (nnode,) bool:
        node may be shared with another rank; variables [(fem, vec)]; nn_global = total number of nodes; enum_type: the
        GlobEnumeration type ordering the dofs inside every rank's interval (closed forms of include/anifem_b200/enumerator.hpp
        with the rank's own entity counts and ELEM_ID = position among the rank's owned entities)."""
        if enum_type not in self.ENUM_TYPES:
            raise ValueError("Faced unknown ASSEMBLING_TYPE")
        self.enum_type = enum_type
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = tets.device
        self.vars = list(variables)
        for fem, _ in self.vars:
            if fem not in NDOF:
                raise NotImplementedError("multi-GPU numbering covers P1 / P2 / P3 based variables")
        tets = tets.long()
        need_e = any(NDOF[f][1] for f, _ in self.vars)
        need_f = any(NDOF[f][2] for f, _ in self.vars)
        # ---- local entities with canonical keys
        ent_key = [gnode.long()]
        ent_of_tet = [tets]
        ent_iface = [iface_node.bool()]
        if need_e:
            ga = torch.stack([gnode[tets[:, a]] for a, b in LOCAL_EDGES], 1).long()
            gb = torch.stack([gnode[tets[:, b]] for a, b in LOCAL_EDGES], 1).long()
            key = torch.minimum(ga, gb) * nn_global + torch.maximum(ga, gb)
            ekeys, inv = torch.unique(key.reshape(-1), sorted=True, return_inverse=True)
            both = torch.stack([iface_node[tets[:, a]] & iface_node[tets[:, b]] for a, b in LOCAL_EDGES], 1).reshape(-1)
            eif = torch.zeros(ekeys.numel(), dtype=torch.bool, device=dev)
            eif[inv[both]] = True
            ent_key.append(ekeys)
            ent_of_tet.append(inv.reshape(-1, 6))
            ent_iface.append(eif)
        if need_f:
            # faces: canonical key = sorted global node triple as one integer (lexicographic order of the triples, like the
            # edges); the three factors must fit 63 bits
            if nn_global >= (1 << 21):
                raise NotImplementedError("P3 on a partitioned mesh: face keys need more than 63 bits above 2^21 global nodes")
            tri = torch.stack([torch.stack([gnode[tets[:, a]] for a in f], 1).long() for f in LOCAL_FACES], 1)   # (ntet, 4, 3)
            tri = torch.sort(tri, dim=2).values
            key = (tri[..., 0] * nn_global + tri[..., 1]) * nn_global + tri[..., 2]
            fkeys, inv = torch.unique(key.reshape(-1), sorted=True, return_inverse=True)
            allif = torch.stack([iface_node[tets[:, f[0]]] & iface_node[tets[:, f[1]]] & iface_node[tets[:, f[2]]] for f in LOCAL_FACES], 1).reshape(-1)
            fif = torch.zeros(fkeys.numel(), dtype=torch.bool, device=dev)
            fif[inv[allif]] = True
            if not need_e:   # keep the (node, edge, face) slots aligned
                ent_key.append(torch.zeros(0, dtype=torch.int64, device=dev))
                ent_of_tet.append(torch.zeros((tets.shape[0], 0), dtype=torch.int64, device=dev))
                ent_iface.append(torch.zeros(0, dtype=torch.bool, device=dev))
            ent_key.append(fkeys)
            ent_of_tet.append(inv.reshape(-1, 4))
            ent_iface.append(fif)
        nd_types = len(ent_key)
        # ---- ownership: lowest rank among the ranks that hold the entity
        self.owner, self.pos = [], []
        num = torch.zeros((nd_types, self.world), dtype=torch.int64)
        gathered = []
        for d in range(nd_types):
            cand = torch.sort(ent_key[d][ent_iface[d]]).values
            allc = _all_gather_var(cand, group)
            owner = torch.full((ent_key[d].numel(),), self.rank, dtype=torch.int64, device=dev)
            for p in range(self.rank - 1, -1, -1):  # lower ranks win; scan downwards so the minimum sticks
                if allc[p].numel() == 0:
                    continue
                ix = torch.searchsorted(allc[p], ent_key[d]).clamp(max=allc[p].numel() - 1)
                has = (allc[p][ix] == ent_key[d]) & ent_iface[d]
                owner[has] = p
            self.owner.append(owner)
            owned = owner == self.rank
            # position among the owned entities in canonical (key) order
            order = torch.argsort(ent_key[d][owned])
            pos = torch.full((ent_key[d].numel(),), -1, dtype=torch.int64, device=dev)
            idx_owned = torch.nonzero(owned).reshape(-1)
            pos[idx_owned[order]] = torch.arange(idx_owned.numel(), device=dev)
            self.pos.append(pos)
            cnt = torch.tensor([int(owned.sum().item())], dtype=torch.int64, device=dev)
            cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
            dist.all_gather(cnts, cnt, group=group)
            num[d] = torch.tensor([int(c.item()) for c in cnts])
            # table of my owned interface entities (key -> pos) for the ghosts of other ranks
            mine = owned & ent_iface[d]
            k_m, o_m = torch.sort(ent_key[d][mine])
            gathered.append((_all_gather_var(k_m, group), _all_gather_var(pos[mine][o_m], group)))
        for d in range(nd_types):  # positions of ghost entities inside their owner
            ghost = self.owner[d] != self.rank
            for p in range(self.world):
                sel = ghost & (self.owner[d] == p)
                if not bool(sel.any()):
                    continue
                keys_p, pos_p = gathered[d]
                ix = torch.searchsorted(keys_p[p], ent_key[d][sel])
                assert bool((keys_p[p][ix] == ent_key[d][sel]).all()), "ghost entity missing at its owner"
                self.pos[d][sel] = pos_p[p][ix]
        # ---- NATURAL intervals (global_enumerator.cpp:594-604) and group offsets per rank
        beg = torch.cumsum(num, 1) - num                     # BegElemID[d][r]
        ndof_ent = [sum(NDOF[f][d] * v for f, v in self.vars) for d in range(nd_types)]
        self.beg_ind = sum(beg[d] * ndof_ent[d] for d in range(nd_types))
        self.end_ind = self.beg_ind + sum(num[d] * ndof_ent[d] for d in range(nd_types))
        self.nrows_global = int(sum(int(num[d].sum()) * ndof_ent[d] for d in range(nd_types)))
        grp_off, off = {}, torch.zeros(self.world, dtype=torch.int64)
        et_first = enum_type in ("BYELEMTYPE", "ETDIMBLOCKS")            # ELEM_TYPE is the leading key of the arrangement
        dim_last = enum_type in ("DIMUNION", "ETDIMBLOCKS")              # the component is the innermost key
        groups = [(v, c, d) for v, (fem, vec) in enumerate(self.vars) for c in range(vec) for d in range(nd_types)]
        if et_first:
            groups.sort(key=lambda t: (t[2], t[0], t[1]))
        if dim_last:   # one group per (variable, entity type): all components interleaved
            groups = [t for t in groups if t[1] == 0]
        for v, c, d in groups:
            fem, vec = self.vars[v]
            if NDOF[fem][d]:
                size = num[d] * NDOF[fem][d] * (vec if dim_last else 1)
                for cc in (range(vec) if dim_last else (c,)):
                    grp_off[(v, cc, d)] = off.clone()
                off = off + size
        # SimpleEnumerator (global_enumerator.cpp:671-700,818-835): InitElemIndex per entity type, position of a dof among all
        # dofs of its entity (GetElemDofId: variables in order, component-major inside a vector variable)
        init = torch.zeros((nd_types + 1, self.world), dtype=torch.int64)
        for d in range(nd_types):
            init[d + 1] = init[d] + num[d] * ndof_ent[d]
        shift = {}
        for d in range(nd_types):
            o = 0
            for v, (fem, vec) in enumerate(self.vars):
                shift[(v, d)] = o
                o += NDOF[fem][d] * vec

        def index(v, c, d, ow, pos, k):
            """global id of dof k (int or tensor) of component c of variable v on the entities (owner ow, position pos)"""
            fem, vec = self.vars[v]
            ns = NDOF[fem][d]
            if enum_type in ("NATURAL", "BYELEMTYPE"):
                loc = grp_off[(v, c, d)].to(dev)[ow] + pos * ns + k
            elif dim_last:
                loc = grp_off[(v, c, d)].to(dev)[ow] + (pos * ns + k) * vec + c
            elif enum_type == "ANITYPE":
                loc = init[d].to(dev)[ow] + pos + (shift[(v, d)] + c * ns + k) * num[d].to(dev)[ow]
            else:   # MINIBLOCKS: the layout the reference's inverse map decodes (:866-871)
                loc = init[d].to(dev)[ow] + pos * ndof_ent[d] + shift[(v, d)] + c * ns + k
            return beg_ind_dev[ow] + loc
        # ---- element -> global dof (local order: variable, component, 4 vertices, 6 edges)
        cols = []
        beg_ind_dev = self.beg_ind.to(dev)
        node_gid = beg[0].to(dev)[self.owner[0]] + self.pos[0]   # GlobalID of every local node (BegElemID of its owner + position)
        for v, (fem, vec) in enumerate(self.vars):
            for c in range(vec):
                for d in range(nd_types):
                    nd = NDOF[fem][d]
                    if not nd:
                        continue
                    ents = ent_of_tet[d]
                    ow = self.owner[d][ents]
                    ps = self.pos[d][ents]
                    for le in range(ents.shape[1]):
                        if d == 1 and nd == 2:
                            # the two dofs of a P3 edge follow the orientation of the edge by the GLOBAL IDS of its end points
                            # (tetdofmap.inl:98-104).  Under a partition these are the rank-major ids INMOST assigns (owner's first
                            # node id + position among its owned nodes), not the mesh node index: the two orders differ as soon
                            # as the partition cuts the fastest mesh axis (found by the 8-rank parity leg, 2 x 2 x 2 blocks)
                            a, b = LOCAL_EDGES[le]
                            flip = (node_gid[tets[:, a]] > node_gid[tets[:, b]]).long()
                            cols.append(index(v, c, d, ow[:, le], ps[:, le], flip))
                            cols.append(index(v, c, d, ow[:, le], ps[:, le], 1 - flip))
                            continue
                        for k in range(nd):
                            cols.append(index(v, c, d, ow[:, le], ps[:, le], k))
        self.elem2dof = torch.stack(cols, 1).contiguous()
        self.nloc = self.elem2dof.shape[1]
        self.row_begin, self.row_end = int(self.beg_ind[self.rank]), int(self.end_ind[self.rank])
        self.num, self.grp_off, self.nd_types = num, grp_off, nd_types

    def fields(self, plan):
        """Scalar fields (variable, component) for afb_fields_set: (fem, first local dof, row intervals, column intervals).
        A field is contiguous inside every rank's interval (NATURAL: VAR, component, DIM, ... global_enumerator.cpp:702-777):
        its global columns are one interval per rank; its local rows are the own interval followed by the runs of its
        dofs inside the sorted list of foreign rows (one run per peer)."""
        if self.enum_type != "NATURAL":
            return []   # other arrangements interleave components or split a field by entity type: no field intervals, generic path
        out, loff = [], 0
        foreign = plan.foreign.cpu()
        for v, (fem, vec) in enumerate(self.vars):
            ds = [d for d in range(self.nd_types) if NDOF[fem][d]]
            nl = 4 * NDOF[fem][0] + 6 * NDOF[fem][1] + 4 * NDOF[fem][2]
            for c in range(vec):
                cols = []
                for p in range(self.world):
                    start = int(self.beg_ind[p] + self.grp_off[(v, c, ds[0])][p])
                    count = int(sum(int(self.num[d][p]) * NDOF[fem][d] for d in ds))
                    cols.append((start, count))
                rows = [(cols[self.rank][0] - self.row_begin, cols[self.rank][1])]
                for p in range(self.world):
                    if p == self.rank or foreign.numel() == 0:
                        continue
                    lo = int(torch.searchsorted(foreign, torch.tensor([cols[p][0]])).item())
                    hi = int(torch.searchsorted(foreign, torch.tensor([cols[p][0] + cols[p][1]])).item())
                    rows.append((plan.n_own + lo, hi - lo))
                out.append((fem, loff, rows, cols))
                loff += nl
        return out


class InterfacePlan:
    """Extended local row space (owned rows, then the interface rows of other ranks) + exchange plan."""

    def __init__(self, numbering):
        nb = self.nb = numbering
        dev = nb.elem2dof.device
        g = nb.elem2dof
        own = (g >= nb.row_begin) & (g < nb.row_end)
        self.n_own = nb.row_end - nb.row_begin
        self.foreign = torch.unique(g[~own], sorted=True)          # global ids of rows owned elsewhere, ascending
        self.n_for = int(self.foreign.numel())
        loc = torch.where(own, g - nb.row_begin, self.n_own + torch.searchsorted(self.foreign, g).clamp(max=max(self.n_for - 1, 0)))
        self.rowcode = (loc + 1).contiguous()                      # local rows, 1-based codes
        self.colcode = (g + 1).contiguous()                        # global columns
        self.diag_col = torch.cat([torch.arange(nb.row_begin, nb.row_end, device=dev),
                                   torch.full((self.n_for,), -1, dtype=torch.int64, device=dev)]).contiguous()
        end_ind = nb.end_ind.to(dev)
        self.for_owner = torch.searchsorted(end_ind, self.foreign, right=True)   # owner rank of each foreign row
        self.for_per_peer = [int((self.for_owner == p).sum().item()) for p in range(nb.world)]

    def finalize_pattern(self, rowptr, colind):
        """rowptr/colind: local pattern of the extended rows (from the local elements only).  Exchanges the column
        lists of the foreign rows, returns the final pattern (owned rows = union with what the peers contribute)
        and stores the receive-side slot maps."""
        nb = self.nb
        dev = colind.device
        NC = nb.nrows_global
        rowptr = rowptr.long()
        cols = colind.long()
        counts = rowptr[1:] - rowptr[:-1]
        rows = torch.repeat_interleave(torch.arange(self.n_own + self.n_for, device=dev), counts)
        own_keys = rows[:int(rowptr[self.n_own])] * NC + cols[:int(rowptr[self.n_own])]
        # foreign part: (global row, col) pairs in CSR order = the order in which their values will be sent
        f_rows_g = self.foreign[rows[int(rowptr[self.n_own]):] - self.n_own]
        f_cols = cols[int(rowptr[self.n_own]):]
        f_owner = torch.searchsorted(nb.end_ind.to(dev), f_rows_g, right=True)
        send = [torch.stack([f_rows_g[f_owner == p], f_cols[f_owner == p]], 1).reshape(-1) for p in range(nb.world)]
        self.send_nnz = [int((f_owner == p).sum().item()) for p in range(nb.world)]
        recv = _all_to_all_var(send, nb.group)
        recv_keys = [(r.reshape(-1, 2)[:, 0] - nb.row_begin) * NC + r.reshape(-1, 2)[:, 1] for r in recv]
        final_keys = torch.unique(torch.cat([own_keys] + recv_keys), sorted=True)
        own_rows = final_keys // NC
        rowptr_own = torch.zeros(self.n_own + 1, dtype=torch.int64, device=dev)
        rowptr_own[1:] = torch.cumsum(torch.bincount(own_rows, minlength=self.n_own), 0)
        colind_own = (final_keys % NC).to(torch.int32)
        self.nnz_own = int(final_keys.numel())
        self.recv_nnz = [int(k.numel()) for k in recv_keys]
        self.val_slots = [torch.searchsorted(final_keys, k).contiguous() for k in recv_keys]   # slot in my CSR per received value
        # rhs: one value per foreign row, grouped by owner; receiver adds at the local row
        rrow = _all_to_all_var([self.foreign[self.for_owner == p] for p in range(nb.world)], nb.group)
        self.rhs_slots = [(r - nb.row_begin).contiguous() for r in rrow]
        # final extended pattern = final owned rows + local foreign rows
        f_counts = counts[self.n_own:]
        rowptr_ext = torch.cat([rowptr_own, rowptr_own[-1] + torch.cumsum(f_counts, 0)])
        colind_ext = torch.cat([colind_own, f_cols.to(torch.int32)])
        self.nnz_ext = int(colind_ext.numel())
        return rowptr_ext.contiguous(), colind_ext.contiguous()

    def exchange_start(self, val_ext, rhs_ext):
        """launch the interface exchange (NCCL all_to_all_single, asynchronous): the foreign rows are sorted by global id =
        grouped by owner, so the tail of val_ext / rhs_ext IS the send buffer; all sizes are known from the plan, no size
        exchange, no host synchronisation.  Returns the handles for exchange_finish."""
        nb = self.nb
        works = []
        if val_ext is not None:
            if getattr(self, "_vrecv", None) is None:
                self._vrecv = torch.empty(sum(self.recv_nnz), dtype=val_ext.dtype, device=val_ext.device)
            works.append(dist.all_to_all_single(self._vrecv, val_ext[self.nnz_own:], self.recv_nnz, self.send_nnz, group=nb.group, async_op=True))
        if rhs_ext is not None:
            sizes = [int(r.numel()) for r in self.rhs_slots]
            if getattr(self, "_rrecv", None) is None:
                self._rrecv = torch.empty(sum(sizes), dtype=rhs_ext.dtype, device=rhs_ext.device)
            works.append(dist.all_to_all_single(self._rrecv, rhs_ext[self.n_own:], sizes, self.for_per_peer, group=nb.group, async_op=True))
        return works

    def exchange_finish(self, works, val_ext, rhs_ext, add):
        """wait for the exchange and add what arrived, peers in rank order (add = afb_halo_add: dst[slots] += contrib)"""
        for w in works:
            w.wait()
        if val_ext is not None:
            for p, r in enumerate(torch.split(self._vrecv, self.recv_nnz)):
                if r.numel():
                    add(self.val_slots[p], r, val_ext)
        if rhs_ext is not None:
            for p, r in enumerate(torch.split(self._rrecv, [int(r.numel()) for r in self.rhs_slots])):
                if r.numel():
                    add(self.rhs_slots[p], r, rhs_ext)

    def exchange(self, val_ext, rhs_ext, add):
        """send the foreign-row values / rhs entries to their owners and add what arrives, peers in rank order.
        add(slots, contrib, dst) performs dst[slots] += contrib (afb_halo_add on the GPU)."""
        nb = self.nb
        fixed = val_ext.is_cuda if val_ext is not None else (rhs_ext is not None and rhs_ext.is_cuda)
        if fixed:
            self.exchange_finish(self.exchange_start(val_ext, rhs_ext), val_ext, rhs_ext, add)
            return
        if val_ext is not None:
            for p, r in enumerate(_all_to_all_var(list(torch.split(val_ext[self.nnz_own:], self.send_nnz)), nb.group)):
                if r.numel():
                    add(self.val_slots[p], r.contiguous(), val_ext)
        if rhs_ext is not None:
            for p, r in enumerate(_all_to_all_var(list(torch.split(rhs_ext[self.n_own:], self.for_per_peer)), nb.group)):
                if r.numel():
                    add(self.rhs_slots[p], r.contiguous(), rhs_ext)


class DistributedAssembler:
    """One rank of a multi-GPU assembly on the cube: local block mesh, global numbering, final pattern, exchange.
    Mirrors what a reference user gets from GenerateCube + Assembler::PrepareProblem under mpirun."""

    def __init__(self, ctx, dims, variables, group=None, enum_type="NATURAL"):
        import torch
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nx, ny, nz = dims
        self.block = block_of_rank(self.rank, self.world, dims)
        bx, by, bz, lx, ly, lz = self.block
        ctx.mesh_cube(nx, ny, nz, 1.0, self.block)
        self.coords, self.tets = ctx.mesh_get_torch()
        dev = self.tets.device
        ii = torch.arange(lx + 1, device=dev).view(-1, 1, 1).expand(lx + 1, ly + 1, lz + 1)
        jj = torch.arange(ly + 1, device=dev).view(1, -1, 1).expand(lx + 1, ly + 1, lz + 1)
        kk = torch.arange(lz + 1, device=dev).view(1, 1, -1).expand(lx + 1, ly + 1, lz + 1)
        gnode = (((bx + ii) * (ny + 1) + (by + jj)) * (nz + 1) + (bz + kk)).reshape(-1)
        iface = (((ii == 0) & (bx > 0)) | ((ii == lx) & (bx + lx < nx)) | ((jj == 0) & (by > 0)) | ((jj == ly) & (by + ly < ny)) |
                 ((kk == 0) & (bz > 0)) | ((kk == lz) & (bz + lz < nz))).reshape(-1)
        nn_global = (nx + 1) * (ny + 1) * (nz + 1)
        self.numbering = nb = Numbering(self.tets, gnode, iface, variables, nn_global, group, enum_type)
        self.plan = plan = InterfacePlan(nb)
        n_ext = plan.n_own + plan.n_for
        ctx.dofmap_set_any(plan.rowcode, plan.colcode, 0, n_ext, nb.nrows_global, plan.diag_col)
        self.fields = nb.fields(plan)
        if len(self.fields) > 1 and hasattr(ctx, "fields_set"):
            # vector-valued / mixed spaces: the assembly runs block by block on scalar gather plans (afb_blocks.cu)
            ctx.fields_set(self.fields)
        ctx.pattern_build()
        rp, ci = ctx.pattern_get_torch()
        self.rowptr_ext, self.colind_ext = plan.finalize_pattern(rp, ci)
        ctx.pattern_set(self.rowptr_ext, self.colind_ext)
        # phased assembly: the interface rows of other ranks (local rows >= n_own) are produced first, their exchange
        # overlaps the rest of the assembly (afb_assemble_phase)
        self.phased = bool(self.val_is_cuda()) and self.world > 1
        if self.phased:
            ctx.priority_rows_set(plan.n_own)
        self.val = torch.zeros(plan.nnz_ext, dtype=torch.float64, device=dev)
        self.rhs = torch.zeros(n_ext, dtype=torch.float64, device=dev)
        self.ntet = int(self.tets.shape[0])
        # data path of the exchange inside the library (afb_comm.cu: NCCL send / recv + additions issued from C); torch only
        # carries the 128-byte NCCL id to the ranks.  AFB_PY_EXCHANGE=1 keeps the torch.distributed exchange (A/B); tensors on the
        # CPU (gloo tests of the host logic) always use it.
        self.c_exchange = bool(self.val_is_cuda()) and self.world > 1 and not os.environ.get("AFB_PY_EXCHANGE") and hasattr(ctx, "comm_init")
        if self.c_exchange:
            ids = [ctx.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ctx.comm_init(ids[0], self.rank, self.world)
            cat = lambda ts: torch.cat([t.reshape(-1) for t in ts]).contiguous() if len(ts) else None
            ctx.halo_plan_set(plan.n_own, plan.nnz_own, plan.send_nnz, plan.for_per_peer, plan.recv_nnz,
                              [int(r.numel()) for r in plan.rhs_slots], cat(plan.val_slots), cat(plan.rhs_slots))

    def val_is_cuda(self):
        return self.tets.is_cuda

    @property
    def rowptr(self):
        return self.rowptr_ext[:self.plan.n_own + 1]

    @property
    def colind(self):
        return self.colind_ext[:self.plan.nnz_own]

    def _ctx_stream(self):
        """torch stream context that makes the context's CUDA stream torch's current stream: the NCCL exchange and the halo
        additions are ordered against the assembly kernels only through torch's current stream, so the whole sequence must be
        issued on the stream the library launches on (a Context created without a stream owns a private one)."""
        import contextlib
        if not self.val_is_cuda():
            return contextlib.nullcontext()
        h = self.ctx.stream_handle()
        cur = torch.cuda.current_stream()
        if h == cur.cuda_stream:
            return contextlib.nullcontext()
        ext = torch.cuda.ExternalStream(h)
        ext.wait_stream(cur)          # inputs produced on the caller's stream are visible to the assembly
        self._ext_stream = ext
        return torch.cuda.stream(ext)

    def assemble(self, forms, rhs_forms, drop_val=1e-100):
        """Assemble the owned rows [row_begin,row_end): local element contributions + interface contributions of peers.
        Results: self.val[:nnz_own] (CSR values of the owned rows), self.rhs[:n_own]."""
        self._ext_stream = None
        with self._ctx_stream():
            st = self._assemble(forms, rhs_forms, drop_val)
        if self._ext_stream is not None:   # results are consumed on the caller's stream
            torch.cuda.current_stream().wait_stream(self._ext_stream)
        return st

    def _assemble(self, forms, rhs_forms, drop_val):
        if self.c_exchange:
            return self.ctx.assemble_distributed(forms, rhs_forms, self.val, self.rhs, drop_val=drop_val)
        if self.phased:
            self.ctx.assemble_phase(forms, rhs_forms, self.val, self.rhs, 1, drop_val=drop_val)
            works = self.plan.exchange_start(self.val, self.rhs)
            st = self.ctx.assemble_phase(forms, rhs_forms, self.val, self.rhs, 2, drop_val=drop_val)
            self.plan.exchange_finish(works, self.val, self.rhs, self.ctx.halo_add)
            return st
        st = self.ctx.assemble(forms, rhs_forms, self.val, self.rhs, accumulate=False, drop_val=drop_val)
        self.plan.exchange(self.val, self.rhs, self.ctx.halo_add)
        return st
