// anifem_b200/dofmap.hpp -- local degree-of-freedom maps of a tetrahedron: host twin of the reference's Ani::DofT
// (fem/tetdofmap.h:20-767, tetdofmap.cpp / tetdofmap.inl).
//
// A dof map says where the local dofs of a finite-element space sit on the tetrahedron: dof `gid` (contiguous index on the tet)
// <-> (type of the geometric entity, number of the entity on the tet, number of the dof on that entity).  The reference's
// assembler walks these maps in fill_assemble_templates (inmost_interface/assembler.inl:139-184); the device numbering kernels of
// this library hard-wire the same order for P0..P3 (afb_ctx.cu) -- this header is the API a reference user builds spaces with:
//     UniteDofMap (dofs per entity type), VectorDofMap (k copies), ComplexDofMap (concatenation), the DofMap handle with
//     operator* / operator^ / pow / merge / merge_with_simplifications, TetGeomSparsity (a selection of entities of the tet) and
//     the iteration over the dofs of a selection.
// Orders reproduced (checked against the reference build by tests/test_cxx_shim.py and on the reference's own test scenario,
// tests/fem/tetdofmap_test.cpp): entity types in the order NODE, EDGE_UNORIENT, EDGE_ORIENT, FACE_UNORIENT, FACE_ORIENT, CELL;
// inside a type entity by entity; components of vector / complex maps one after the other, their `leid` shifted by the dofs the
// preceding components have on the same entity.  TetDofID is the exact inverse of LocalOrderOnTet for every map (the reference's
// VectorDofMap::TetDofIDExt divides by the per-tet count, tetdofmap.cpp:449-451, and is not the inverse there).
// Not provided: DofSymmetries (S4 action on the dofs of an entity; the numbering kernels apply the only case the supported
// spaces need, the P3 edge pair, directly).
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

namespace Ani {
namespace DofT {

using uchar = unsigned char;
using uint = unsigned;

static constexpr uchar UNDEF = 0x0, NODE = 0x1, EDGE_UNORIENT = 0x2, EDGE_ORIENT = 0x4, EDGE = 0x2 | 0x4, FACE_UNORIENT = 0x8, FACE_ORIENT = 0x10,
                       FACE = 0x8 | 0x10, CELL = 0x20;
static constexpr uchar NGEOM_TYPES = 6;

/// sequential number of a primitive type (NODE -> 0 ... CELL -> 5), -1 for UNDEF / non-primitive
inline int GeomTypeToNum(uchar etype) {
    for (int k = 0; k < NGEOM_TYPES; ++k) if (etype == (1u << k)) return k;
    return -1;
}
inline uchar NumToGeomType(int num) { return (num >= 0 && num < NGEOM_TYPES) ? static_cast<uchar>(1u << num) : UNDEF; }
/// dimension of the entity a primitive type lives on
inline int NumToGeomDim(int num) { return num < 0 ? -1 : (num == 0 ? 0 : (num <= 2 ? 1 : (num <= 4 ? 2 : 3))); }
inline int GeomTypeDim(uchar etype) {
    int dim = -1;
    for (int k = 0; k < NGEOM_TYPES; ++k) if (etype & (1u << k)) dim = std::max(dim, NumToGeomDim(k));
    return dim;
}
inline uchar DimToGeomType(int dim) { return dim == 0 ? NODE : dim == 1 ? EDGE : dim == 2 ? FACE : dim == 3 ? CELL : UNDEF; }
inline bool GeomTypeIsValid(uchar etype) { return etype < (1u << NGEOM_TYPES); }
/// entities of that dimension on a tetrahedron: 4 nodes, 6 edges, 4 faces, 1 cell
inline uchar DimTetElems(int dim) { return dim == 0 ? 4 : dim == 1 ? 6 : dim == 2 ? 4 : dim == 3 ? 1 : 0; }
inline uchar GeomTypeTetElems(uchar etype) { return DimTetElems(GeomTypeDim(etype)); }

struct TetOrder {
    uint gid = uint(-1);
    TetOrder() = default;
    TetOrder(uint id) : gid(id) {}
    operator uint() const { return gid; }
    bool isValid() const { return gid != uint(-1); }
};
struct LocGeomOrder {
    uint leid = uint(-1);   ///< number of the dof on its entity
    uchar etype = UNDEF;    ///< primitive type of the entity
    uchar nelem = 0;        ///< number of the entity on the tet
    LocGeomOrder() = default;
    LocGeomOrder(uchar etype, uchar nelem, uint leid) : leid(leid), etype(etype), nelem(nelem) {}
    bool operator==(const LocGeomOrder& o) const { return etype == o.etype && nelem == o.nelem && leid == o.leid; }
    bool operator!=(const LocGeomOrder& o) const { return !(*this == o); }
};
struct LocalOrder {
    uint gid = uint(-1), leid = uint(-1);
    uchar etype = UNDEF, nelem = 0;
    uchar stype = uchar(-1), lsid = uchar(-1);   ///< symmetry group data of the reference; not filled here
    LocalOrder() = default;
    LocalOrder(uint gid, uchar etype, uchar nelem, uint leid) : gid(gid), leid(leid), etype(etype), nelem(nelem) {}
    LocalOrder(TetOrder t, LocGeomOrder g) : gid(t.gid), leid(g.leid), etype(g.etype), nelem(g.nelem) {}
    LocGeomOrder getGeomOrder() const { return LocGeomOrder(etype, nelem, leid); }
    TetOrder getTetOrder() const { return TetOrder(gid); }
    bool isValid() const { return gid != uint(-1) && etype != UNDEF; }
    bool operator==(const LocalOrder& o) const { return gid == o.gid; }
    bool operator!=(const LocalOrder& o) const { return gid != o.gid; }
};

/// local numbering of the tetrahedron: edge e joins nodes tet_edge_nodes(e); face f = nodes {f, f+1, f+2 mod 4}
inline std::array<uchar, 2> tet_edge_nodes(int e) {
    static const uchar t[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    return {{t[e][0], t[e][1]}};
}

/// a selection of entities of the tetrahedron (tetdofmap.h:144-196): one bit per entity, by dimension
struct TetGeomSparsity {
    struct Pos {
        uchar elem_dim = 6, elem_num = 6;
        Pos() = default;
        Pos(uchar d, uchar n) : elem_dim(d), elem_num(n) {}
        bool isValid() const { return elem_dim != 6; }
        bool operator==(const Pos& o) const { return elem_dim == o.elem_dim && elem_num == o.elem_num; }
        bool operator!=(const Pos& o) const { return !(*this == o); }
    };
    void clear() { bits = {{0, 0, 0, 0}}; }
    bool empty() const { return !(bits[0] | bits[1] | bits[2] | bits[3]); }
    bool empty(uchar dim) const { return bits[dim] == 0; }
    bool has(uchar dim, int i) const { return (bits[dim] >> i) & 1u; }
    /// closure = the lower-dimensional entities on the boundary of the chosen one
    TetGeomSparsity& set(uchar dim, int i, bool with_closure = false) { return apply(dim, i, with_closure, true); }
    TetGeomSparsity& unset(uchar dim, int i, bool with_closure = false) { return apply(dim, i, with_closure, false); }
    TetGeomSparsity& set(uchar dim) { bits[dim] = static_cast<uchar>((1u << DimTetElems(dim)) - 1u); return *this; }
    TetGeomSparsity& unset(uchar dim) { bits[dim] = 0; return *this; }
    TetGeomSparsity& set(const TetGeomSparsity& o) { for (int d = 0; d < 4; ++d) bits[d] |= o.bits[d]; return *this; }
    TetGeomSparsity& unset(const TetGeomSparsity& o) { for (int d = 0; d < 4; ++d) bits[d] &= static_cast<uchar>(~o.bits[d]); return *this; }
    TetGeomSparsity& setCell(bool with_closure = false) { return set(3, 0, with_closure); }
    TetGeomSparsity& unsetCell(bool with_closure = false) { return unset(3, 0, with_closure); }
    TetGeomSparsity& setFace(int f, bool with_closure = false) { return set(2, f, with_closure); }
    TetGeomSparsity& unsetFace(int f, bool with_closure = false) { return unset(2, f, with_closure); }
    TetGeomSparsity& setEdge(int e, bool with_closure = false) { return set(1, e, with_closure); }
    TetGeomSparsity& unsetEdge(int e, bool with_closure = false) { return unset(1, e, with_closure); }
    TetGeomSparsity& setNode(int n) { return set(0, n); }
    TetGeomSparsity& unsetNode(int n) { return unset(0, n); }
    TetGeomSparsity& setFaces() { return set(2); }
    TetGeomSparsity& unsetFaces() { return unset(2); }
    TetGeomSparsity& setEdges() { return set(1); }
    TetGeomSparsity& unsetEdges() { return unset(1); }
    TetGeomSparsity& setNodes() { return set(0); }
    TetGeomSparsity& unsetNodes() { return unset(0); }
    /// ids of the chosen entities of one dimension and their count
    std::pair<std::array<uchar, 6>, uchar> getElemsIds(uchar dim) const {
        std::pair<std::array<uchar, 6>, uchar> r{{{0, 0, 0, 0, 0, 0}}, 0};
        for (uchar i = 0; i < DimTetElems(dim); ++i) if (has(dim, i)) r.first[r.second++] = i;
        return r;
    }
    Pos endPos() const { return Pos(); }
    Pos beginPos() const { return scan(0, 0, 4); }
    Pos beginPos(uchar dim) const { return scan(dim, 0, dim + 1); }
    Pos nextPos(Pos p) const { return p.isValid() ? scan(p.elem_dim, p.elem_num + 1, 4) : endPos(); }
    Pos nextPosOnDim(Pos p) const { return p.isValid() ? scan(p.elem_dim, p.elem_num + 1, p.elem_dim + 1) : endPos(); }
    friend TetGeomSparsity operator&(TetGeomSparsity a, const TetGeomSparsity& b) { for (int d = 0; d < 4; ++d) a.bits[d] &= b.bits[d]; return a; }
    friend TetGeomSparsity operator|(TetGeomSparsity a, const TetGeomSparsity& b) { for (int d = 0; d < 4; ++d) a.bits[d] |= b.bits[d]; return a; }
    friend TetGeomSparsity operator^(TetGeomSparsity a, const TetGeomSparsity& b) { for (int d = 0; d < 4; ++d) a.bits[d] ^= b.bits[d]; return a; }
    friend TetGeomSparsity operator~(TetGeomSparsity a) {
        for (int d = 0; d < 4; ++d) a.bits[d] = static_cast<uchar>(~a.bits[d] & ((1u << DimTetElems(d)) - 1u));
        return a;
    }

private:
    std::array<uchar, 4> bits{{0, 0, 0, 0}};   // nodes, edges, faces, cell
    Pos scan(int dim, int first, int dim_end) const {
        for (int d = dim; d < dim_end; ++d, first = 0)
            for (int i = first; i < DimTetElems(d); ++i) if (has(static_cast<uchar>(d), i)) return Pos(static_cast<uchar>(d), static_cast<uchar>(i));
        return Pos();
    }
    TetGeomSparsity& apply(uchar dim, int i, bool closure, bool on) {
        if (dim > 3) throw std::runtime_error("Wrong dimesion");
        auto put = [&](uchar d, uchar m) { if (on) bits[d] |= m; else bits[d] &= static_cast<uchar>(~m); };
        if (dim == 3 && !on && closure) { clear(); return *this; }
        put(dim, static_cast<uchar>(1u << i));
        if (!closure) return *this;
        uchar nodes = 0;
        if (dim == 1) { auto n = tet_edge_nodes(i); nodes = static_cast<uchar>((1u << n[0]) | (1u << n[1])); }
        if (dim == 2) nodes = static_cast<uchar>((1u << i) | (1u << ((i + 1) & 3)) | (1u << ((i + 2) & 3)));
        if (dim == 3) nodes = 0xF;
        if (dim >= 1) put(0, nodes);
        if (dim >= 2) {   // the edges whose two end points belong to the entity
            uchar edges = 0;
            for (int e = 0; e < 6; ++e) { auto n = tet_edge_nodes(e); if (((nodes >> n[0]) & 1u) && ((nodes >> n[1]) & 1u)) edges |= static_cast<uchar>(1u << e); }
            put(1, edges);
        }
        if (dim == 3) put(2, 0xF);
        return *this;
    }
};

/// interface of a local dof map (tetdofmap.h:272-346)
struct BaseDofMap {
    enum class BaseTypes { Unknown = 0, UniteType = 1, VectorType = 2, ComplexType = 3 };
    virtual ~BaseDofMap() = default;
    virtual uint ActualType() const { return static_cast<uint>(BaseTypes::Unknown); }
    /// dofs on ONE entity of the primitive type
    virtual uint NumDof(uchar etype) const = 0;
    /// dofs on all entities of that type of the tet
    virtual uint NumDofOnTet(uchar etype) const { return NumDof(etype) * GeomTypeTetElems(etype); }
    virtual uint NumDofOnTet() const { uint s = 0; for (int t = 0; t < NGEOM_TYPES; ++t) s += NumDofOnTet(NumToGeomType(t)); return s; }
    std::array<uint, NGEOM_TYPES> NumDofs() const { std::array<uint, NGEOM_TYPES> r; for (int t = 0; t < NGEOM_TYPES; ++t) r[t] = NumDof(NumToGeomType(t)); return r; }
    std::array<uint, NGEOM_TYPES> NumDofsOnTet() const { std::array<uint, NGEOM_TYPES> r; for (int t = 0; t < NGEOM_TYPES; ++t) r[t] = NumDofOnTet(NumToGeomType(t)); return r; }
    virtual LocalOrder LocalOrderOnTet(TetOrder dof_id) const = 0;
    LocalOrder LocalOrderOnTet(LocGeomOrder g) const { return LocalOrder(TetOrder(TetDofID(g)), g); }
    virtual uint TetDofID(LocGeomOrder dof_id) const = 0;
    virtual uint TypeOnTet(uint dof) const { return LocalOrderOnTet(TetOrder(dof)).etype; }
    bool DefinedOn(uchar etype) const { return NumDof(etype) > 0; }
    uint GetGeomMask() const { uint m = 0; for (int t = 0; t < NGEOM_TYPES; ++t) if (DefinedOn(NumToGeomType(t))) m |= 1u << t; return m; }
    bool isValidIndex(TetOrder d) const { return d.gid < NumDofOnTet(); }
    bool isValidIndex(LocGeomOrder g) const { return GeomTypeToNum(g.etype) >= 0 && g.nelem < GeomTypeTetElems(g.etype) && g.leid < NumDof(g.etype); }
    LocalOrder operator[](TetOrder d) const { return LocalOrderOnTet(d); }
    LocalOrder operator[](LocGeomOrder g) const { return LocalOrderOnTet(g); }
    LocalOrder operator[](uint d) const { return LocalOrderOnTet(TetOrder(d)); }
    LocalOrder at(TetOrder d) const { if (!isValidIndex(d)) throw std::out_of_range("Not valid index"); return LocalOrderOnTet(d); }
    LocalOrder at(LocGeomOrder g) const { if (!isValidIndex(g)) throw std::out_of_range("Not valid index"); return LocalOrderOnTet(g); }
    /// number of first-level components (0 for a UniteDofMap)
    virtual uint NestedDim() const = 0;
    /// component reached by the path ext_dims[0..ndims) through nested vector / complex maps (nullptr if the path leaves the map)
    virtual std::shared_ptr<BaseDofMap> GetSubDofMap(const int* ext_dims, int ndims) const = 0;
    /// first tet index and per-type `leid` shift of that component inside this map
    virtual bool GetNestedShift(const int* ext_dims, int ndims, uint& shift_on_tet, std::array<uint, NGEOM_TYPES>& shift_leid) const = 0;
    virtual bool operator==(const BaseDofMap& o) const = 0;
    bool operator!=(const BaseDofMap& o) const { return !(*this == o); }
    virtual std::shared_ptr<BaseDofMap> Copy() const = 0;

    /// all dofs in tet order
    struct DofIterator {
        const BaseDofMap* map = nullptr;
        uint gid = 0;
        LocalOrder operator*() const { return map->LocalOrderOnTet(TetOrder(gid)); }
        DofIterator& operator++() { ++gid; return *this; }
        bool operator==(const DofIterator& o) const { return gid == o.gid; }
        bool operator!=(const DofIterator& o) const { return gid != o.gid; }
    };
    DofIterator begin() const { return DofIterator{this, 0}; }
    DofIterator end() const { return DofIterator{this, NumDofOnTet()}; }
    /// dofs on the entities of a selection: ascending tet index, or (preferGeomOrdering) entity by entity in the order of the
    /// selection's positions and inside an entity by primitive type, then `leid`
    std::vector<LocalOrder> DofsBySparsity(const TetGeomSparsity& sp, bool preferGeomOrdering = false) const {
        std::vector<LocalOrder> r;
        for (auto p = sp.beginPos(); p != sp.endPos(); p = sp.nextPos(p))
            for (int t = 0; t < NGEOM_TYPES; ++t) {
                if (NumToGeomDim(t) != p.elem_dim) continue;
                const uchar et = NumToGeomType(t);
                for (uint k = 0; k < NumDof(et); ++k) r.push_back(LocalOrderOnTet(LocGeomOrder(et, p.elem_num, k)));
            }
        if (!preferGeomOrdering) std::sort(r.begin(), r.end(), [](const LocalOrder& a, const LocalOrder& b) { return a.gid < b.gid; });
        return r;
    }
    struct DofSparsedIterator {
        std::shared_ptr<std::vector<LocalOrder>> list;
        std::size_t k = 0;
        const LocalOrder& operator*() const { return (*list)[k]; }
        const LocalOrder* operator->() const { return &(*list)[k]; }
        DofSparsedIterator& operator++() { ++k; return *this; }
        bool atEnd() const { return !list || k >= list->size(); }
        bool operator==(const DofSparsedIterator& o) const { return (atEnd() && o.atEnd()) || (list == o.list && k == o.k); }
        bool operator!=(const DofSparsedIterator& o) const { return !(*this == o); }
    };
    DofSparsedIterator beginBySparsity(const TetGeomSparsity& sp, bool preferGeomOrdering = false) const {
        return DofSparsedIterator{std::make_shared<std::vector<LocalOrder>>(DofsBySparsity(sp, preferGeomOrdering)), 0};
    }
    DofSparsedIterator endBySparsity() const { return DofSparsedIterator{}; }
};

/// n[t] dofs on every entity of primitive type t (tetdofmap.h:592-620)
struct UniteDofMap : public BaseDofMap {
    std::array<uint, NGEOM_TYPES> m_n{{0, 0, 0, 0, 0, 0}};
    UniteDofMap() = default;
    explicit UniteDofMap(std::array<uint, NGEOM_TYPES> n) : m_n(n) {}
    uint ActualType() const override { return static_cast<uint>(BaseTypes::UniteType); }
    uint NumDof(uchar etype) const override { const int t = GeomTypeToNum(etype); return t < 0 ? 0 : m_n[t]; }
    LocalOrder LocalOrderOnTet(TetOrder d) const override {
        uint g = d.gid;
        for (int t = 0; t < NGEOM_TYPES; ++t) {
            const uchar et = NumToGeomType(t);
            const uint block = m_n[t] * GeomTypeTetElems(et);
            if (g < block) return LocalOrder(d.gid, et, static_cast<uchar>(g / m_n[t]), g % m_n[t]);
            g -= block;
        }
        throw std::out_of_range("Not valid index");
    }
    uint TetDofID(LocGeomOrder q) const override {
        const int t = GeomTypeToNum(q.etype);
        if (t < 0) throw std::out_of_range("Not valid index");
        uint off = 0;
        for (int s = 0; s < t; ++s) off += m_n[s] * GeomTypeTetElems(NumToGeomType(s));
        return off + q.nelem * m_n[t] + q.leid;
    }
    uint NestedDim() const override { return 0; }
    std::shared_ptr<BaseDofMap> GetSubDofMap(const int*, int ndims) const override { return ndims == 0 ? Copy() : nullptr; }
    bool GetNestedShift(const int*, int ndims, uint&, std::array<uint, NGEOM_TYPES>&) const override { return ndims == 0; }
    bool operator==(const BaseDofMap& o) const override { return o.ActualType() == ActualType() && static_cast<const UniteDofMap&>(o).m_n == m_n; }
    std::shared_ptr<BaseDofMap> Copy() const override { return std::make_shared<UniteDofMap>(*this); }
};

/// concatenation of component maps; VectorDofMap is the case of k equal components (tetdofmap.h:621-767)
struct ComplexDofMap : public BaseDofMap {
    std::vector<std::shared_ptr<BaseDofMap>> m_spaces;
    ComplexDofMap() = default;
    explicit ComplexDofMap(std::vector<std::shared_ptr<BaseDofMap>> spaces) : m_spaces(std::move(spaces)) {}
    /// runs of equal neighbours become VectorDofMaps (tetdofmap.cpp:563-588)
    static ComplexDofMap makeCompressed(const std::vector<std::shared_ptr<BaseDofMap>>& spaces);
    uint ActualType() const override { return static_cast<uint>(BaseTypes::ComplexType); }
    uint NumDof(uchar etype) const override { uint s = 0; for (auto& m : m_spaces) s += m->NumDof(etype); return s; }
    LocalOrder LocalOrderOnTet(TetOrder d) const override {
        uint g = d.gid;
        std::array<uint, NGEOM_TYPES> lshift{{0, 0, 0, 0, 0, 0}};
        for (auto& m : m_spaces) {
            const uint n = m->NumDofOnTet();
            if (g < n) {
                LocalOrder lo = m->LocalOrderOnTet(TetOrder(g));
                lo.gid = d.gid;
                lo.leid += lshift[GeomTypeToNum(lo.etype)];
                return lo;
            }
            g -= n;
            for (int t = 0; t < NGEOM_TYPES; ++t) lshift[t] += m->NumDof(NumToGeomType(t));
        }
        throw std::out_of_range("Not valid index");
    }
    uint TetDofID(LocGeomOrder q) const override {
        uint off = 0, leid = q.leid;
        for (auto& m : m_spaces) {
            const uint n = m->NumDof(q.etype);
            if (leid < n) return off + m->TetDofID(LocGeomOrder(q.etype, q.nelem, leid));
            leid -= n;
            off += m->NumDofOnTet();
        }
        throw std::out_of_range("Not valid index");
    }
    uint NestedDim() const override { return static_cast<uint>(m_spaces.size()); }
    std::shared_ptr<BaseDofMap> GetSubDofMap(const int* ext_dims, int ndims) const override {
        if (ndims == 0) return Copy();
        if (ext_dims[0] < 0 || ext_dims[0] >= static_cast<int>(m_spaces.size())) return nullptr;
        return ndims == 1 ? m_spaces[ext_dims[0]] : m_spaces[ext_dims[0]]->GetSubDofMap(ext_dims + 1, ndims - 1);
    }
    bool GetNestedShift(const int* ext_dims, int ndims, uint& shift_on_tet, std::array<uint, NGEOM_TYPES>& shift_leid) const override {
        if (ndims == 0) return true;
        if (ext_dims[0] < 0 || ext_dims[0] >= static_cast<int>(m_spaces.size())) return false;
        for (int c = 0; c < ext_dims[0]; ++c) {
            shift_on_tet += m_spaces[c]->NumDofOnTet();
            for (int t = 0; t < NGEOM_TYPES; ++t) shift_leid[t] += m_spaces[c]->NumDof(NumToGeomType(t));
        }
        return m_spaces[ext_dims[0]]->GetNestedShift(ext_dims + 1, ndims - 1, shift_on_tet, shift_leid);
    }
    bool operator==(const BaseDofMap& o) const override {
        if (o.ActualType() != ActualType()) return false;
        const auto& c = static_cast<const ComplexDofMap&>(o);
        if (c.m_spaces.size() != m_spaces.size()) return false;
        for (std::size_t k = 0; k < m_spaces.size(); ++k)
            if (m_spaces[k].get() != c.m_spaces[k].get() && *m_spaces[k] != *c.m_spaces[k]) return false;
        return true;
    }
    std::shared_ptr<BaseDofMap> Copy() const override { return std::make_shared<ComplexDofMap>(*this); }
};

struct VectorDofMap : public BaseDofMap {
    uint m_dim = 0;
    std::shared_ptr<BaseDofMap> base;
    VectorDofMap() = default;
    VectorDofMap(uint dim, std::shared_ptr<BaseDofMap> b) : m_dim(dim), base(std::move(b)) {}
    uint ActualType() const override { return static_cast<uint>(BaseTypes::VectorType); }
    uint NumDof(uchar etype) const override { return m_dim * base->NumDof(etype); }
    LocalOrder LocalOrderOnTet(TetOrder d) const override {
        const uint n = base->NumDofOnTet();
        if (n == 0 || d.gid >= m_dim * n) throw std::out_of_range("Not valid index");
        const uint c = d.gid / n;
        LocalOrder lo = base->LocalOrderOnTet(TetOrder(d.gid - c * n));
        lo.gid = d.gid;
        lo.leid += c * base->NumDof(lo.etype);
        return lo;
    }
    uint TetDofID(LocGeomOrder q) const override {
        const uint n = base->NumDof(q.etype);
        if (n == 0 || q.leid >= m_dim * n) throw std::out_of_range("Not valid index");
        const uint c = q.leid / n;
        return c * base->NumDofOnTet() + base->TetDofID(LocGeomOrder(q.etype, q.nelem, q.leid - c * n));
    }
    uint NestedDim() const override { return m_dim; }
    std::shared_ptr<BaseDofMap> GetSubDofMap(const int* ext_dims, int ndims) const override {
        if (ndims == 0) return Copy();
        if (ext_dims[0] < 0 || ext_dims[0] >= static_cast<int>(m_dim)) return nullptr;
        return ndims == 1 ? base : base->GetSubDofMap(ext_dims + 1, ndims - 1);
    }
    bool GetNestedShift(const int* ext_dims, int ndims, uint& shift_on_tet, std::array<uint, NGEOM_TYPES>& shift_leid) const override {
        if (ndims == 0) return true;
        if (ext_dims[0] < 0 || ext_dims[0] >= static_cast<int>(m_dim)) return false;
        shift_on_tet += ext_dims[0] * base->NumDofOnTet();
        for (int t = 0; t < NGEOM_TYPES; ++t) shift_leid[t] += ext_dims[0] * base->NumDof(NumToGeomType(t));
        return base->GetNestedShift(ext_dims + 1, ndims - 1, shift_on_tet, shift_leid);
    }
    bool operator==(const BaseDofMap& o) const override {
        if (o.ActualType() != ActualType()) return false;
        const auto& v = static_cast<const VectorDofMap&>(o);
        return v.m_dim == m_dim && (v.base.get() == base.get() || *v.base == *base);
    }
    std::shared_ptr<BaseDofMap> Copy() const override { return std::make_shared<VectorDofMap>(*this); }
};

inline ComplexDofMap ComplexDofMap::makeCompressed(const std::vector<std::shared_ptr<BaseDofMap>>& spaces) {
    std::vector<std::shared_ptr<BaseDofMap>> out;
    for (std::size_t i = 0; i < spaces.size();) {
        std::size_t j = i + 1;
        while (j < spaces.size() && (spaces[j].get() == spaces[i].get() || *spaces[j] == *spaces[i])) ++j;
        if (j - i == 1) out.push_back(spaces[i]);
        else out.push_back(std::make_shared<VectorDofMap>(static_cast<uint>(j - i), spaces[i]));
        i = j;
    }
    return ComplexDofMap(out);
}

/// value-semantics handle (tetdofmap.h:492-590)
struct DofMap {
    std::shared_ptr<BaseDofMap> m_invoker;
    DofMap() = default;
    explicit DofMap(std::shared_ptr<BaseDofMap> m) : m_invoker(std::move(m)) {}
    template <typename DofMapT, typename = typename std::enable_if<std::is_base_of<BaseDofMap, DofMapT>::value>::type>
    explicit DofMap(const DofMapT& m) : m_invoker(std::make_shared<DofMapT>(m)) {}
    std::shared_ptr<BaseDofMap> base() const { return m_invoker; }
    template <typename DofMapT = BaseDofMap> DofMapT* target() { return static_cast<DofMapT*>(m_invoker.get()); }
    template <typename DofMapT = BaseDofMap> const DofMapT* target() const { return static_cast<const DofMapT*>(m_invoker.get()); }
    bool isValid() const { return static_cast<bool>(m_invoker); }
    uint ActualType() const { return m_invoker->ActualType(); }
    uint NumDof(uchar etype) const { return m_invoker->NumDof(etype); }
    uint NumDofOnTet(uchar etype) const { return m_invoker->NumDofOnTet(etype); }
    uint NumDofOnTet() const { return m_invoker->NumDofOnTet(); }
    std::array<uint, NGEOM_TYPES> NumDofs() const { return m_invoker->NumDofs(); }
    std::array<uint, NGEOM_TYPES> NumDofsOnTet() const { return m_invoker->NumDofsOnTet(); }
    LocalOrder LocalOrderOnTet(TetOrder d) const { return m_invoker->LocalOrderOnTet(d); }
    LocalOrder LocalOrderOnTet(LocGeomOrder g) const { return m_invoker->LocalOrderOnTet(g); }
    uint TetDofID(LocGeomOrder g) const { return m_invoker->TetDofID(g); }
    uint TypeOnTet(uint dof) const { return m_invoker->TypeOnTet(dof); }
    bool DefinedOn(uchar etype) const { return m_invoker->DefinedOn(etype); }
    uint GetGeomMask() const { return m_invoker->GetGeomMask(); }
    uint NestedDim() const { return m_invoker->NestedDim(); }
    LocalOrder operator[](uint d) const { return (*m_invoker)[d]; }
    LocalOrder operator[](LocGeomOrder g) const { return (*m_invoker)[g]; }
    LocalOrder at(TetOrder d) const { return m_invoker->at(d); }
    LocalOrder at(LocGeomOrder g) const { return m_invoker->at(g); }
    BaseDofMap::DofIterator begin() const { return m_invoker->begin(); }
    BaseDofMap::DofIterator end() const { return m_invoker->end(); }
    BaseDofMap::DofSparsedIterator beginBySparsity(const TetGeomSparsity& sp, bool preferGeomOrdering = false) const { return m_invoker->beginBySparsity(sp, preferGeomOrdering); }
    BaseDofMap::DofSparsedIterator endBySparsity() const { return m_invoker->endBySparsity(); }
    DofMap GetSubDofMap(const int* ext_dims, int ndims) const { return DofMap(m_invoker->GetSubDofMap(ext_dims, ndims)); }
    bool operator==(const DofMap& o) const { return m_invoker.get() == o.m_invoker.get() || (m_invoker && o.m_invoker && *m_invoker == *o.m_invoker); }
    bool operator!=(const DofMap& o) const { return !(*this == o); }
    /// product of spaces with the simplifications of the reference (tetdofmap.cpp:814-890): equal factors fuse into vectors,
    /// complex factors are spliced, a vector absorbs a neighbour equal to its base
    DofMap operator*(const DofMap& o) const;
};
inline DofMap pow(const DofMap& d, uint k) { return DofMap(std::make_shared<VectorDofMap>(k, d.base())); }
inline DofMap operator^(const DofMap& d, uint k) {
    if (d.ActualType() == static_cast<uint>(BaseDofMap::BaseTypes::VectorType)) {
        auto v = d.target<VectorDofMap>();
        return DofMap(std::make_shared<VectorDofMap>(v->m_dim * k, v->base));
    }
    return pow(d, k);
}
/// concatenation without simplification
inline DofMap merge(const std::vector<DofMap>& maps) {
    std::vector<std::shared_ptr<BaseDofMap>> s;
    for (auto& m : maps) s.push_back(m.base());
    return DofMap(std::make_shared<ComplexDofMap>(s));
}
inline DofMap merge_with_simplifications(const std::vector<DofMap>& maps) {
    if (maps.empty()) return DofMap();
    DofMap r = maps[0];
    for (std::size_t k = 1; k < maps.size(); ++k) r = r * maps[k];
    return r;
}
inline DofMap DofMap::operator*(const DofMap& o) const {
    const uint VEC = static_cast<uint>(BaseDofMap::BaseTypes::VectorType), CPX = static_cast<uint>(BaseDofMap::BaseTypes::ComplexType);
    auto same = [](const std::shared_ptr<BaseDofMap>& a, const std::shared_ptr<BaseDofMap>& b) { return a.get() == b.get() || (a && b && *a == *b); };
    auto parts = [&](const DofMap& m) {   // first-level factors of a complex map, else the map itself
        return m.ActualType() == CPX ? m.target<ComplexDofMap>()->m_spaces : std::vector<std::shared_ptr<BaseDofMap>>{m.base()};
    };
    const uint t1 = ActualType(), t2 = o.ActualType();
    if (t1 != CPX && t2 != CPX) {   // two simple factors
        if (t1 == t2 && *this == o) {
            if (t1 == VEC) return DofMap(std::make_shared<VectorDofMap>(target<VectorDofMap>()->m_dim + o.target<VectorDofMap>()->m_dim, target<VectorDofMap>()->base));
            return DofMap(std::make_shared<VectorDofMap>(2, base()));
        }
        if (t1 == VEC && t2 != VEC && same(target<VectorDofMap>()->base, o.base()))
            return DofMap(std::make_shared<VectorDofMap>(target<VectorDofMap>()->m_dim + 1, target<VectorDofMap>()->base));
        if (t2 == VEC && t1 != VEC && same(o.target<VectorDofMap>()->base, base()))
            return DofMap(std::make_shared<VectorDofMap>(o.target<VectorDofMap>()->m_dim + 1, o.target<VectorDofMap>()->base));
        return DofMap(std::make_shared<ComplexDofMap>(std::vector<std::shared_ptr<BaseDofMap>>{base(), o.base()}));
    }
    if (t1 == CPX && t2 == CPX && *this == o) return DofMap(std::make_shared<VectorDofMap>(2, base()));
    // at least one complex factor: splice, fusing the two factors that meet
    std::vector<std::shared_ptr<BaseDofMap>> a = parts(*this), b = parts(o), u;
    if (a.empty() || b.empty()) {
        u = a;
        u.insert(u.end(), b.begin(), b.end());
        return DofMap(std::make_shared<ComplexDofMap>(u));
    }
    u.assign(a.begin(), a.end() - 1);
    const DofMap mid = DofMap(a.back()) * DofMap(b.front());
    const auto midp = parts(mid);
    u.insert(u.end(), midp.begin(), midp.end());
    u.insert(u.end(), b.begin() + 1, b.end());
    return DofMap(std::make_shared<ComplexDofMap>(u));
}

}  // namespace DofT
}  // namespace Ani
