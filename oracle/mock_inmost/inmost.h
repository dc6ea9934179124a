// TEST INFRASTRUCTURE ONLY.  A small serial stand-in for the part of the INMOST API that AniFem++'s `inmost_interface`
// touches (INMOST itself is an un-vendored external: cmake/Downloadinmost.cmake:4, commit f3392cef...).  It exists so that the
// reference's OWN, UNMODIFIED assembler sources (anifem++/inmost_interface/{assembler.inl, global_enumerator.cpp, ordering.inl,
// elemental_assembler.cpp}) compile and run in this container and pin oracle/asm_oracle.py (and through it the GPU path) at the
// assembler level: dof numbering of all six GlobEnumeration types, fill_assemble_templates, AssembleTemplate, the scatter with
// drop_val / signs, status codes.  Nothing here is INMOST code; the names follow INMOST's public interface because the reference
// includes "inmost.h".  One rank only (GetProcessorsNumber() == 1, every element Owned): the MPI side of INMOST is not modelled.
//
// Conventions standing in for INMOST (the same ones oracle/asm_oracle.py and the product use):
//   * nodes keep their creation order; edges / faces are created in lexicographic order of their sorted node tuples; cells in
//     creation order; LocalID == GlobalID == that position;
//   * cell -> nodes (HighConn) in the order given at creation; cell -> faces (LowConn): face k = nodes {k, k+1, k+2} mod 4;
//     face -> edges (LowConn): edges (n0 n1), (n1 n2), (n2 n0) of its sorted nodes; edge -> nodes (LowConn): sorted pair.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <iostream>
#include <iterator>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define INMOST_DATA_ENUM_TYPE unsigned int
#define INMOST_DATA_REAL_TYPE double
#define INMOST_DATA_INTEGER_TYPE int
#define INMOST_DATA_BULK_TYPE unsigned char
#define INMOST_MPI_COMM_WORLD 0
#define ENUMUNDEF (~(INMOST_DATA_ENUM_TYPE)0)

inline double Timer() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace INMOST {

typedef unsigned int HandleType;
typedef unsigned char ElementType;
typedef unsigned char DataType;
typedef unsigned char MarkerType;
static const ElementType NONE = 0x00, NODE = 0x01, EDGE = 0x02, FACE = 0x04, CELL = 0x08, ESET = 0x10, MESH = 0x20;
static const DataType DATA_REAL = 0, DATA_INTEGER = 1, DATA_BULK = 2, DATA_REFERENCE = 3;
static const HandleType InvalidHandle() { return 0; }

inline int ElementNum(ElementType t) { int n = 0; while (t > 1) { t >>= 1; ++n; } return n; }
inline ElementType ElementTypeFromDim(int d) { return (ElementType)(1u << d); }
inline HandleType ComposeHandle(ElementType et, int id) { return ((HandleType)ElementNum(et) << 29) + (HandleType)(id + 1); }
inline int GetHandleID(HandleType h) { return (int)(h & 0x1fffffffu) - 1; }
inline int GetHandleElementNum(HandleType h) { return (int)(h >> 29); }
inline ElementType GetHandleElementType(HandleType h) { return h == 0 ? NONE : (ElementType)(1u << GetHandleElementNum(h)); }
inline bool isValidHandle(HandleType h) { return h != 0; }


class Mesh;
class Element;
class Node;
class Edge;
class Face;
class Cell;
class Storage;

// view of a contiguous array owned by the mesh (tag data, adjacency, coordinates)
template <typename T>
class shell {
    std::vector<T>* v = nullptr;   // resizable storage, or
    T* p = nullptr;                // fixed view
    std::size_t n = 0;
public:
    typedef T* iterator;
    typedef const T* const_iterator;
    shell() = default;
    explicit shell(std::vector<T>& vec) : v(&vec) {}
    shell(T* ptr, std::size_t sz) : p(ptr), n(sz) {}
    std::size_t size() const { return v ? v->size() : n; }
    bool empty() const { return size() == 0; }
    T* data() { return v ? v->data() : p; }
    const T* data() const { return v ? v->data() : p; }
    T& operator[](std::size_t i) { return data()[i]; }
    const T& operator[](std::size_t i) const { return data()[i]; }
    T& at(std::size_t i) { if (i >= size()) throw std::out_of_range("shell"); return data()[i]; }
    iterator begin() { return data(); }
    iterator end() { return data() + size(); }
    const_iterator begin() const { return data(); }
    const_iterator end() const { return data() + size(); }
    void resize(std::size_t sz, T val = T()) { if (!v) throw std::runtime_error("mock inmost: fixed-size array"); v->resize(sz, val); }
    void push_back(const T& x) { if (!v) throw std::runtime_error("mock inmost: fixed-size array"); v->push_back(x); }
    void clear() { resize(0); }
    T& back() { return data()[size() - 1]; }
};

struct TagMemory {
    std::string name;
    DataType dtype;
    ElementType defined, sparse;
    INMOST_DATA_ENUM_TYPE size;    // ENUMUNDEF: variable size
    Mesh* mesh;
    // data[element type number][local id] -> values
    std::array<std::vector<std::vector<double>>, 6> rdata;
    std::array<std::vector<std::vector<int>>, 6> idata;
};

class Tag {
    std::shared_ptr<TagMemory> mem;
    friend class Mesh;
public:
    Tag() = default;
    bool isValid() const { return (bool)mem; }
    INMOST_DATA_ENUM_TYPE GetSize() const { return mem->size; }
    DataType GetDataType() const { return mem->dtype; }
    bool isDefined(ElementType t) const { return (mem->defined & t) != 0; }
    bool isDefinedMask(ElementType mask) const { return (mem->defined & mask) == mask; }
    bool isSparse(ElementType t) const { return (mem->sparse & t) != 0; }
    Mesh* GetMeshLink() const { return mem ? mem->mesh : nullptr; }
    std::string GetTagName() const { return mem->name; }
    bool operator==(const Tag& o) const { return mem == o.mem; }
    bool operator!=(const Tag& o) const { return mem != o.mem; }
    TagMemory* raw() const { return mem.get(); }
};

class Storage {
protected:
    Mesh* m_link = nullptr;
    HandleType handle = 0;
public:
    typedef INMOST_DATA_REAL_TYPE real;
    typedef INMOST_DATA_INTEGER_TYPE integer;
    typedef INMOST_DATA_BULK_TYPE bulk;
    typedef INMOST_DATA_ENUM_TYPE enumerator;
    typedef shell<real> real_array;
    typedef shell<integer> integer_array;
    typedef shell<bulk> bulk_array;
    Storage() = default;
    Storage(Mesh* m, HandleType h) : m_link(m), handle(h) {}
    HandleType GetHandle() const { return handle; }
    Mesh* GetMeshLink() const { return m_link; }
    bool isValid() const { return m_link != nullptr && handle != 0; }
    ElementType GetElementType() const { return GetHandleElementType(handle); }
    integer GetElementNum() const { return GetHandleElementNum(handle); }
    integer GetElementDimension() const { return GetHandleElementNum(handle); }   // 3D mesh: node 0, edge 1, face 2, cell 3
    integer LocalID() const { return GetHandleID(handle); }
    integer DataLocalID() const { return GetHandleID(handle); }
    inline integer GlobalID() const;
    inline real_array RealArray(const Tag& t) const;
    inline integer_array IntegerArray(const Tag& t) const;
    inline real& Real(const Tag& t) const;
    inline integer& Integer(const Tag& t) const;
    inline bool HaveData(const Tag& t) const;
    Storage* operator->() { return this; }
    const Storage* operator->() const { return this; }
};

template <typename StorageType>
class ElementArray {
    Mesh* m_link = nullptr;
    std::vector<HandleType> container;
public:
    typedef std::size_t size_type;
    ElementArray() = default;
    explicit ElementArray(Mesh* m) : m_link(m) {}
    ElementArray(Mesh* m, size_type n, HandleType h = 0) : m_link(m), container(n, h) {}
    ElementArray(Mesh* m, const HandleType* b, const HandleType* e) : m_link(m), container(b, e) {}
    size_type size() const { return container.size(); }
    bool empty() const { return container.empty(); }
    void resize(size_type n, HandleType h = 0) { container.resize(n, h); }
    void clear() { container.clear(); }
    void reserve(size_type n) { container.reserve(n); }
    HandleType* data() { return container.data(); }
    const HandleType* data() const { return container.data(); }
    HandleType& at(size_type i) { return container.at(i); }
    const HandleType& at(size_type i) const { return container.at(i); }
    StorageType operator[](size_type i) const { return StorageType(m_link, container[i]); }
    void push_back(const Storage& e) { container.push_back(e.GetHandle()); }
    void push_back(HandleType h) { container.push_back(h); }
    Mesh* GetMeshLink() const { return m_link; }
    void SetMeshLink(Mesh* m) { m_link = m; }
    class iterator {
        const ElementArray* a; size_type i;
    public:
        iterator(const ElementArray* arr, size_type k) : a(arr), i(k) {}
        StorageType operator*() const { return (*a)[i]; }
        StorageType operator->() const { return (*a)[i]; }
        iterator& operator++() { ++i; return *this; }
        bool operator!=(const iterator& o) const { return i != o.i; }
        bool operator==(const iterator& o) const { return i == o.i; }
    };
    iterator begin() const { return iterator(this, 0); }
    iterator end() const { return iterator(this, container.size()); }
};

class Element : public Storage {
public:
    typedef INMOST_DATA_BULK_TYPE Status;
    static const Status Owned = 1, Shared = 2, Ghost = 4, Any = 0;
    typedef shell<HandleType> adj_type;
    Element() = default;
    Element(Mesh* m, HandleType h) : Storage(m, h) {}
    Element(const Storage& s) : Storage(s) {}
    Status GetStatus() const { return Owned; }
    bool Hidden() const { return false; }
    inline ElementArray<Node> getNodes() const;
    inline ElementArray<Edge> getEdges() const;
    inline ElementArray<Face> getFaces() const;
    inline ElementArray<Cell> getCells() const;
    inline Node getAsNode() const;
    inline Edge getAsEdge() const;
    inline Face getAsFace() const;
    inline Cell getAsCell() const;
    Element getAsElement() const { return *this; }
    Element* operator->() { return this; }
    const Element* operator->() const { return this; }
    bool operator==(const Element& o) const { return handle == o.handle && m_link == o.m_link; }
    bool operator!=(const Element& o) const { return !(*this == o); }
};
class Node : public Element {
public:
    Node() = default;
    Node(Mesh* m, HandleType h) : Element(m, h) {}
    Node(const Element& e) : Element(e) {}
    inline Storage::real_array Coords() const;
    Node* operator->() { return this; }
    const Node* operator->() const { return this; }
};
class Edge : public Element {
public:
    Edge() = default;
    Edge(Mesh* m, HandleType h) : Element(m, h) {}
    Edge(const Element& e) : Element(e) {}
    inline Node getBeg() const;
    inline Node getEnd() const;
    Edge* operator->() { return this; }
    const Edge* operator->() const { return this; }
};
class Face : public Element {
public:
    Face() = default;
    Face(Mesh* m, HandleType h) : Element(m, h) {}
    Face(const Element& e) : Element(e) {}
    Face* operator->() { return this; }
    const Face* operator->() const { return this; }
};
class Cell : public Element {
public:
    Cell() = default;
    Cell(Mesh* m, HandleType h) : Element(m, h) {}
    Cell(const Element& e) : Element(e) {}
    Cell* operator->() { return this; }
    const Cell* operator->() const { return this; }
};

class Mesh : public Storage {
    std::vector<std::array<double, 3>> coords;
    // adjacency per element type number: low = towards nodes, high = towards cells
    std::array<std::vector<std::vector<HandleType>>, 4> low, high;
    std::array<int, 6> count{{0, 0, 0, 0, 0, 1}};
    std::map<std::string, Tag> tags;
    Tag gid_tag;
    ElementType have_gid = NONE;
public:
    Mesh() : Storage(this, ComposeHandle(MESH, 0)) {}
    Mesh(const Mesh&) = delete;
    // ---- construction (mock only): all nodes, then all tetrahedra
    void BuildTets(long nnode, const double* xyz /*nnode x 3*/, long ntet, const long* tets /*ntet x 4*/) {
        coords.resize(nnode);
        for (long i = 0; i < nnode; ++i) coords[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        count[0] = (int)nnode; count[3] = (int)ntet;
        std::vector<std::array<long, 2>> ek;
        std::vector<std::array<long, 3>> fk;
        for (long e = 0; e < ntet; ++e) {
            const long* t = tets + 4 * e;
            for (int a = 0; a < 4; ++a)
                for (int b = a + 1; b < 4; ++b) ek.push_back({std::min(t[a], t[b]), std::max(t[a], t[b])});
            for (int k = 0; k < 4; ++k) { std::array<long, 3> f{t[k], t[(k + 1) % 4], t[(k + 2) % 4]}; std::sort(f.begin(), f.end()); fk.push_back(f); }
        }
        std::sort(ek.begin(), ek.end()); ek.erase(std::unique(ek.begin(), ek.end()), ek.end());
        std::sort(fk.begin(), fk.end()); fk.erase(std::unique(fk.begin(), fk.end()), fk.end());
        count[1] = (int)ek.size(); count[2] = (int)fk.size();
        for (int d = 0; d < 4; ++d) { low[d].assign(count[d], {}); high[d].assign(count[d], {}); }
        auto edge_id = [&](long a, long b) { std::array<long, 2> k{std::min(a, b), std::max(a, b)}; return (int)(std::lower_bound(ek.begin(), ek.end(), k) - ek.begin()); };
        auto face_id = [&](std::array<long, 3> k) { std::sort(k.begin(), k.end()); return (int)(std::lower_bound(fk.begin(), fk.end(), k) - fk.begin()); };
        for (int i = 0; i < count[1]; ++i) {
            low[1][i] = {ComposeHandle(NODE, (int)ek[i][0]), ComposeHandle(NODE, (int)ek[i][1])};
            high[0][ek[i][0]].push_back(ComposeHandle(EDGE, i)); high[0][ek[i][1]].push_back(ComposeHandle(EDGE, i));
        }
        for (int i = 0; i < count[2]; ++i) {
            const auto& f = fk[i];
            const int e0 = edge_id(f[0], f[1]), e1 = edge_id(f[1], f[2]), e2 = edge_id(f[2], f[0]);
            low[2][i] = {ComposeHandle(EDGE, e0), ComposeHandle(EDGE, e1), ComposeHandle(EDGE, e2)};
            for (int e : {e0, e1, e2}) high[1][e].push_back(ComposeHandle(FACE, i));
        }
        for (long e = 0; e < ntet; ++e) {
            const long* t = tets + 4 * e;
            for (int k = 0; k < 4; ++k) {
                const int f = face_id({t[k], t[(k + 1) % 4], t[(k + 2) % 4]});
                low[3][e].push_back(ComposeHandle(FACE, f));
                high[2][f].push_back(ComposeHandle(CELL, (int)e));
                high[3][e].push_back(ComposeHandle(NODE, (int)t[k]));   // HighConn of a cell = its nodes (INMOST convention)
            }
        }
    }
    // ---- queries used by the reference
    int GetProcessorRank() const { return 0; }
    int GetProcessorsNumber() const { return 1; }
    Element::adj_type LowConn(HandleType h) { return Element::adj_type(low[GetHandleElementNum(h)][GetHandleID(h)]); }
    Element::adj_type HighConn(HandleType h) { return Element::adj_type(high[GetHandleElementNum(h)][GetHandleID(h)]); }
    Storage::real_array NodeCoords(HandleType h) { return Storage::real_array(coords[GetHandleID(h)].data(), 3); }
    Storage::integer NumberOf(ElementType t) const { Storage::integer n = 0; for (int d = 0; d < 4; ++d) if (t & (1 << d)) n += count[d]; return n; }
    Storage::integer TotalNumberOf(ElementType t) const { return NumberOf(t); }
    Storage::integer FirstLocalID(ElementType) const { return 0; }
    Storage::integer LastLocalID(ElementType t) const { return count[ElementNum(t)]; }
    Storage::integer LastLocalIDNum(int n) const { return count[n]; }
    Storage::integer NextLocalID(ElementType t, Storage::integer lid) const { (void)t; return lid + 1; }
    Storage::integer CellLastLocalID() const { return count[3]; }
    Storage::integer NodeLastLocalID() const { return count[0]; }
    Storage::integer EdgeLastLocalID() const { return count[1]; }
    Storage::integer FaceLastLocalID() const { return count[2]; }
    Element ElementByLocalID(ElementType t, Storage::integer lid) { return Element(this, lid >= 0 && lid < count[ElementNum(t)] ? ComposeHandle(t, lid) : 0); }
    Element ElementByLocalIDNum(int n, Storage::integer lid) { return ElementByLocalID((ElementType)(1 << n), lid); }
    Cell CellByLocalID(Storage::integer lid) { return Cell(this, lid >= 0 && lid < count[3] ? ComposeHandle(CELL, lid) : 0); }
    Node NodeByLocalID(Storage::integer lid) { return Node(this, ComposeHandle(NODE, lid)); }
    Edge EdgeByLocalID(Storage::integer lid) { return Edge(this, ComposeHandle(EDGE, lid)); }
    Face FaceByLocalID(Storage::integer lid) { return Face(this, ComposeHandle(FACE, lid)); }
    Element::Status GetStatus(HandleType) const { return Element::Owned; }
    bool Hidden(HandleType) const { return false; }
    bool isValidElement(HandleType h) const { return h != 0; }
    bool HaveGlobalID(ElementType t) const { return (have_gid & t) == t; }
    void AssignGlobalID(ElementType mask) { have_gid |= mask; }
    Storage::integer GlobalID(HandleType h) const { return GetHandleID(h); }   // serial: creation order
    Tag GlobalIDTag() {
        if (!gid_tag.isValid()) {
            gid_tag = CreateTag("GLOBAL_ID", DATA_INTEGER, NODE | EDGE | FACE | CELL, NONE, 1);
            for (int d = 0; d < 4; ++d) for (int i = 0; i < count[d]; ++i) gid_tag.mem->idata[d][i][0] = i;
        }
        return gid_tag;
    }
    // ---- tags
    Tag CreateTag(const std::string& name, DataType dtype, ElementType etype, ElementType sparse, INMOST_DATA_ENUM_TYPE size = ENUMUNDEF) {
        auto it = tags.find(name);
        if (it != tags.end()) return it->second;
        Tag t;
        t.mem = std::make_shared<TagMemory>();
        t.mem->name = name; t.mem->dtype = dtype; t.mem->defined = etype; t.mem->sparse = sparse; t.mem->size = size; t.mem->mesh = this;
        const std::size_t n0 = size == ENUMUNDEF ? 0 : size;
        for (int d = 0; d < 6; ++d) if (etype & (1 << d)) {
            if (dtype == DATA_REAL) t.mem->rdata[d].assign(count[d], std::vector<double>(n0, 0.0));
            else t.mem->idata[d].assign(count[d], std::vector<int>(n0, 0));
        }
        tags[name] = t;
        return t;
    }
    Tag DeleteTag(Tag t, ElementType mask = NODE | EDGE | FACE | CELL | ESET | MESH) { (void)mask; if (t.isValid()) tags.erase(t.GetTagName()); return Tag(); }
    bool HaveTag(const std::string& name) const { return tags.count(name) != 0; }
    Tag GetTag(const std::string& name) const { auto it = tags.find(name); if (it == tags.end()) throw std::runtime_error("mock inmost: no tag " + name); return it->second; }
    Storage::real_array RealArray(HandleType h, const Tag& t) { return Storage::real_array(t.raw()->rdata[GetHandleElementNum(h)][GetHandleID(h)]); }
    Storage::integer_array IntegerArray(HandleType h, const Tag& t) { return Storage::integer_array(t.raw()->idata[GetHandleElementNum(h)][GetHandleID(h)]); }
    Storage::real& Real(HandleType h, const Tag& t) { return t.raw()->rdata[GetHandleElementNum(h)][GetHandleID(h)][0]; }
    Storage::integer& Integer(HandleType h, const Tag& t) { return t.raw()->idata[GetHandleElementNum(h)][GetHandleID(h)][0]; }
    bool HaveData(HandleType h, const Tag& t) const { return (t.raw()->defined & GetHandleElementType(h)) != 0; }
    void ExchangeData(const Tag&, ElementType, MarkerType = 0) {}                 // one rank: nothing to exchange
    void ExchangeData(const std::vector<Tag>&, ElementType, MarkerType = 0) {}
    Storage::integer Integrate(Storage::integer x) const { return x; }
    Storage::real Integrate(Storage::real x) const { return x; }
    Storage::integer ExclusiveSum(Storage::integer) const { return 0; }
    Storage::integer AggregateMax(Storage::integer x) const { return x; }
    Storage::real AggregateMax(Storage::real x) const { return x; }
    Storage::integer AggregateMin(Storage::integer x) const { return x; }
    // ---- iteration
    template <typename T>
    class base_iterator {
        Mesh* m; ElementType mask; int d; int lid;
        void settle() { while (d < 4 && (!(mask & (1 << d)) || lid >= m->count[d])) { ++d; lid = 0; } }
    public:
        typedef std::forward_iterator_tag iterator_category;
        typedef T value_type;
        typedef std::ptrdiff_t difference_type;
        typedef T* pointer;
        typedef T reference;
        base_iterator(Mesh* mesh, ElementType msk, bool end) : m(mesh), mask(msk), d(end ? 4 : 0), lid(0) { if (!end) settle(); }
        T operator*() const { return T(m, ComposeHandle((ElementType)(1 << d), lid)); }
        T operator->() const { return T(m, ComposeHandle((ElementType)(1 << d), lid)); }
        base_iterator& operator++() { ++lid; settle(); return *this; }
        base_iterator operator++(int) { base_iterator t = *this; ++(*this); return t; }
        bool operator==(const base_iterator& o) const { return d == o.d && (d == 4 || lid == o.lid); }
        bool operator!=(const base_iterator& o) const { return !(*this == o); }
    };
    typedef base_iterator<Element> iteratorElement;
    typedef base_iterator<Node> iteratorNode;
    typedef base_iterator<Edge> iteratorEdge;
    typedef base_iterator<Face> iteratorFace;
    typedef base_iterator<Cell> iteratorCell;
    iteratorElement BeginElement(ElementType mask) { return iteratorElement(this, mask, false); }
    iteratorElement EndElement() { return iteratorElement(this, NONE, true); }
    iteratorNode BeginNode() { return iteratorNode(this, NODE, false); }
    iteratorNode EndNode() { return iteratorNode(this, NONE, true); }
    iteratorEdge BeginEdge() { return iteratorEdge(this, EDGE, false); }
    iteratorEdge EndEdge() { return iteratorEdge(this, NONE, true); }
    iteratorFace BeginFace() { return iteratorFace(this, FACE, false); }
    iteratorFace EndFace() { return iteratorFace(this, NONE, true); }
    iteratorCell BeginCell() { return iteratorCell(this, CELL, false); }
    iteratorCell EndCell() { return iteratorCell(this, NONE, true); }
};

inline Storage::integer Storage::GlobalID() const { return m_link->GlobalID(handle); }
inline Storage::real_array Storage::RealArray(const Tag& t) const { return m_link->RealArray(handle, t); }
inline Storage::integer_array Storage::IntegerArray(const Tag& t) const { return m_link->IntegerArray(handle, t); }
inline Storage::real& Storage::Real(const Tag& t) const { return m_link->Real(handle, t); }
inline Storage::integer& Storage::Integer(const Tag& t) const { return m_link->Integer(handle, t); }
inline bool Storage::HaveData(const Tag& t) const { return m_link->HaveData(handle, t); }
inline Storage::real_array Node::Coords() const { return m_link->NodeCoords(handle); }
inline Node Element::getAsNode() const { return Node(m_link, handle); }
inline Edge Element::getAsEdge() const { return Edge(m_link, handle); }
inline Face Element::getAsFace() const { return Face(m_link, handle); }
inline Cell Element::getAsCell() const { return Cell(m_link, handle); }
inline Node Edge::getBeg() const { return Node(m_link, m_link->LowConn(handle)[0]); }
inline Node Edge::getEnd() const { return Node(m_link, m_link->LowConn(handle)[1]); }
namespace mock_detail {
inline void collect(Mesh* m, HandleType h, int target, std::vector<HandleType>& out) {
    const int d = GetHandleElementNum(h);
    if (d == target) { if (std::find(out.begin(), out.end(), h) == out.end()) out.push_back(h); return; }
    if (d == 3 && target == 0) { auto hc = m->HighConn(h); for (std::size_t i = 0; i < hc.size(); ++i) collect(m, hc[i], 0, out); return; }
    if (d > target) { auto lc = m->LowConn(h); for (std::size_t i = 0; i < lc.size(); ++i) collect(m, lc[i], target, out); }
    else { auto hc = m->HighConn(h); for (std::size_t i = 0; i < hc.size(); ++i) if (GetHandleElementNum(hc[i]) > d) collect(m, hc[i], target, out); }
}
}  // namespace mock_detail
inline ElementArray<Node> Element::getNodes() const { std::vector<HandleType> v; mock_detail::collect(m_link, handle, 0, v); return ElementArray<Node>(m_link, v.data(), v.data() + v.size()); }
inline ElementArray<Edge> Element::getEdges() const { std::vector<HandleType> v; mock_detail::collect(m_link, handle, 1, v); return ElementArray<Edge>(m_link, v.data(), v.data() + v.size()); }
inline ElementArray<Face> Element::getFaces() const { std::vector<HandleType> v; mock_detail::collect(m_link, handle, 2, v); return ElementArray<Face>(m_link, v.data(), v.data() + v.size()); }
inline ElementArray<Cell> Element::getCells() const { std::vector<HandleType> v; mock_detail::collect(m_link, handle, 3, v); return ElementArray<Cell>(m_link, v.data(), v.data() + v.size()); }

namespace Sparse {

// row of an INMOST matrix: unsorted (index, value) pairs, operator[] = find or append (the reference's default scatter path)
class Row {
public:
    struct entry {
        INMOST_DATA_ENUM_TYPE first;
        INMOST_DATA_REAL_TYPE second;
        entry() : first(ENUMUNDEF), second(0.0) {}
        entry(INMOST_DATA_ENUM_TYPE i, INMOST_DATA_REAL_TYPE v) : first(i), second(v) {}
        bool operator<(const entry& o) const { return first < o.first; }
    };
    typedef std::vector<entry>::iterator iterator;
    typedef std::vector<entry>::const_iterator const_iterator;
private:
    std::vector<entry> data;
public:
    INMOST_DATA_REAL_TYPE& operator[](INMOST_DATA_ENUM_TYPE i) {
        for (auto& e : data) if (e.first == i) return e.second;
        data.push_back(entry(i, 0.0));
        return data.back().second;
    }
    INMOST_DATA_REAL_TYPE get_safe(INMOST_DATA_ENUM_TYPE i) const { for (auto& e : data) if (e.first == i) return e.second; return 0.0; }
    INMOST_DATA_ENUM_TYPE Size() const { return (INMOST_DATA_ENUM_TYPE)data.size(); }
    bool Empty() const { return data.empty(); }
    void Clear() { data.clear(); }
    void Resize(INMOST_DATA_ENUM_TYPE n) { data.resize(n); }
    void Push(INMOST_DATA_ENUM_TYPE i, INMOST_DATA_REAL_TYPE v) { data.push_back(entry(i, v)); }
    INMOST_DATA_ENUM_TYPE& GetIndex(INMOST_DATA_ENUM_TYPE k) { return data[k].first; }
    INMOST_DATA_REAL_TYPE& GetValue(INMOST_DATA_ENUM_TYPE k) { return data[k].second; }
    INMOST_DATA_ENUM_TYPE GetIndex(INMOST_DATA_ENUM_TYPE k) const { return data[k].first; }
    INMOST_DATA_REAL_TYPE GetValue(INMOST_DATA_ENUM_TYPE k) const { return data[k].second; }
    iterator Begin() { return data.begin(); }
    iterator End() { return data.end(); }
    const_iterator Begin() const { return data.begin(); }
    const_iterator End() const { return data.end(); }
    void Swap(Row& o) { data.swap(o.data); }
};

class Vector {
    std::string name;
    INMOST_DATA_ENUM_TYPE first = 0, last = 0;
    std::vector<INMOST_DATA_REAL_TYPE> data;
public:
    explicit Vector(std::string nm = "", INMOST_DATA_ENUM_TYPE b = 0, INMOST_DATA_ENUM_TYPE e = 0) : name(std::move(nm)) { SetInterval(b, e); }
    void SetInterval(INMOST_DATA_ENUM_TYPE b, INMOST_DATA_ENUM_TYPE e) {   // keeps existing entries (Assemble accumulates)
        if (b == first && e == last) return;
        std::vector<INMOST_DATA_REAL_TYPE> nd(e - b, 0.0);
        for (INMOST_DATA_ENUM_TYPE i = std::max(b, first); i < std::min(e, last); ++i) nd[i - b] = data[i - first];
        data.swap(nd); first = b; last = e;
    }
    void GetInterval(INMOST_DATA_ENUM_TYPE& b, INMOST_DATA_ENUM_TYPE& e) const { b = first; e = last; }
    INMOST_DATA_ENUM_TYPE GetFirstIndex() const { return first; }
    INMOST_DATA_ENUM_TYPE GetLastIndex() const { return last; }
    INMOST_DATA_ENUM_TYPE Size() const { return (INMOST_DATA_ENUM_TYPE)data.size(); }
    INMOST_DATA_REAL_TYPE& operator[](INMOST_DATA_ENUM_TYPE i) { return data[i - first]; }
    INMOST_DATA_REAL_TYPE operator[](INMOST_DATA_ENUM_TYPE i) const { return data[i - first]; }
    void Clear() { data.clear(); first = last = 0; }
    std::vector<INMOST_DATA_REAL_TYPE>::iterator Begin() { return data.begin(); }
    std::vector<INMOST_DATA_REAL_TYPE>::iterator End() { return data.end(); }
};

class Matrix {
    std::string name;
    INMOST_DATA_ENUM_TYPE first = 0, last = 0;
    std::vector<Row> rows;
public:
    explicit Matrix(std::string nm = "", INMOST_DATA_ENUM_TYPE b = 0, INMOST_DATA_ENUM_TYPE e = 0) : name(std::move(nm)) { SetInterval(b, e); }
    void SetInterval(INMOST_DATA_ENUM_TYPE b, INMOST_DATA_ENUM_TYPE e) {
        if (b == first && e == last) return;
        std::vector<Row> nr(e - b);
        for (INMOST_DATA_ENUM_TYPE i = std::max(b, first); i < std::min(e, last); ++i) nr[i - b].Swap(rows[i - first]);
        rows.swap(nr); first = b; last = e;
    }
    void GetInterval(INMOST_DATA_ENUM_TYPE& b, INMOST_DATA_ENUM_TYPE& e) const { b = first; e = last; }
    INMOST_DATA_ENUM_TYPE GetFirstIndex() const { return first; }
    INMOST_DATA_ENUM_TYPE GetLastIndex() const { return last; }
    INMOST_DATA_ENUM_TYPE Size() const { return (INMOST_DATA_ENUM_TYPE)rows.size(); }
    Row& operator[](INMOST_DATA_ENUM_TYPE i) { return rows[i - first]; }
    const Row& operator[](INMOST_DATA_ENUM_TYPE i) const { return rows[i - first]; }
    void Clear() { rows.clear(); first = last = 0; }
};

class LockService {   // single-threaded use only in the mock
public:
    void SetInterval(INMOST_DATA_ENUM_TYPE, INMOST_DATA_ENUM_TYPE) {}
    bool Lock(INMOST_DATA_ENUM_TYPE) { return true; }
    bool UnLock(INMOST_DATA_ENUM_TYPE) { return true; }
    bool TestLock(INMOST_DATA_ENUM_TYPE) { return true; }
};

}  // namespace Sparse
}  // namespace INMOST
