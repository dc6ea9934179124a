// anifem_b200/inmost_adapter.hpp -- glue between an INMOST mesh / INMOST sparse containers and the array interface of this
// library.  The reference's Assembler reads the mesh through INMOST handles (inmost_interface/assembler.inl:
// 339-395, ordering.inl:8-26) and writes into INMOST::Sparse::Matrix / Vector (assembler.inl:397-481); here the mesh is copied once
// into SoA arrays (afb_mesh_set) and the assembled CSR rows are written back row by row.
//
// Include AFTER <inmost.h>: the header uses only the public INMOST API (Mesh::BeginNode/BeginCell, Element::getNodes,
// Node::Coords, Storage::GlobalID / LocalID, Element::GetStatus, Sparse::Matrix / Row / Vector).  Tested in this repository against
// oracle/mock_inmost/inmost.h (tests/cxx/test_inmost_adapter.cpp); INMOST itself is not vendored (cmake/Downloadinmost.cmake:4).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "assembler.hpp"

namespace Ani {
namespace b200 {

/// tetrahedral mesh of an INMOST::Mesh as arrays; node k of the arrays is the k-th node met by the node iteration
struct InmostMeshArrays {
    std::vector<double> x, y, z;
    std::vector<int32_t> v[4];                     ///< node indices of every tetrahedron (INMOST's getNodes order; orientation is fixed by the library)
    std::vector<INMOST::HandleType> node, cell;    ///< handles of the nodes / cells behind the array entries
    std::vector<long> node_gid;                    ///< GlobalID of the nodes when the mesh has them, else the array index
};

/// Copies nodes and tetrahedral cells (ghost cells as well unless only_owned: the reference assembles over all local cells and
/// skips ghost ROWS, assembler.inl:162-183).  Throws like the reference on a non-tetrahedral cell (assembler.inl:309-312).
inline InmostMeshArrays mesh_to_arrays(INMOST::Mesh* m, bool only_owned_cells = false) {
    if (!m) throw std::runtime_error("Mesh was not specified");
    InmostMeshArrays a;
    const long nn = static_cast<long>(m->NumberOf(INMOST::NODE));
    a.x.reserve(nn); a.y.reserve(nn); a.z.reserve(nn);
    std::vector<int32_t> index_of(static_cast<std::size_t>(m->NodeLastLocalID()), -1);   // LocalID -> array index
    const bool have_gid = m->HaveGlobalID(INMOST::NODE);
    for (auto it = m->BeginNode(); it != m->EndNode(); ++it) {
        INMOST::Node n = it->getAsNode();
        auto c = n.Coords();
        index_of[static_cast<std::size_t>(n.LocalID())] = static_cast<int32_t>(a.x.size());
        a.node_gid.push_back(have_gid ? static_cast<long>(n.GlobalID()) : static_cast<long>(a.x.size()));
        a.x.push_back(c[0]); a.y.push_back(c[1]); a.z.push_back(c[2]);
        a.node.push_back(n.GetHandle());
    }
    for (auto it = m->BeginCell(); it != m->EndCell(); ++it) {
        INMOST::Cell c = it->getAsCell();
        if (only_owned_cells && c.GetStatus() == INMOST::Element::Ghost) continue;
        auto nodes = c.getNodes();
        if (nodes.size() != 4) throw std::runtime_error("The assembler supports only tetrahedral cells");
        for (int k = 0; k < 4; ++k) a.v[k].push_back(index_of[static_cast<std::size_t>(nodes[k].LocalID())]);
        a.cell.push_back(c.GetHandle());
    }
    return a;
}

/// Assembler::SetMesh(INMOST::Mesh*) of the reference (assembler.h:318): copies the mesh to the device, orientation included
inline InmostMeshArrays SetMesh(Assembler& discr, INMOST::Mesh* m) {
    InmostMeshArrays a = mesh_to_arrays(m);
    discr.SetMesh(static_cast<int64_t>(a.x.size()), a.x.data(), a.y.data(), a.z.data(), static_cast<int64_t>(a.v[0].size()), a.v[0].data(), a.v[1].data(),
                  a.v[2].data(), a.v[3].data());
    return a;
}

/// assembled rows -> INMOST::Sparse::Matrix: rows [row_begin, row_end), entries in ascending column order (what the reference
/// produces with use_ordered_insert, assembler.inl:428-438).  add = true adds to entries that already exist.
inline void csr_to_inmost(const CsrMatrix& A, INMOST::Sparse::Matrix& M, bool add = false) {
    M.SetInterval(static_cast<INMOST_DATA_ENUM_TYPE>(A.row_begin), static_cast<INMOST_DATA_ENUM_TYPE>(A.row_end));
    for (int64_t r = A.row_begin; r < A.row_end; ++r) {
        INMOST::Sparse::Row& row = M[static_cast<INMOST_DATA_ENUM_TYPE>(r)];
        const int64_t p0 = A.rowptr[static_cast<std::size_t>(r - A.row_begin)], p1 = A.rowptr[static_cast<std::size_t>(r - A.row_begin) + 1];
        if (add && !row.Empty()) {
            for (int64_t p = p0; p < p1; ++p) row[static_cast<INMOST_DATA_ENUM_TYPE>(A.colind[static_cast<std::size_t>(p)])] += A.val[static_cast<std::size_t>(p)];
            continue;
        }
        row.Resize(static_cast<INMOST_DATA_ENUM_TYPE>(p1 - p0));
        for (int64_t p = p0; p < p1; ++p) {
            row.GetIndex(static_cast<INMOST_DATA_ENUM_TYPE>(p - p0)) = static_cast<INMOST_DATA_ENUM_TYPE>(A.colind[static_cast<std::size_t>(p)]);
            row.GetValue(static_cast<INMOST_DATA_ENUM_TYPE>(p - p0)) = A.val[static_cast<std::size_t>(p)];
        }
    }
}
inline void rhs_to_inmost(const std::vector<double>& rhs, int64_t row_begin, INMOST::Sparse::Vector& v, bool add = false) {
    v.SetInterval(static_cast<INMOST_DATA_ENUM_TYPE>(row_begin), static_cast<INMOST_DATA_ENUM_TYPE>(row_begin + static_cast<int64_t>(rhs.size())));
    for (std::size_t k = 0; k < rhs.size(); ++k) {
        const INMOST_DATA_ENUM_TYPE i = static_cast<INMOST_DATA_ENUM_TYPE>(row_begin + static_cast<int64_t>(k));
        if (add) v[i] += rhs[k]; else v[i] = rhs[k];
    }
}
/// INMOST::Sparse::Matrix -> CSR with ascending columns (e.g. to compare a matrix the reference assembled with ours)
inline CsrMatrix inmost_to_csr(const INMOST::Sparse::Matrix& M) {
    CsrMatrix A;
    A.row_begin = static_cast<int64_t>(M.GetFirstIndex()); A.row_end = static_cast<int64_t>(M.GetLastIndex());
    A.rowptr.assign(static_cast<std::size_t>(A.row_end - A.row_begin) + 1, 0);
    std::vector<std::pair<int32_t, double>> tmp;
    for (int64_t r = A.row_begin; r < A.row_end; ++r) {
        const INMOST::Sparse::Row& row = M[static_cast<INMOST_DATA_ENUM_TYPE>(r)];
        tmp.clear();
        for (INMOST_DATA_ENUM_TYPE k = 0; k < row.Size(); ++k) tmp.emplace_back(static_cast<int32_t>(row.GetIndex(k)), row.GetValue(k));
        std::sort(tmp.begin(), tmp.end(), [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) { return a.first < b.first; });
        for (auto& e : tmp) { A.colind.push_back(e.first); A.val.push_back(e.second); }
        A.rowptr[static_cast<std::size_t>(r - A.row_begin) + 1] = static_cast<int64_t>(A.colind.size());
    }
    return A;
}

}  // namespace b200
}  // namespace Ani
