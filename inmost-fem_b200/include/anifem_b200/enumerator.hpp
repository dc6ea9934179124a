// anifem_b200/enumerator.hpp -- host-side twin of the reference's GlobEnumeration family for ONE rank
// (inmost_interface/global_enumerator.h:393-428, global_enumerator.cpp:671-777,818-890).
//
// The reference numbers the dofs on the host at set-up time (integer tags on mesh entities); so does this header: it turns a
// tetrahedral connectivity + a list of variables into the elem -> global dof table that afb_dofmap_set (include/anifem_b200.h)
// consumes.  NATURAL is also available on the device (afb_dofmap_natural, used by the benchmarks); the other types go through
// this table.  Pure host code, no CUDA, no dependency on the library.
//
//   ASSEMBLING_TYPE           index of dof k (component c) of variable v on entity g of geometric type et
//   NATURAL      (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID)   off[v][c][et] + g * ns + k
//   DIMUNION     (VAR, ELEM_TYPE, ELEM_ID, DOF_ID, DIM)   off[v][et] + (g * ns + k) * ndim + c
//   BYELEMTYPE   (ELEM_TYPE, VAR, DIM, ELEM_ID, DOF_ID)   off[et][v][c] + g * ns + k
//   ETDIMBLOCKS  (ELEM_TYPE, VAR, ELEM_ID, DOF_ID, DIM)   off[et][v] + (g * ns + k) * ndim + c
//   ANITYPE      InitElemIndex[et] + g + iodf * NumElem[et]              (global_enumerator.cpp:823-825)
//   MINIBLOCKS   InitElemIndex[et] + g * nd[et] + iodf                   (the layout the inverse map decodes, :866-871)
// with ns = dofs of the scalar base space per entity, ndim = components, off[..] = number of dofs whose tuple precedes the group
// in the lexicographic order of the arrangement, iodf = position among all dofs on the entity = GetElemDofId (variables in order,
// inside a vector variable component-major: c * ns + k, global_enumerator.h:76, .cpp:262-274; pinned against the reference's
// own SimpleEnumerator compiled on oracle/mock_inmost), nd[et] = dofs per entity over all variables.
// Entity ids (our stand-in for INMOST GlobalIDs, as in afb_dofmap_natural): nodes by id, edges and faces in lexicographic order of
// their sorted node tuples, cells by id.  Local order on the tet: variable, component, 4 vertices, 6 edges (01,02,03,12,13,23;
// the two dofs of a P3 edge ordered by the node ids, tetdofmap.inl:98-104), 4 faces (012,123,230,301), cell (fem_space.h:27-69).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace Ani {

enum ASSEMBLING_TYPE { ANITYPE = 0, MINIBLOCKS = 1, NATURAL = 2, DIMUNION = 3, BYELEMTYPE = 4, ETDIMBLOCKS = 5 };

struct EnumVar { int fem; int vec; };   // fem: AFB_FEM_P0..P3 values (1..4), vec: 1 or 3

struct DofEnumeration {
    int64_t nrows = 0;                  // MatrSize = BegInd..EndInd of the single rank
    int nloc = 0;                       // dofs per tetrahedron
    std::vector<int64_t> elem2dof;      // [i + nloc*e], 0-based ids (afb_dofmap_set takes id + 1)
};

namespace enum_detail {
// dofs per (node, edge, face, cell) of the scalar spaces (fem/spaces/poly_{0,1,2,3}.h Dof<>::Map())
inline std::array<int, 4> ndof(int fem) {
    switch (fem) {
        case 1: return {0, 0, 0, 1};
        case 2: return {1, 0, 0, 0};
        case 3: return {1, 1, 0, 0};
        case 4: return {1, 2, 1, 0};
    }
    throw std::runtime_error("enumerate_dofs: unsupported finite element space");
}
}  // namespace enum_detail

inline DofEnumeration enumerate_dofs(ASSEMBLING_TYPE type, int64_t nnode, int64_t ntet, const int32_t* v0, const int32_t* v1, const int32_t* v2,
                                     const int32_t* v3, const std::vector<EnumVar>& vars) {
    using enum_detail::ndof;
    if (vars.empty() || ntet <= 0 || nnode <= 0) throw std::runtime_error("enumerate_dofs: empty problem");
    const int32_t* vv[4] = {v0, v1, v2, v3};
    static const int LE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int LF[4][3] = {{0, 1, 2}, {1, 2, 3}, {2, 3, 0}, {3, 0, 1}};
    bool need_e = false, need_f = false;
    for (const EnumVar& v : vars) {
        if (v.vec != 1 && v.vec != 3) throw std::runtime_error("enumerate_dofs: vec must be 1 or 3");
        need_e = need_e || ndof(v.fem)[1] > 0;
        need_f = need_f || ndof(v.fem)[2] > 0;
    }
    // ---- entity ids
    std::vector<int64_t> ekeys;                         // sorted unique min*nnode + max
    std::vector<int64_t> tet_edge;                      // [6*e + le]
    if (need_e) {
        ekeys.resize((size_t)6 * ntet);
        for (int64_t e = 0; e < ntet; ++e)
            for (int le = 0; le < 6; ++le) {
                const int64_t a = vv[LE[le][0]][e], b = vv[LE[le][1]][e];
                ekeys[6 * e + le] = std::min(a, b) * nnode + std::max(a, b);
            }
        std::vector<int64_t> inst = ekeys;
        std::sort(ekeys.begin(), ekeys.end());
        ekeys.erase(std::unique(ekeys.begin(), ekeys.end()), ekeys.end());
        tet_edge.resize(inst.size());
        for (size_t t = 0; t < inst.size(); ++t) tet_edge[t] = std::lower_bound(ekeys.begin(), ekeys.end(), inst[t]) - ekeys.begin();
    }
    std::vector<std::array<int32_t, 3>> fkeys;
    std::vector<int64_t> tet_face;
    if (need_f) {
        fkeys.resize((size_t)4 * ntet);
        for (int64_t e = 0; e < ntet; ++e)
            for (int lf = 0; lf < 4; ++lf) {
                std::array<int32_t, 3> t = {vv[LF[lf][0]][e], vv[LF[lf][1]][e], vv[LF[lf][2]][e]};
                std::sort(t.begin(), t.end());
                fkeys[4 * e + lf] = t;
            }
        std::vector<std::array<int32_t, 3>> inst = fkeys;
        std::sort(fkeys.begin(), fkeys.end());
        fkeys.erase(std::unique(fkeys.begin(), fkeys.end()), fkeys.end());
        tet_face.resize(inst.size());
        for (size_t t = 0; t < inst.size(); ++t) tet_face[t] = std::lower_bound(fkeys.begin(), fkeys.end(), inst[t]) - fkeys.begin();
    }
    const int64_t nent[4] = {nnode, (int64_t)ekeys.size(), (int64_t)fkeys.size(), ntet};
    const int nv = (int)vars.size();
    // ---- group offsets
    // off[(v*3 + c)*4 + d] for the arrangements that keep the component outside (NATURAL, BYELEMTYPE); off2[v*4 + d] for the
    // ones with the component innermost (DIMUNION, ETDIMBLOCKS); init[d], shift[v*4 + d], nd[d] for the simple enumerators
    std::vector<int64_t> off((size_t)nv * 3 * 4, 0), off2((size_t)nv * 4, 0), shift((size_t)nv * 4, 0);
    int64_t nd[4] = {0, 0, 0, 0}, init[5] = {0, 0, 0, 0, 0};
    for (int d = 0; d < 4; ++d)
        for (int v = 0; v < nv; ++v) { shift[(size_t)v * 4 + d] = nd[d]; nd[d] += (int64_t)ndof(vars[v].fem)[d] * vars[v].vec; }
    for (int d = 0; d < 4; ++d) init[d + 1] = init[d] + nd[d] * nent[d];
    int64_t run = 0;
    if (type == NATURAL) {
        for (int v = 0; v < nv; ++v) for (int c = 0; c < vars[v].vec; ++c) for (int d = 0; d < 4; ++d) {
            off[((size_t)v * 3 + c) * 4 + d] = run; run += (int64_t)ndof(vars[v].fem)[d] * nent[d];
        }
    } else if (type == BYELEMTYPE) {
        for (int d = 0; d < 4; ++d) for (int v = 0; v < nv; ++v) for (int c = 0; c < vars[v].vec; ++c) {
            off[((size_t)v * 3 + c) * 4 + d] = run; run += (int64_t)ndof(vars[v].fem)[d] * nent[d];
        }
    } else if (type == DIMUNION) {
        for (int v = 0; v < nv; ++v) for (int d = 0; d < 4; ++d) { off2[(size_t)v * 4 + d] = run; run += (int64_t)ndof(vars[v].fem)[d] * vars[v].vec * nent[d]; }
    } else if (type == ETDIMBLOCKS) {
        for (int d = 0; d < 4; ++d) for (int v = 0; v < nv; ++v) { off2[(size_t)v * 4 + d] = run; run += (int64_t)ndof(vars[v].fem)[d] * vars[v].vec * nent[d]; }
    } else if (type != ANITYPE && type != MINIBLOCKS) throw std::runtime_error("Faced unknown ASSEMBLING_TYPE");
    auto index = [&](int v, int c, int d, int64_t g, int k) -> int64_t {
        const int ns = ndof(vars[v].fem)[d], ndim = vars[v].vec;
        switch (type) {
            case NATURAL: case BYELEMTYPE: return off[((size_t)v * 3 + c) * 4 + d] + g * ns + k;
            case DIMUNION: case ETDIMBLOCKS: return off2[(size_t)v * 4 + d] + (g * ns + k) * ndim + c;
            case ANITYPE: return init[d] + g + (shift[(size_t)v * 4 + d] + (int64_t)c * ns + k) * nent[d];
            default: return init[d] + g * nd[d] + shift[(size_t)v * 4 + d] + (int64_t)c * ns + k;   // MINIBLOCKS
        }
    };
    // ---- elem -> dof
    DofEnumeration out;
    out.nrows = init[4];
    for (const EnumVar& v : vars) { const auto n = ndof(v.fem); out.nloc += v.vec * (4 * n[0] + 6 * n[1] + 4 * n[2] + n[3]); }
    out.elem2dof.resize((size_t)out.nloc * ntet);
    for (int64_t e = 0; e < ntet; ++e) {
        int64_t* row = out.elem2dof.data() + (size_t)out.nloc * e;
        int i = 0;
        for (int v = 0; v < nv; ++v) {
            const auto n = ndof(vars[v].fem);
            for (int c = 0; c < vars[v].vec; ++c) {
                if (n[0]) for (int l = 0; l < 4; ++l) row[i++] = index(v, c, 0, vv[l][e], 0);
                if (n[1]) for (int le = 0; le < 6; ++le) {
                    const int64_t g = tet_edge[6 * e + le];
                    if (n[1] == 2) {
                        const int flip = vv[LE[le][0]][e] > vv[LE[le][1]][e] ? 1 : 0;
                        row[i++] = index(v, c, 1, g, flip);
                        row[i++] = index(v, c, 1, g, 1 - flip);
                    } else row[i++] = index(v, c, 1, g, 0);
                }
                if (n[2]) for (int lf = 0; lf < 4; ++lf) row[i++] = index(v, c, 2, tet_face[4 * e + lf], 0);
                if (n[3]) row[i++] = index(v, c, 3, e, 0);
            }
        }
    }
    return out;
}

}  // namespace Ani
