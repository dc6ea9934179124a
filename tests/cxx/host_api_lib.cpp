// C entry to the host-side Dirichlet helpers of anifem_b200/dc_on_dof.hpp, mirroring oracle/ref_driver.cpp::ref_dirichlet_local,
// so that the CPU test suite can compare them with the reference's own helpers on the same random data.
#include <vector>

#include "anifem_b200/dofmap.hpp"
#include "anifem_b200/fem.hpp"

extern "C" int mine_dirichlet_local(int what, int n, double* A, double* F, int d, const unsigned* dof_id, const double* Vorth, int ndc, const double* bc,
                                    const unsigned* dc_orth) {
    using namespace Ani;
    try {
        DenseMatrix<double> Am(A, n, n), Fm(F, n, 1), V(const_cast<double*>(Vorth), d, d);
        std::vector<double> work(static_cast<std::size_t>(d) * n + 16);
        ArrayView<double> mem(work.data(), work.size()), bcv(const_cast<double*>(bc), ndc);
        switch (what) {
            case 0: applyVectorDir<double>(Am, Fm, dof_id, V, bcv, mem, static_cast<uint>(ndc), dc_orth); break;
            case 1: applyVectorDirMatrix<double>(Am, dof_id, V, mem, static_cast<uint>(ndc), dc_orth); break;
            case 2: applyVectorDirResidual<double>(Fm, dof_id, V, mem, static_cast<uint>(ndc), dc_orth); break;
            case 3: applyDir<double>(Am, Fm, static_cast<int>(dof_id[0]), bc[0]); break;
            case 4: { ArrayView<double> col(A, n); applyVectorDirMatrixExtCol<double>(col, dof_id, V, mem, static_cast<uint>(ndc), dc_orth); break; }
            case 5: { ArrayView<double> row(A, n); applyVectorDirMatrixExtRow<double>(row, dof_id, V, static_cast<uint>(ndc), dc_orth); break; }
            default: return -7;
        }
        return 0;
    } catch (std::exception&) { return -1; }
}

// ---- local dof maps (anifem_b200/dofmap.hpp): the same flat description as oracle/ref_driver.cpp::ref_dofmap_table
static Ani::DofT::DofMap mine_build_dofmap(const int*& p) {
    using namespace Ani::DofT;
    const int kind = *p++;
    if (kind == 1) { std::array<uint, NGEOM_TYPES> n; for (int t = 0; t < NGEOM_TYPES; ++t) n[t] = static_cast<uint>(*p++); return DofMap(std::make_shared<UniteDofMap>(n)); }
    if (kind == 2) { const int dim = *p++; DofMap b = mine_build_dofmap(p); return DofMap(std::make_shared<VectorDofMap>(dim, b.base())); }
    if (kind == 3) { const int k = *p++; std::vector<DofMap> v; for (int i = 0; i < k; ++i) v.push_back(mine_build_dofmap(p)); return merge(v); }
    if (kind == 4) { const int k = *p++; std::vector<DofMap> v; for (int i = 0; i < k; ++i) v.push_back(mine_build_dofmap(p)); return merge_with_simplifications(v); }
    if (kind == 5) { const int k = *p++; DofMap b = mine_build_dofmap(p); return b ^ static_cast<uint>(k); }
    throw std::runtime_error("bad dof map description");
}
extern "C" int mine_dofmap_table(const int* spec, int* out, int cap, const int* sel, int* by_sp) {
    using namespace Ani::DofT;
    try {
        const int* p = spec;
        DofMap m = mine_build_dofmap(p);
        const int n = static_cast<int>(m.NumDofOnTet());
        if (n > cap) return -2;
        for (int g = 0; g < n; ++g) {
            LocalOrder lo = m.LocalOrderOnTet(TetOrder(static_cast<uint>(g)));
            out[3 * g] = lo.etype; out[3 * g + 1] = lo.nelem; out[3 * g + 2] = static_cast<int>(lo.leid);
            if (static_cast<int>(m.TetDofID(lo.getGeomOrder())) != g) return -3;
        }
        if (sel && by_sp) {
            TetGeomSparsity sp;
            for (int d = 0; d < 4; ++d) for (int i = 0; i < 6; ++i) if ((sel[d] >> i) & 1) sp.set(static_cast<uchar>(d), i, false);
            std::vector<int> ids;
            for (auto it = m.beginBySparsity(sp, false); it != m.endBySparsity(); ++it) ids.push_back(static_cast<int>((*it).gid));
            // also the entity-ordered walk must visit the same dofs
            std::vector<int> ids2;
            for (auto it = m.beginBySparsity(sp, true); it != m.endBySparsity(); ++it) ids2.push_back(static_cast<int>((*it).gid));
            std::sort(ids2.begin(), ids2.end());
            if (!std::is_sorted(ids.begin(), ids.end()) || ids != ids2) return -4;
            by_sp[0] = static_cast<int>(ids.size());
            for (std::size_t k = 0; k < ids.size(); ++k) by_sp[1 + k] = ids[k];
        }
        return n;
    } catch (std::exception&) { return -1; }
}
extern "C" int mine_dofmap_equal(const int* spec_a, const int* spec_b) {
    try {
        const int *pa = spec_a, *pb = spec_b;
        return mine_build_dofmap(pa) == mine_build_dofmap(pb) ? 1 : 0;
    } catch (std::exception&) { return -1; }
}
extern "C" int mine_sparsity_ops(int dim, int i, int closure, int udim, int ui, int uclosure, int* bits) {
    using namespace Ani::DofT;
    try {
        TetGeomSparsity sp;
        sp.set(static_cast<uchar>(dim), i, closure != 0);
        if (udim >= 0) sp.unset(static_cast<uchar>(udim), ui, uclosure != 0);
        for (int d = 0; d < 4; ++d) { auto ids = sp.getElemsIds(static_cast<uchar>(d)); bits[d] = 0; for (int k = 0; k < ids.second; ++k) bits[d] |= 1 << ids.first[k]; }
        return 0;
    } catch (std::exception&) { return -1; }
}
