// C entry to the host-side Dirichlet helpers of anifem_b200/dc_on_dof.hpp, mirroring oracle/ref_driver.cpp::ref_dirichlet_local,
// so that the CPU test suite can compare them with the reference's own helpers on the same random data.
#include <vector>

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
