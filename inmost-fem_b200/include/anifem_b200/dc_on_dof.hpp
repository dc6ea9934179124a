// anifem_b200/dc_on_dof.hpp -- essential conditions on a LOCAL matrix / right-hand side: the host-side helpers a local assembler
// lambda calls (fem/operations/dc_on_dof.h:11-128, dc_on_dof.inl).  Plain dense arithmetic on the caller's element matrix; used by
// the MatFuncWrap path (func_wrap.hpp), where the lambda of the reference runs unchanged on the host.
//
// Scalar conditions: applyDirNonExists / applyDirMatrix / applyDir / applyDirResidual.
// Vector conditions V u = b for a vector variable u = (u_1 .. u_d) whose components use the same scalar space: dof_id[0..d) are the
// local dofs of ONE basis function in the d components, Vorth (d x d) is orthogonal and the constrained directions are its rows
// dc_orth[0..ndc) (rows 0..ndc-1 when dc_orth is NULL).  The helpers first express the d equations of that basis function in the
// basis of the rows of Vorth (test functions rotated), then eliminate the constrained directions:
//     columns:  A(:, dof) -= (A(:, dof) v_k) v_k^T,   rhs:  F -= (A(:, dof) v_k) b_k,   row of direction k := v_k^T on the dof columns.
#pragma once
#include <cstddef>
#include <stdexcept>

namespace Ani {

template <typename ScalarType> struct DenseMatrix;
template <typename T> struct ArrayView;
using uint = unsigned int;

template <typename Scalar>
inline void applyDirNonExists(DenseMatrix<Scalar>& A, int k) {
    for (std::size_t i = 0; i < A.nRow; ++i) A.data[i + A.nRow * k] = 0;
    for (std::size_t j = 0; j < A.nCol; ++j) A.data[k + A.nRow * j] = 0;
}
template <typename Scalar>
inline void applyDirNonExists(DenseMatrix<Scalar>& A, DenseMatrix<Scalar>& F, int k) {
    applyDirNonExists(A, k);
    F.data[k] = 0;
}
template <typename Scalar>
inline void applyDirMatrix(DenseMatrix<Scalar>& A, int k) {
    applyDirNonExists(A, k);
    A.data[k + A.nRow * k] = 1.0;
}
/// F(i) -= A(i,k) * bc for all i;  F(k) = bc;  row and column k of A zeroed;  A(k,k) = 1   (dc_on_dof.h:33-45)
template <typename Scalar>
inline void applyDir(DenseMatrix<Scalar>& A, DenseMatrix<Scalar>& F, int k, Scalar bc) {
    for (std::size_t i = 0; i < A.nRow; ++i) F.data[i] -= A.data[i + A.nRow * k] * bc;
    F.data[k] = bc;
    applyDirMatrix(A, k);
}
template <typename Scalar>
inline void applyDirResidual(DenseMatrix<Scalar>& F, int k) { F.data[k] = 0.0; }

namespace dc_detail {
struct DirIds {   // constrained direction k -> row of Vorth
    const uint* map;
    uint operator[](uint k) const { return map ? map[k] : k; }
};
// rows dof_id[0..d) of M (nr x nc, col-major) <- Vorth * those rows
template <typename Scalar, typename RandomIt>
inline void rotate_rows(Scalar* M, std::size_t nr, std::size_t nc, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, Scalar* work) {
    const uint d = static_cast<uint>(Vorth.nCol);
    for (std::size_t j = 0; j < nc; ++j) {
        for (uint p = 0; p < d; ++p) work[p] = M[dof_id[p] + nr * j];
        for (uint m = 0; m < d; ++m) {
            Scalar s = 0;
            for (uint p = 0; p < d; ++p) s += work[p] * Vorth.data[m + Vorth.nRow * p];
            M[dof_id[m] + nr * j] = s;
        }
    }
}
template <typename Scalar, typename RandomIt>
inline void eliminate(DenseMatrix<Scalar>& A, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, Scalar* work, uint ndc, DirIds id) {
    const uint d = static_cast<uint>(Vorth.nCol);
    auto V = [&](uint r, uint c) { return Vorth.data[r + Vorth.nRow * c]; };
    for (uint k = 0; k < ndc; ++k) {
        const uint vk = id[k];
        for (std::size_t j = 0; j < A.nRow; ++j) {
            Scalar s = 0;
            for (uint p = 0; p < d; ++p) s += A.data[j + A.nRow * dof_id[p]] * V(vk, p);
            work[j] = s;
        }
        for (uint l = 0; l < d; ++l)
            for (std::size_t j = 0; j < A.nRow; ++j) A.data[j + A.nRow * dof_id[l]] -= work[j] * V(vk, l);
        const std::size_t row = dof_id[vk];
        for (std::size_t j = 0; j < A.nCol; ++j) A.data[row + A.nRow * j] = 0;
        for (uint m = 0; m < d; ++m) A.data[row + A.nRow * dof_id[m]] = V(vk, m);
    }
}
template <typename Scalar>
inline void need(const ArrayView<Scalar>& mem, std::size_t n) {
    if (mem.size < n) throw std::runtime_error("Not enough of memory");
}
}  // namespace dc_detail

/// matrix part of the vector condition; mem holds at least d * max(A.nCol, A.nRow) scalars (dc_on_dof.h:53-70)
template <typename Scalar, typename RandomIt>
inline void applyVectorDirMatrix(DenseMatrix<Scalar>& A, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, ArrayView<Scalar> mem, const uint ndc,
                                 const uint* dc_orth = nullptr) {
    dc_detail::need(mem, Vorth.nCol * (A.nCol > A.nRow ? A.nCol : A.nRow));
    dc_detail::rotate_rows(A.data, A.nRow, A.nCol, dof_id, Vorth, mem.data);
    dc_detail::eliminate(A, dof_id, Vorth, mem.data, ndc, dc_detail::DirIds{dc_orth});
}
/// B: a column of the element matrix that belongs to another variable (rows of all dofs): its entries at dof_id are rotated and
/// the constrained directions zeroed (dc_on_dof.h:71-78)
template <typename Scalar, typename RandomIt>
inline void applyVectorDirMatrixExtCol(ArrayView<Scalar>& B, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, ArrayView<Scalar> mem, const uint ndc,
                                       const uint* dc_orth = nullptr) {
    const uint d = static_cast<uint>(Vorth.nCol);
    dc_detail::need(mem, d);
    dc_detail::DirIds id{dc_orth};
    for (uint p = 0; p < d; ++p) mem.data[p] = B.data[dof_id[p]];
    for (uint m = 0; m < d; ++m) {
        Scalar s = 0;
        for (uint p = 0; p < d; ++p) s += mem.data[p] * Vorth.data[m + Vorth.nRow * p];
        B.data[dof_id[m]] = s;
    }
    for (uint k = 0; k < ndc; ++k) B.data[dof_id[id[k]]] = 0;
}
/// C: a row of the element matrix that belongs to another variable: the constrained directions are projected out of its entries
/// at dof_id (dc_on_dof.h:79-84)
template <typename Scalar, typename RandomIt>
inline void applyVectorDirMatrixExtRow(ArrayView<Scalar>& C, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, const uint ndc, const uint* dc_orth = nullptr) {
    const uint d = static_cast<uint>(Vorth.nCol);
    dc_detail::DirIds id{dc_orth};
    for (uint k = 0; k < ndc; ++k) {
        const uint vk = id[k];
        Scalar c = 0;
        for (uint p = 0; p < d; ++p) c += C.data[dof_id[p]] * Vorth.data[vk + Vorth.nRow * p];
        for (uint l = 0; l < d; ++l) C.data[dof_id[l]] -= c * Vorth.data[vk + Vorth.nRow * l];
    }
}
/// matrix and right-hand side; bc[k] is the prescribed value of direction k (dc_on_dof.h:85-95)
template <typename Scalar, typename RandomIt>
inline void applyVectorDir(DenseMatrix<Scalar>& A, DenseMatrix<Scalar>& F, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, const ArrayView<Scalar>& bc,
                           ArrayView<Scalar> mem, const uint ndc, const uint* dc_orth = nullptr) {
    const uint d = static_cast<uint>(Vorth.nCol);
    dc_detail::need(mem, d * (A.nCol > A.nRow ? A.nCol : A.nRow));
    dc_detail::DirIds id{dc_orth};
    dc_detail::rotate_rows(A.data, A.nRow, A.nCol, dof_id, Vorth, mem.data);
    dc_detail::rotate_rows(F.data, F.nRow, 1, dof_id, Vorth, mem.data);
    for (uint k = 0; k < ndc; ++k) {
        const uint vk = id[k];
        for (uint m = 0; m < d; ++m) {
            const Scalar w = Vorth.data[vk + Vorth.nRow * m] * bc.data[k];
            for (std::size_t j = 0; j < A.nRow; ++j) F.data[j] -= A.data[j + A.nRow * dof_id[m]] * w;
        }
        F.data[dof_id[vk]] = bc.data[k];
    }
    dc_detail::eliminate(A, dof_id, Vorth, mem.data, ndc, id);
}
/// residual form: rotated right-hand side with zeros in the constrained directions (dc_on_dof.h:96-103)
template <typename Scalar, typename RandomIt>
inline void applyVectorDirResidual(DenseMatrix<Scalar>& F, RandomIt dof_id, const DenseMatrix<Scalar>& Vorth, ArrayView<Scalar> mem, const uint ndc,
                                   const uint* dc_orth = nullptr) {
    dc_detail::need(mem, Vorth.nCol);
    dc_detail::DirIds id{dc_orth};
    dc_detail::rotate_rows(F.data, F.nRow, 1, dof_id, Vorth, mem.data);
    for (uint k = 0; k < ndc; ++k) F.data[dof_id[id[k]]] = 0;
}

}  // namespace Ani
