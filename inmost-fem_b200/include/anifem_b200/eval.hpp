// anifem_b200/eval.hpp -- FE-function evaluation on tetrahedra: Ani::fem3DapplyL / fem3DapplyX (fem/operations/eval.h:13-120,
// eval.inl:103-165) on top of the C ABI (afb_fem3dapply_batched).  Standalone header: include it next to fem.hpp.
//
//   fem3DapplyL<Op>(XYZ, XYL, dofs, opU):  opU[k + dim*(n + q*r)] = sum_i Op(phi_i)(x_n)[k] * dofs[i + nfa*r] at the q points given by
//                                          barycentric coordinates XYL[4q] (the same on every tet of the batch);
//   fem3DapplyX<Op>(XYZ, X, dofs, opU):    the same on ONE tet at physical points X[3q]: the points are converted to barycentric
//                                          coordinates on the host (core.inl:119-137: lambda_{1..3} = PSI (x - P0), lambda_0 = 1 - sum)
//                                          and handed to the batched kernel.
#pragma once
#include "fem.hpp"

namespace Ani {

/// one tetrahedron (fem/geometry.h:143-160)
template <typename ScalarType = const double>
struct Tetra : public Tetras<ScalarType> {
    Tetra() : Tetras<ScalarType>{nullptr, nullptr, nullptr, nullptr, 1} {}
    Tetra(ScalarType* XY0, ScalarType* XY1, ScalarType* XY2, ScalarType* XY3) : Tetras<ScalarType>{XY0, XY1, XY2, XY3, 1} {}
    std::array<double, 3> centroid() const {
        std::array<double, 3> c;
        for (int k = 0; k < 3; ++k) c[k] = (this->XY0[k] + this->XY1[k] + this->XY2[k] + this->XY3[k]) / 4;
        return c;
    }
};

namespace b200 {
/// barycentric coordinates XYL[4n + 0..3] of the physical points X[3n + 0..2] in the tet (P0, P1, P2, P3)
inline void bary_coords_host(const double* P0, const double* P1, const double* P2, const double* P3, const double* X, std::size_t q, double* XYL) {
    // columns of M: P1 - P0, P2 - P0, P3 - P0;  lambda_{1..3} = M^{-1} (x - P0) by Cramer's rule
    const double m[3][3] = {{P1[0] - P0[0], P2[0] - P0[0], P3[0] - P0[0]}, {P1[1] - P0[1], P2[1] - P0[1], P3[1] - P0[1]}, {P1[2] - P0[2], P2[2] - P0[2], P3[2] - P0[2]}};
    const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2], c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    if (det == 0.0) throw std::runtime_error("degenerate tetrahedron");
    const double inv[3][3] = {{c00 / det, (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det, (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det},
                              {c01 / det, (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det, (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det},
                              {c02 / det, (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det, (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det}};
    for (std::size_t n = 0; n < q; ++n) {
        const double d[3] = {X[3 * n] - P0[0], X[3 * n + 1] - P0[1], X[3 * n + 2] - P0[2]};
        double s = 0;
        for (int i = 0; i < 3; ++i) {
            const double l = inv[i][0] * d[0] + inv[i][1] * d[1] + inv[i][2] * d[2];
            XYL[4 * n + i + 1] = l;
            s += l;
        }
        XYL[4 * n] = 1.0 - s;
    }
}
inline void check_apply_args(int f, std::size_t q, int nfa, int dim, const DenseMatrix<double>& dofs, const DenseMatrix<double>& opU) {
    if (dofs.nRow != static_cast<std::size_t>(nfa) || dofs.nCol != static_cast<std::size_t>(f))
        throw std::runtime_error("Expected dimension of dofs is " + std::to_string(nfa) + "x" + std::to_string(f));
    if (f <= 0) throw std::runtime_error("Coordinate arrays shouldn't be free");
    if (opU.size < static_cast<std::size_t>(dim) * q * f)
        throw std::runtime_error("opU.size = " + std::to_string(opU.size) + " but required at least " + std::to_string(dim * q * f));
}
}  // namespace b200

/// eval.h:13-23, 46-75
template <typename Op>
void fem3DapplyL(const Tetras<const double>& XYZ, ArrayView<double> XYL, const DenseMatrix<double>& dofs, DenseMatrix<double>& opU) {
    const int f = XYZ.fusion, nfa = Op::Nfa::value, dim = Op::Dim::value;
    const std::size_t q = XYL.size / 4;
    b200::check_apply_args(f, q, nfa, dim, dofs, opU);
    opU.nRow = static_cast<std::size_t>(dim) * q; opU.nCol = static_cast<std::size_t>(f);
    afb_ctx* ctx = b200::default_context();
    b200::check(ctx, afb_fem3dapply_batched(ctx, Op::op, Op::fem, Op::vec, static_cast<int>(q), XYL.data, f, XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, dofs.data,
                                            opU.data, AFB_HOST));
}
template <typename Op, typename ScalarType, typename IndexType>
void fem3DapplyL(const Tetras<const double>& XYZ, ArrayView<double> XYL, const DenseMatrix<double>& dofs, DenseMatrix<double>& opU,
                 PlainMemory<ScalarType, IndexType>) {
    fem3DapplyL<Op>(XYZ, XYL, dofs, opU);
}
template <typename Op>
void fem3DapplyL(const DenseMatrix<double>& XY0, const DenseMatrix<double>& XY1, const DenseMatrix<double>& XY2, const DenseMatrix<double>& XY3,
                 ArrayView<double> XYL, const DenseMatrix<double>& dofs, DenseMatrix<double>& opU) {
    fem3DapplyL<Op>(make_tetras(XY0.data, XY1.data, XY2.data, XY3.data, static_cast<int>(XY0.nCol)), XYL, dofs, opU);
}
/// eval.h:24-38: Op(u_h) at physical points of one tetrahedron
template <typename Op>
void fem3DapplyX(const Tetra<const double>& XYZ, const ArrayView<const double> X, const ArrayView<double>& dofs, ArrayView<double> opU) {
    const std::size_t q = X.size / 3;
    std::vector<double> xyl(4 * q);
    b200::bary_coords_host(XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, X.data, q, xyl.data());
    DenseMatrix<double> d(dofs.data, dofs.size, 1), o(opU.data, opU.size, 1);
    fem3DapplyL<Op>(static_cast<const Tetras<const double>&>(XYZ), ArrayView<double>(xyl.data(), xyl.size()), d, o);
}
template <typename Op, typename ScalarType, typename IndexType>
void fem3DapplyX(const Tetra<const double>& XYZ, const ArrayView<const double> X, const ArrayView<double>& dofs, ArrayView<double> opU,
                 PlainMemory<ScalarType, IndexType>) {
    fem3DapplyX<Op>(XYZ, X, dofs, opU);
}
/// no host scratch is needed (eval.inl:8-22)
template <typename Op, typename ScalarType = double, typename IndexType = int>
PlainMemory<ScalarType, IndexType> fem3DapplyL_memory_requirements(int /*pnt_per_tetra*/, int /*fusion*/ = 1) { return PlainMemory<ScalarType, IndexType>(); }
template <typename Op, typename ScalarType = double, typename IndexType = int>
PlainMemory<ScalarType, IndexType> fem3DapplyX_memory_requirements(int /*pnt_per_tetra*/, int /*fusion*/ = 1) { return PlainMemory<ScalarType, IndexType>(); }

}  // namespace Ani
