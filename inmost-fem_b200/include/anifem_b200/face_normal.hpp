// anifem_b200/face_normal.hpp -- fem3DfaceN: surface integrals weighted with the face normal (fem/operations/int_face.h:32-47,
// 97-133; dyn_ops.h:36-41), included by fem.hpp.
//
//     A = int_f ((D OpA(u)) . N) . OpB(v) ds,    N = unit normal of face f pointing away from the opposite vertex.
// The callback fills a col-major (3 Dim(OpB) x Dim(OpA)) tensor per point, the normal component being the fastest row index
// (fem/operations/core.h:176-195); a scalar / identity tensor needs 3 Dim(OpB) == Dim(OpA) (e.g. OpA = GRAD, OpB = IDEN: s du/dn v).
// Since N is constant on a face, contracting the tensor with it on the host leaves an ordinary fem3Dface call with the general
// tensor D_N(k, j) = sum_m N_m D(m + 3k, j): that call is the validated batched face kernel (afb_fem3dface_batched).  The
// contraction takes the face evaluator as a parameter; tests/cxx/test_composite.cpp plugs in the reference build's fem3Dface and
// compares with the reference's own fem3DfaceN.
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace Ani {
namespace b200 {

/// outward unit normal of face {face, face+1, face+2 mod 4} of tet r (geometry.h:77-93: the orientation is fixed by the opposite
/// vertex); returns the face area
inline double face_unit_normal(const Tetras<const double>& XYZ, int r, int face_num, double n[3]) {
    const double* P[4] = {XYZ.XY0 + 3 * r, XYZ.XY1 + 3 * r, XYZ.XY2 + 3 * r, XYZ.XY3 + 3 * r};
    const double *p0 = P[face_num % 4], *p1 = P[(face_num + 1) % 4], *p2 = P[(face_num + 2) % 4], *p3 = P[(face_num + 3) % 4];
    const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    n[0] = a[1] * b[2] - a[2] * b[1];
    n[1] = a[2] * b[0] - a[0] * b[2];
    n[2] = a[0] * b[1] - a[1] * b[0];
    const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double side = n[0] * (p3[0] - p0[0]) + n[1] * (p3[1] - p0[1]) + n[2] * (p3[2] - p0[2]);
    const double s = side > 0 ? -1.0 / len : 1.0 / len;
    for (int k = 0; k < 3; ++k) n[k] *= s;
    return 0.5 * len;
}

/// points of the triangle rule lifted to the face (int_face.inl:69-76, core.inl:249-269), 3 x q x f
inline std::vector<double> face_quad_points_host(const Tetras<const double>& XYZ, int face_num, int order, int& q) {
    q = afb_tri_quadrature(order, nullptr, nullptr, 0);
    if (q < 0) throw std::runtime_error("Numerical triangle integration formula implemented only for 0 <= order <= 20");
    std::vector<double> p(static_cast<std::size_t>(3) * q), w(q), X(static_cast<std::size_t>(3) * q * XYZ.fusion);
    afb_tri_quadrature(order, p.data(), w.data(), q);
    const double* P[4] = {XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3};
    for (int r = 0; r < XYZ.fusion; ++r)
        for (int n = 0; n < q; ++n)
            for (int k = 0; k < 3; ++k) {
                double s = 0;
                for (int m = 0; m < 3; ++m) s += p[3 * n + m] * P[(face_num + m) % 4][k + 3 * r];
                X[k + 3 * (n + static_cast<std::size_t>(q) * r)] = s;
            }
    return X;
}

/// eval(D_N, nrec): D_N general tensors, col-major (dimB x dimA) per record, nrec = 1 per TET for a constant tensor (the normal
/// differs from tet to tet) or q per tet -> fills A through a fem3Dface evaluation
template <typename FuncTraits, typename Functor, typename FaceEval>
void fem3DfaceN_contract(int dimA, int dimB, const Tetras<const double>& XYZ, int face_num, const Functor& Dfnc, int order, void* user_data, FaceEval&& eval) {
    const int f = XYZ.fusion;
    if (f <= 0) return;
    if (face_num < 0 || face_num > 3) throw std::runtime_error("Wrong face index");
    constexpr bool is_constant = FuncTraits::IsConstant::value && FuncTraits::AggregateType::value == OnePointTensor;
    const int jd = 3 * dimB;
    const TensorDims dims{static_cast<std::size_t>(jd), static_cast<std::size_t>(dimA)};
    const std::size_t dl = static_cast<std::size_t>(jd) * dimA, dn = static_cast<std::size_t>(dimB) * dimA;
    int q = 0;
    std::vector<double> XYG = face_quad_points_host(XYZ, face_num, order, q);
    std::vector<double> D;
    std::vector<int> types;
    const std::size_t per_tet = is_constant ? 1 : static_cast<std::size_t>(q);
    if (is_constant) {
        D.assign(dl, 0.0);
        std::vector<double> X0(3, 0.0);
        eval_tensor_points(typename FuncTraits::AggregateType(), PerPoint, Dfnc, X0, 1, 1, dl, dims, user_data, nullptr, nullptr, D, types);
    } else {
        D.assign(dl * q * f, 0.0);
        eval_tensor_points(typename FuncTraits::AggregateType(), FuncTraits::TensorSparsity::value, Dfnc, XYG, q, f, dl, dims, user_data, nullptr, nullptr, D, types);
    }
    std::vector<double> DN(dn * per_tet * f, 0.0);
    for (int r = 0; r < f; ++r) {
        double N[3];
        face_unit_normal(XYZ, r, face_num, N);
        for (std::size_t n = 0; n < per_tet; ++n) {
            const std::size_t src = is_constant ? 0 : n + static_cast<std::size_t>(q) * r, dst = n + per_tet * r;
            const int t = types[src];
            for (int j = 0; j < dimA; ++j)
                for (int k = 0; k < dimB; ++k) {
                    double s = 0;
                    for (int m = 0; m < 3; ++m) {
                        const std::size_t row = static_cast<std::size_t>(m) + 3 * k;
                        double d;
                        if (t == TENSOR_NULL || t == TENSOR_SCALAR) {
                            if (jd != dimA) throw std::runtime_error("Identity tensor defined only for compatible (with same dimensions) operators A and B");
                            d = row == static_cast<std::size_t>(j) ? (t == TENSOR_SCALAR ? D[dl * src] : 1.0) : 0.0;
                        } else d = D[dl * src + row + static_cast<std::size_t>(jd) * j];
                        s += N[m] * d;
                    }
                    DN[dn * dst + k + static_cast<std::size_t>(dimB) * j] = s;
                }
        }
    }
    eval(DN, per_tet);
}

}  // namespace b200
}  // namespace Ani
