// anifem_b200/composite.hpp -- composite element spaces: FemVecT<DIM, FEM> and FemCom<FEM...> under IDEN / GRAD
// (fem/operators.h:71-74, 157-259), included by fem.hpp.
//
// In the reference Operator<OP, FemCom<S_1..S_m>> is block diagonal: part i maps its Nfa_i basis functions to its Dim_i
// components, the parts are stacked (Nfa = sum Nfa_i, Dim = sum Dim_i); FemVecT<DIM, S> is DIM copies of S.  Hence
//     fem3Dtet<Operator<OpA, A>, Operator<OpB, B>>(D)  =  blocks  A[rows of B_j, columns of A_i] = fem3Dtet<OpA(A_i), OpB(B_j)>(D_ji)
// with D_ji the (Dim(B_j) x Dim(A_i)) sub-block of the tensor at the parts' dimension offsets.  Every composite space flattens to
// scalar parts (FemFix), so each block is one batched call of the element kernel on scalar spaces (afb_fem3dtet_batched); blocks
// whose sub-tensor vanishes at every point are skipped.  The composition (flattening, sub-tensors, placement) is written against a
// pluggable block evaluator: the product uses the GPU entry, tests/cxx/test_composite.cpp plugs in the reference build's scalar
// fem3Dtet and compares the composed matrices with the reference's own FemCom / FemVecT operators.
#pragma once
#include <array>
#include <cstddef>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <vector>

namespace Ani {

/// fem/operators.h:71-74
template <int DIM, typename FEMTYPE>
struct FemVecT {
    using Dim = std::integral_constant<int, DIM>;
    using Base = FEMTYPE;
};
template <typename... Types>
struct FemCom {
    using Base = std::tuple<Types...>;
};

namespace b200 {

/// one scalar part of a flattened composite space
struct OpPart {
    int fem;       ///< FEM_P0..P3
    int nfa_off;   ///< first basis function of the part inside the composite space
    int comp;      ///< number of the part = component index (its dimension offset is comp * dim(op))
};

template <typename FEMTYPE> struct Flatten;
template <int F> struct Flatten<FemFix<F>> {
    static constexpr int nf = b200_detail::base_nf(F), ncomp = 1;
    static void append(std::vector<OpPart>& out, int& nfa, int& comp) { out.push_back(OpPart{F, nfa, comp}); nfa += nf; ++comp; }
};
template <int D, int F> struct Flatten<FemVec<D, F>> {
    static constexpr int nf = D * b200_detail::base_nf(F), ncomp = D;
    static void append(std::vector<OpPart>& out, int& nfa, int& comp) { for (int k = 0; k < D; ++k) Flatten<FemFix<F>>::append(out, nfa, comp); }
};
template <int D, typename T> struct Flatten<FemVecT<D, T>> {
    static constexpr int nf = D * Flatten<T>::nf, ncomp = D * Flatten<T>::ncomp;
    static void append(std::vector<OpPart>& out, int& nfa, int& comp) { for (int k = 0; k < D; ++k) Flatten<T>::append(out, nfa, comp); }
};
template <> struct Flatten<FemCom<>> {
    static constexpr int nf = 0, ncomp = 0;
    static void append(std::vector<OpPart>&, int&, int&) {}
};
template <typename T, typename... Rest> struct Flatten<FemCom<T, Rest...>> {
    static constexpr int nf = Flatten<T>::nf + Flatten<FemCom<Rest...>>::nf, ncomp = Flatten<T>::ncomp + Flatten<FemCom<Rest...>>::ncomp;
    static void append(std::vector<OpPart>& out, int& nfa, int& comp) { Flatten<T>::append(out, nfa, comp); Flatten<FemCom<Rest...>>::append(out, nfa, comp); }
};

/// runtime description of Operator<op, composite space>
struct CompositeOp {
    int op = IDEN;
    int nfa = 0, dim = 0, part_dim = 1;
    std::vector<OpPart> parts;
};
template <int OPERATOR, typename FEMTYPE>
inline CompositeOp make_composite_op() {
    static_assert(OPERATOR == IDEN || OPERATOR == GRAD, "composite spaces support IDEN and GRAD (fem/operators.h:152, 257)");
    CompositeOp c;
    c.op = OPERATOR;
    c.part_dim = OPERATOR == GRAD ? 3 : 1;
    int nfa = 0, comp = 0;
    Flatten<FEMTYPE>::append(c.parts, nfa, comp);
    c.nfa = nfa;
    c.dim = comp * c.part_dim;
    return c;
}

/// The composition.  D: tensors at the points of the call, col-major (dimB x dimA) per record, `nrec` records (1 for a constant
/// tensor, q*f per point), one TensorType per record.  eval(opA, femA, opB, femB, Dsub, Ablk) evaluates ONE scalar block with
/// the general sub-tensor Dsub (col-major dB x dA per record, same number of records) into Ablk (nfB_j x nfA_i*f, col-major).
template <typename BlockEval>
void compose_blocks(const CompositeOp& ca, const CompositeOp& cb, int f, std::size_t nrec, const std::vector<double>& D, const std::vector<int>& types,
                    DenseMatrix<double>& A, BlockEval&& eval) {
    const int dimA = ca.dim, dimB = cb.dim, da = ca.part_dim, db = cb.part_dim;
    const std::size_t dl = static_cast<std::size_t>(dimA) * dimB;
    // every record as a full general tensor: TENSOR_NULL = identity, TENSOR_SCALAR = s * identity (equal dimensions only, as in the
    // reference: "Identity tensor defined only for compatible (with same dimensions) operators A and B", diff_tensor.h:315-317)
    std::vector<double> G(dl * nrec, 0.0);
    for (std::size_t p = 0; p < nrec; ++p) {
        const int t = types[p];
        if (t == TENSOR_NULL || t == TENSOR_SCALAR) {
            if (dimA != dimB) throw std::runtime_error("Identity tensor defined only for compatible (with same dimensions) operators A and B");
            const double s = t == TENSOR_SCALAR ? D[dl * p] : 1.0;
            for (int k = 0; k < dimB; ++k) G[dl * p + k + static_cast<std::size_t>(dimB) * k] = s;
        } else {
            for (std::size_t k = 0; k < dl; ++k) G[dl * p + k] = D[dl * p + k];
        }
    }
    for (std::size_t k = 0; k < static_cast<std::size_t>(cb.nfa) * ca.nfa * f; ++k) A.data[k] = 0.0;
    std::vector<double> Dsub(static_cast<std::size_t>(da) * db * nrec), Ablk;
    for (const OpPart& pj : cb.parts)
        for (const OpPart& pi : ca.parts) {
            bool nonzero = false;
            for (std::size_t p = 0; p < nrec; ++p)
                for (int l = 0; l < da; ++l)
                    for (int k = 0; k < db; ++k) {
                        const double v = G[dl * p + (static_cast<std::size_t>(pj.comp) * db + k) + static_cast<std::size_t>(dimB) * (static_cast<std::size_t>(pi.comp) * da + l)];
                        Dsub[(p * da + l) * db + k] = v;
                        nonzero = nonzero || v != 0.0;
                    }
            if (!nonzero) continue;
            const int nfi = b200_detail::base_nf(pi.fem), nfj = b200_detail::base_nf(pj.fem);
            Ablk.assign(static_cast<std::size_t>(nfi) * nfj * f, 0.0);
            eval(ca.op, pi.fem, cb.op, pj.fem, Dsub, Ablk);
            for (int r = 0; r < f; ++r)
                for (int a = 0; a < nfi; ++a)
                    for (int b = 0; b < nfj; ++b)
                        A.data[(pj.nfa_off + b) + static_cast<std::size_t>(cb.nfa) * (static_cast<std::size_t>(r) * ca.nfa + pi.nfa_off + a)] =
                            Ablk[b + static_cast<std::size_t>(nfj) * (static_cast<std::size_t>(r) * nfi + a)];
        }
}

/// physical quadrature points of the tetrahedra on the host (core.inl:249-269: x = sum_k lambda_k P_k), 3 x q x f
inline std::vector<double> quad_points_host(const Tetras<const double>& XYZ, int order, int& q) {
    q = afb_tet_quadrature(order, nullptr, nullptr, 0);
    if (q < 0) throw std::runtime_error("Numerical tetrahedron integration formula implemented only for 0 <= order <= 20");
    std::vector<double> lam(static_cast<std::size_t>(4) * q), w(q), X(static_cast<std::size_t>(3) * q * XYZ.fusion);
    afb_tet_quadrature(order, lam.data(), w.data(), q);
    const double* P[4] = {XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3};
    for (int r = 0; r < XYZ.fusion; ++r)
        for (int n = 0; n < q; ++n)
            for (int k = 0; k < 3; ++k) {
                double s = 0;
                for (int v = 0; v < 4; ++v) s += lam[4 * n + v] * P[v][k + 3 * r];
                X[k + 3 * (n + static_cast<std::size_t>(q) * r)] = s;
            }
    return X;
}

/// front end shared by the product (eval = GPU entry) and the CPU test (eval = reference build)
template <typename FuncTraits, typename Functor, typename BlockEval>
void fem3Dtet_composite(const CompositeOp& ca, const CompositeOp& cb, const Tetras<const double>& XYZ, const Functor& Dfnc, DenseMatrix<double>& A, int order,
                        void* user_data, BlockEval&& eval) {
    const int f = XYZ.fusion;
    if (f <= 0) return;
    if (A.size < static_cast<std::size_t>(ca.nfa) * cb.nfa * f)
        throw std::runtime_error("Not enough memory for local matrix, expected size = " + std::to_string(ca.nfa * cb.nfa * f) + " but A has size = " +
                                 std::to_string(A.size));
    A.nRow = cb.nfa; A.nCol = static_cast<std::size_t>(ca.nfa) * f;
    constexpr bool is_constant = FuncTraits::IsConstant::value && FuncTraits::AggregateType::value == OnePointTensor;
    const TensorDims dims{static_cast<std::size_t>(cb.dim), static_cast<std::size_t>(ca.dim)};
    const std::size_t dl = static_cast<std::size_t>(ca.dim) * cb.dim;
    std::vector<double> D;
    std::vector<int> types;
    int q = 0;
    std::vector<double> XYG = quad_points_host(XYZ, order, q);
    std::size_t nrec;
    if (is_constant) {
        nrec = 1;
        D.assign(dl, 0.0);
        std::vector<double> X0(3, 0.0);
        eval_tensor_points(typename FuncTraits::AggregateType(), PerPoint, Dfnc, X0, 1, 1, dl, dims, user_data, nullptr, nullptr, D, types);
    } else {
        nrec = static_cast<std::size_t>(q) * f;
        D.assign(dl * nrec, 0.0);
        std::vector<double> xyl(static_cast<std::size_t>(4) * q), wg(q);
        afb_tet_quadrature(order, xyl.data(), wg.data(), q);
        eval_tensor_points(typename FuncTraits::AggregateType(), FuncTraits::TensorSparsity::value, Dfnc, XYG, q, f, dl, dims, user_data, xyl.data(), wg.data(), D, types);
    }
    compose_blocks(ca, cb, f, nrec, D, types, A, eval);
}

}  // namespace b200

/// Operator<OP, FemVecT<DIM, FEM>> / Operator<OP, FemCom<...>> (fem/operators.h:157-259)
template <int OPERATOR, int DIM, typename FEMTYPE>
struct Operator<OPERATOR, FemVecT<DIM, FEMTYPE>> {
    static constexpr bool composite = true;
    static constexpr int op = OPERATOR;
    using Space = FemVecT<DIM, FEMTYPE>;
    using Nfa = std::integral_constant<int, b200::Flatten<Space>::nf>;
    using Dim = std::integral_constant<int, b200::Flatten<Space>::ncomp * (OPERATOR == GRAD ? 3 : 1)>;
    static b200::CompositeOp describe() { return b200::make_composite_op<OPERATOR, Space>(); }
};
template <int OPERATOR, typename... Types>
struct Operator<OPERATOR, FemCom<Types...>> {
    static constexpr bool composite = true;
    static constexpr int op = OPERATOR;
    using Space = FemCom<Types...>;
    using Nfa = std::integral_constant<int, b200::Flatten<Space>::nf>;
    using Dim = std::integral_constant<int, b200::Flatten<Space>::ncomp * (OPERATOR == GRAD ? 3 : 1)>;
    static b200::CompositeOp describe() { return b200::make_composite_op<OPERATOR, Space>(); }
};

/// Runtime twins (fem/fem_space.h: ComplexFemSpace, VectorFemSpace, FemSpace::operator* / operator^, getOP): a space assembled at
/// run time from scalar / vector parts, e.g. (FemSpace(FEM_P2) ^ 3) * FemSpace(FEM_P1) for Taylor-Hood.
struct ApplyOpComposite {
    b200::CompositeOp c;
    unsigned Nfa() const { return static_cast<unsigned>(c.nfa); }
    unsigned Dim() const { return static_cast<unsigned>(c.dim); }
};
struct ComplexFemSpace {
    std::vector<FemSpace> parts;   ///< first-level factors in order
    ComplexFemSpace() = default;
    ComplexFemSpace(const FemSpace& s) : parts{s} {}
    ComplexFemSpace(std::vector<FemSpace> p) : parts(std::move(p)) {}
    unsigned dofMapSize() const { unsigned n = 0; for (auto& s : parts) n += s.dofMapSize(); return n; }
    /// operator of the whole space: block diagonal over the scalar parts (IDEN, GRAD)
    ApplyOpComposite getOP(OperatorType op) const {
        if (op != IDEN && op != GRAD) throw std::runtime_error("composite spaces support IDEN and GRAD");
        ApplyOpComposite a;
        a.c.op = op; a.c.part_dim = op == GRAD ? 3 : 1;
        int nfa = 0, comp = 0;
        for (auto& s : parts)
            for (int k = 0; k < s.vec; ++k) { a.c.parts.push_back(b200::OpPart{s.fem, nfa, comp}); nfa += b200_detail::base_nf(s.fem); ++comp; }
        a.c.nfa = nfa; a.c.dim = comp * a.c.part_dim;
        return a;
    }
    /// local dof map: the product of the factors' maps with the reference's simplifications
    DofT::DofMap dofMap() const {
        std::vector<DofT::DofMap> m;
        for (auto& s : parts) m.push_back(s.dofMap());
        return DofT::merge_with_simplifications(m);
    }
    ComplexFemSpace operator*(const ComplexFemSpace& o) const {
        ComplexFemSpace r = *this;
        for (auto& s : o.parts) {
            if (!r.parts.empty() && r.parts.back().fem == s.fem) r.parts.back().vec += s.vec;   // equal neighbours fuse into a vector space
            else r.parts.push_back(s);
        }
        return r;
    }
};
inline ComplexFemSpace operator*(const FemSpace& a, const FemSpace& b) { return ComplexFemSpace(a) * ComplexFemSpace(b); }
inline ComplexFemSpace operator*(const FemSpace& a, const ComplexFemSpace& b) { return ComplexFemSpace(a) * b; }

namespace b200 {
/// description of any operator as a composite (simple spaces: FemFix -> one part, FemVec<3,F> -> three parts)
template <typename Op, bool = Op::composite> struct Describe;
template <typename Op> struct Describe<Op, true> { static CompositeOp get() { return Op::describe(); } };
template <typename Op> struct Describe<Op, false> {
    static CompositeOp get() {
        if (Op::op != IDEN && Op::op != GRAD) throw std::runtime_error("composite spaces support IDEN and GRAD");
        CompositeOp c;
        c.op = Op::op; c.part_dim = Op::op == GRAD ? 3 : 1;
        int nfa = 0, comp = 0;
        for (int k = 0; k < Op::vec; ++k) { c.parts.push_back(OpPart{Op::fem, nfa, comp}); nfa += b200_detail::base_nf(Op::fem); ++comp; }
        c.nfa = nfa; c.dim = comp * c.part_dim;
        return c;
    }
};
}  // namespace b200

}  // namespace Ani
