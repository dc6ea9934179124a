// anifem_b200/fem.hpp -- C++ mirror of the AniFem++ element API on top of the C ABI (include/anifem_b200.h).
//
// Same names, template arguments, argument meaning, layouts and error behaviour as the reference so that a user's
// local assembler compiles against it unchanged for the operators in scope:
//   Ani::fem3Dtet<OpA, OpB, FuncTraits>(XYZ | XY0..XY3, Dfnc, A, order, user_data)   fem/operations/int_tet.h:17-78
//   Ani::Operator<GRAD|IDEN|DIV, FemFix<FEM_P0..P3> | FemVec<3,FEM_Pk>>               fem/operators.h:50-67,127-155,320-353
//   Ani::DfuncTraits<...>, TensorType, DenseMatrix<>, Tetras<>, make_tetras            fem/diff_tensor.h:17-53, fem_memory.h:55-130, geometry.h:96-200
// The arithmetic runs on the GPU (afb_fem3dtet_batched); `fusion` (XYZ.fusion tets per call) is the batch axis.
// Tensor callbacks are host functors exactly like the reference's; the shim evaluates them at the quadrature points
// (afb_quad_points) into the FusiveTensor data layout and ships the data (compatibility path, not the
// timed path -- hot loops should pass coefficient arrays through the C ABI directly).
#pragma once
#include <array>
#include <cstddef>
#include <functional>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../../include/anifem_b200.h"
#include "memory.hpp"
#include "dofmap.hpp"

namespace Ani {

enum FiniteElement { FEM_P0 = 1, FEM_P1 = 2, FEM_P2 = 3, FEM_P3 = 4 };
enum OperatorType { IDEN = 1, GRAD = 2, DIV = 3 };
enum TensorType { TENSOR_NULL = 1, TENSOR_SCALAR = 2, TENSOR_SYMMETRIC = 3, TENSOR_GENERAL = 4 };
enum TensorTypeSparsity { PerPoint = 0, PerTetra = -1, PerSelection = -2 };

template <int OP> using FemFix = std::integral_constant<int, OP>;
template <int DIM, int OP> struct FemVec { using Dim = std::integral_constant<int, DIM>; using Base = FemFix<OP>; };

namespace b200_detail {
constexpr int base_nf(int fem) { return fem == FEM_P0 ? 1 : fem == FEM_P1 ? 4 : fem == FEM_P2 ? 10 : 20; }
template <typename FEM> struct FemInfo;
template <int F> struct FemInfo<FemFix<F>> { static constexpr int fem = F, vec = 1; };
template <int D, int F> struct FemInfo<FemVec<D, F>> { static constexpr int fem = F, vec = D; static_assert(D == 3, "only FemVec<3,.> is supported"); };
}  // namespace b200_detail

template <int OPERATOR, typename FEMTYPE>
struct Operator {
    static constexpr bool composite = false;   ///< FemVecT / FemCom spaces are specialised in composite.hpp
    static constexpr int op = OPERATOR, fem = b200_detail::FemInfo<FEMTYPE>::fem, vec = b200_detail::FemInfo<FEMTYPE>::vec;
    static_assert(OPERATOR == IDEN || OPERATOR == GRAD || (OPERATOR == DIV && vec == 3), "operator out of scope of the B200 path");
    using Nfa = std::integral_constant<int, vec * b200_detail::base_nf(fem)>;
    using Dim = std::integral_constant<int, OPERATOR == IDEN ? vec : (OPERATOR == GRAD ? 3 * vec : 1)>;
};

/// fem/diff_tensor.h:29-53: signature class of the tensor callback
enum TensorTypeAggregate {
    OnePointTensor = 0,   ///< TensorType f(const Coord<>& X, double* D, TensorDims Ddims, void* user_data, int iTet), once per point
    FusiveTensor = 1      ///< TensorType f(ArrayView<> X, ArrayView<> D, TensorDims Ddims, void* user_data, const AniMemory<>& mem), once per call
};
template <int TensorTypeSparse = PerPoint, bool isConstant = false, long idim = -1, long jdim = -1, int aggregateType = OnePointTensor>
struct DfuncTraitsCommon {
    using AggregateType = std::integral_constant<int, aggregateType>;
    using IsConstant = std::integral_constant<bool, isConstant>;
    /// > 0: the TensorType is known at compile time; <= 0: one of TensorTypeSparsity (per point / per tet / one type per call)
    using TensorSparsity = std::integral_constant<int, TensorTypeSparse>;
    using iD = std::integral_constant<long, idim>;
    using jD = std::integral_constant<long, jdim>;
};
template <int TensorTypeSparse = PerPoint, bool isConstant = false, long idim = -1, long jdim = -1>
using DfuncTraits = DfuncTraitsCommon<TensorTypeSparse, isConstant, idim, jdim>;
template <long idim = -1, long jdim = -1>
using DfuncTraitsFusive = DfuncTraitsCommon<PerSelection, false, idim, jdim, FusiveTensor>;

using TensorDims = std::pair<std::size_t, std::size_t>;
template <typename Scalar = double> using Coord = std::array<Scalar, 3>;

/// fem/fem_memory.h:15-52: pointer + size
template <typename T = double>
struct ArrayView {
    T* data = nullptr;
    std::size_t size = 0;
    ArrayView() = default;
    ArrayView(T* d, std::size_t n) : data(d), size(n) {}
    T& operator[](std::size_t i) { return data[i]; }
    const T& operator[](std::size_t i) const { return data[i]; }
    T* begin() const { return data; }
    T* end() const { return data + size; }
};
/// fem/fem_memory.h:195-213: what a FusiveTensor callback sees of the running call.  Filled here: q, f, XYG (physical quadrature
/// points, 3 x q x f) and XYL / WG (the rule); the scratch arrays of the reference's host evaluation do not exist (the element
/// kernels run on the device) and stay empty.
template <typename ScalarType = double, typename IndexType = int>
struct AniMemory {
    using Scalar = ScalarType;
    using ArrayR = ArrayView<Scalar>;
    using ArrayI = ArrayView<IndexType>;
    ArrayR XYP, PSI, XYG, DET, MES, NRM, U, V, DIFF, DU, XYL, WG, extraR;
    ArrayI extraI;
    std::size_t q = 0, f = 1;
};

// column-major dense matrix view (fem/fem_memory.h:55-130)
template <typename ScalarType = double>
struct DenseMatrix {
    ScalarType* data = nullptr;
    std::size_t nRow = 0, nCol = 0, size = 0;
    DenseMatrix() = default;
    DenseMatrix(ScalarType* d, std::size_t r, std::size_t c) : data(d), nRow(r), nCol(c), size(r * c) {}
    DenseMatrix(ScalarType* d, std::size_t r, std::size_t c, std::size_t sz) : data(d), nRow(r), nCol(c), size(sz) {}
    ScalarType& operator()(std::size_t i, std::size_t j) { return data[i + nRow * j]; }
    const ScalarType& operator()(std::size_t i, std::size_t j) const { return data[i + nRow * j]; }
    void SetZero() { for (std::size_t i = 0; i < nRow * nCol; ++i) data[i] = 0; }
};

// coordinates of `fusion` tetrahedra: XYk is 3 x fusion col-major (fem/geometry.h:96-200)
template <typename ScalarType = const double>
struct Tetras {
    ScalarType *XY0, *XY1, *XY2, *XY3;
    int fusion = 0;
};
inline Tetras<const double> make_tetras(const double* XY0, const double* XY1, const double* XY2, const double* XY3, int count = 1) {
    return Tetras<const double>{XY0, XY1, XY2, XY3, count};
}

namespace b200 {
// context used by the free functions: one per calling thread (a context is single-owner: one stream, its own buffers), so that
// local assemblers evaluated concurrently (Assembler with num_threads > 1, like the reference's AssemblerP) may call fem3Dtet;
// device 0 unless ANIFEM_B200_DEVICE is set
struct ContextHolder {
    afb_ctx* c = nullptr;
    ContextHolder() {
        int dev = 0;
        if (const char* s = std::getenv("ANIFEM_B200_DEVICE")) dev = std::atoi(s);
        if (afb_ctx_create(dev, nullptr, &c) != 0) throw std::runtime_error(std::string("anifem_b200: ") + afb_last_error(nullptr));
    }
    ~ContextHolder() { if (c) afb_ctx_destroy(c); }
    ContextHolder(const ContextHolder&) = delete;
    ContextHolder& operator=(const ContextHolder&) = delete;
};
inline afb_ctx* default_context() {
    static thread_local ContextHolder holder;
    return holder.c;
}
inline void check(afb_ctx* c, int rc) {
    if (rc < 0 && rc != -1) throw std::runtime_error(afb_last_error(c));
}
}  // namespace b200

// ---- memory arguments of the reference overloads ----------------------------------------------------------------------
// PlainMemory / PlainMemoryX / DynMem (memory.hpp) are the reference's views and planners (fem/fem_memory.h:262-520).  The scratch
// of the element kernels lives on the device, so fem3Dtet_memory_requirements reports zero sizes and the overloads that take a
// memory argument do not draw from it: reference call sites that size, allocate and pass these objects compile and run unchanged.
template <typename ScalarType, typename IndexType>
void* PlainMemoryX<ScalarType, IndexType>::allocateFromRaw(void* mem_in, std::size_t mem_sz, std::size_t dsize, std::size_t isize, std::size_t msize) {
    char *p = static_cast<char*>(mem_in), *end = p + mem_sz;
    ScalarType* d = nullptr;
    IndexType* i = nullptr;
    DenseMatrix<ScalarType>* m = nullptr;
    if (dsize && !(d = mem_detail::carve<ScalarType>(p, end, dsize))) return nullptr;
    if (isize && !(i = mem_detail::carve<IndexType>(p, end, isize))) return nullptr;
    if (msize && !(m = mem_detail::carve<DenseMatrix<ScalarType>>(p, end, msize))) return nullptr;
    if (dsize) { ddata = d; dSize = dsize; }
    if (isize) { idata = i; iSize = isize; }
    if (msize) { mdata = m; mSize = msize; for (std::size_t k = 0; k < msize; ++k) new (m + k) DenseMatrix<ScalarType>(); }
    return p;
}
template <typename ScalarType, typename IndexType>
std::size_t PlainMemoryX<ScalarType, IndexType>::enoughRawSize() const {
    return mem_detail::worst_bytes<ScalarType>(dSize) + mem_detail::worst_bytes<IndexType>(iSize) + mem_detail::worst_bytes<DenseMatrix<ScalarType>>(mSize);
}

// ---- runtime twins of the operators (fem/fem_space.h:170-306, ApplyOpBase / FemSpace::getOP) -------------------------
struct ApplyOpBase {
    int op = IDEN, fem = FEM_P1, vec = 1;
    ApplyOpBase() = default;
    ApplyOpBase(int op_, int fem_, int vec_ = 1) : op(op_), fem(fem_), vec(vec_) {
        int nfa = 0, dim = 0;
        if (afb_op_dims(op, fem, vec, &nfa, &dim) != 0) throw std::runtime_error("operator / space out of scope of the B200 path");
    }
    unsigned Nfa() const { int nfa = 0, dim = 0; afb_op_dims(op, fem, vec, &nfa, &dim); return static_cast<unsigned>(nfa); }
    unsigned Dim() const { int nfa = 0, dim = 0; afb_op_dims(op, fem, vec, &nfa, &dim); return static_cast<unsigned>(dim); }
};
/// runtime description of a space: FemSpace{FEM_P2} (scalar) or FemSpace{FEM_P2, 3} (= P2^3); getOP mirrors fem_space.h:502
struct FemSpace {
    int fem = FEM_P1, vec = 1;
    FemSpace() = default;
    FemSpace(int fem_, int vec_ = 1) : fem(fem_), vec(vec_) {}
    ApplyOpBase getOP(OperatorType op) const { return ApplyOpBase(op, fem, vec); }
    unsigned dofMapSize() const { return static_cast<unsigned>(vec * b200_detail::base_nf(fem)); }
    FemSpace operator^(int k) const { return FemSpace(fem, vec * k); }
    /// local dof map of the space (BaseFemSpace::dofMap, fem_space.h): P0 = one cell dof, P1 = node dofs, P2 = + one dof per edge,
    /// P3 = + an oriented pair per edge and one per face; a vector space is `vec` copies.  These are the orders the device
    /// numbering kernels use (afb_ctx.cu) and the rows / columns of every element matrix of this library.
    DofT::DofMap dofMap() const {
        std::array<DofT::uint, DofT::NGEOM_TYPES> n{{0, 0, 0, 0, 0, 0}};
        if (fem == FEM_P0) n[5] = 1;
        if (fem == FEM_P1 || fem == FEM_P2 || fem == FEM_P3) n[0] = 1;
        if (fem == FEM_P2) n[1] = 1;
        if (fem == FEM_P3) { n[1] = 2; n[3] = 1; }
        DofT::DofMap base(std::make_shared<DofT::UniteDofMap>(n));
        return vec == 1 ? base : DofT::pow(base, static_cast<DofT::uint>(vec));
    }
};

namespace b200 {
/// core shared by the compile-time and the runtime front ends
/// face_num < 0: volume integral (fem3Dtet); 0..3: surface integral over that face of every tet (fem3Dface, int_face.inl:160-199)
/// evaluation of the callback at the points of the call: D[dl*(n + q*r)], one TensorType per point
template <typename Functor>
void eval_tensor_points(std::integral_constant<int, OnePointTensor>, int sparsity, const Functor& Dfnc, std::vector<double>& XYG, int q, int f, std::size_t dl,
                        TensorDims dims, void* user_data, const double*, const double*, std::vector<double>& D, std::vector<int>& types) {
    for (int r = 0; r < f; ++r)
        for (int n = 0; n < q; ++n) {
            const std::size_t p = n + static_cast<std::size_t>(q) * r;
            // PerSelection (diff_tensor.h:657-690): one type for the whole call, the callback is told tet 0
            types.push_back(Dfnc(std::array<double, 3>{XYG[3 * p], XYG[3 * p + 1], XYG[3 * p + 2]}, D.data() + dl * p, dims, user_data, sparsity == PerSelection ? 0 : r));
        }
    if (sparsity == PerSelection) for (auto& t : types) t = types[0];
    else if (sparsity == PerTetra)   // the type of a tet is the type of its first point (diff_tensor.h:559-566)
        for (int r = 0; r < f; ++r) for (int n = 1; n < q; ++n) types[n + static_cast<std::size_t>(q) * r] = types[static_cast<std::size_t>(q) * r];
}
/// FusiveTensor (diff_tensor.h:101-135, 900-905): ONE call fills the tensors of all points; the callback writes the col-major
/// (Ddims.first x Ddims.second) matrix of point (n, r) at D[first*second*(n + q*r)]
template <typename Functor>
void eval_tensor_points(std::integral_constant<int, FusiveTensor>, int, const Functor& Dfnc, std::vector<double>& XYG, int q, int f, std::size_t dl,
                        TensorDims dims, void* user_data, const double* xyl, const double* wg, std::vector<double>& D, std::vector<int>& types) {
    AniMemory<double, int> mem;
    mem.q = static_cast<std::size_t>(q); mem.f = static_cast<std::size_t>(f);
    mem.XYG = ArrayView<double>(XYG.data(), XYG.size());
    mem.XYL = ArrayView<double>(const_cast<double*>(xyl), xyl ? static_cast<std::size_t>(4) * q : 0);
    mem.WG = ArrayView<double>(const_cast<double*>(wg), wg ? static_cast<std::size_t>(q) : 0);
    const int t = Dfnc(ArrayView<double>(XYG.data(), XYG.size()), ArrayView<double>(D.data(), D.size()), dims, user_data, mem);
    (void)dl;
    types.assign(static_cast<std::size_t>(q) * f, t);
}

template <typename FuncTraits, typename Functor>
void fem3Dtet_core(const ApplyOpBase& oa, const ApplyOpBase& ob, const Tetras<const double>& XYZ, const Functor& Dfnc,
                   DenseMatrix<double>& A, int order, void* user_data, int face_num = -1) {
    constexpr bool is_constant = FuncTraits::IsConstant::value && FuncTraits::AggregateType::value == OnePointTensor;
    const int f = XYZ.fusion;
    if (f <= 0) return;
    if (face_num > 3) throw std::runtime_error("Wrong face index");
    const int nfa = static_cast<int>(oa.Nfa()), nfb = static_cast<int>(ob.Nfa()), idim = static_cast<int>(oa.Dim()), jdim = static_cast<int>(ob.Dim());
    if (A.size < static_cast<std::size_t>(nfa) * nfb * f)
        throw std::runtime_error("Not enough memory for local matrix, expected size = " + std::to_string(nfa * nfb * f) +
                                 " but A has size = " + std::to_string(A.size));
    A.nRow = nfb; A.nCol = static_cast<std::size_t>(nfa) * f;
    afb_ctx* ctx = default_context();
    const int q = face_num < 0 ? afb_tet_quadrature(order, nullptr, nullptr, 0) : afb_tri_quadrature(order, nullptr, nullptr, 0);
    if (q < 0) throw std::runtime_error(face_num < 0 ? "Numerical tetrahedron integration formula implemented only for 0 <= order <= 20"
                                                     : "Numerical triangle integration formula implemented only for 0 <= order <= 20");
    // evaluate the callback: once (constant tensor) or per quadrature point of every tet, r-major / n-minor like the
    // reference (fem/diff_tensor.h:492-520)
    const TensorDims dims{static_cast<std::size_t>(jdim), static_cast<std::size_t>(idim)};
    const std::size_t dl = static_cast<std::size_t>(idim) * jdim;
    std::vector<double> D;
    std::vector<int> types;
    int layout;
    if (is_constant) {
        layout = AFB_COEF_CONST;
        D.assign(dl, 0.0);
        std::vector<double> X0(3, 0.0);   // a constant tensor is asked for once (a FusiveTensor callback never takes this branch)
        eval_tensor_points(typename FuncTraits::AggregateType(), PerPoint, Dfnc, X0, 1, 1, dl, dims, user_data, nullptr, nullptr, D, types);
    } else {
        layout = AFB_COEF_PER_POINT;
        std::vector<double> XYG(static_cast<std::size_t>(3) * q * f);
        if (face_num < 0) check(ctx, afb_quad_points(ctx, order, f, XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, XYG.data(), AFB_HOST));
        else {
            // points of the triangle rule lifted to the face {face, face+1, face+2 mod 4} (int_face.inl:175-180, core.inl:249-269)
            std::vector<double> p(static_cast<std::size_t>(3) * q), w(q);
            afb_tri_quadrature(order, p.data(), w.data(), q);
            const double* X[4] = {XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3};
            for (int r = 0; r < f; ++r)
                for (int n = 0; n < q; ++n)
                    for (int k = 0; k < 3; ++k) {
                        double lam[4] = {0, 0, 0, 0};
                        for (int m = 0; m < 3; ++m) lam[(face_num + m) % 4] = p[3 * n + m];
                        double sx = X[0][k + 3 * r];
                        for (int l = 1; l < 4; ++l) sx += lam[l] * (X[l][k + 3 * r] - X[0][k + 3 * r]);
                        XYG[k + 3 * (n + static_cast<std::size_t>(q) * r)] = sx;
                    }
        }
        D.assign(dl * q * f, 0.0);
        types.reserve(static_cast<std::size_t>(q) * f);
        std::vector<double> xyl, wg;
        if (FuncTraits::AggregateType::value == FusiveTensor && face_num < 0) {
            xyl.resize(static_cast<std::size_t>(4) * q); wg.resize(q);
            afb_tet_quadrature(order, xyl.data(), wg.data(), q);
        }
        eval_tensor_points(typename FuncTraits::AggregateType(), FuncTraits::TensorSparsity::value, Dfnc, XYG, q, f, dl, dims, user_data,
                           xyl.empty() ? nullptr : xyl.data(), wg.empty() ? nullptr : wg.data(), D, types);
    }
    bool uniform = true;
    for (int t : types) uniform = uniform && t == types[0];
    int ttype = types[0];
    if (!uniform || ttype == TENSOR_SCALAR) {
        // SCALAR data are compacted to 1 value per point; mixed types are densified to GENERAL
        const std::size_t np = types.size();
        if (uniform) {
            std::vector<double> S(np);
            for (std::size_t p = 0; p < np; ++p) S[p] = D[dl * p];
            D.swap(S);
        } else {
            if (jdim != idim && !(nfa == 1 && idim == 1))
                for (int t : types) if (t <= TENSOR_SCALAR)
                    throw std::runtime_error("Identity tensor defined only for compatible (with same dimensions) operators A and B");
            for (std::size_t p = 0; p < np; ++p) if (types[p] <= TENSOR_SCALAR) {
                const double s = types[p] == TENSOR_SCALAR ? D[dl * p] : 1.0;
                for (std::size_t k = 0; k < dl; ++k) D[dl * p + k] = 0;
                if (jdim == idim) for (int k = 0; k < jdim; ++k) D[dl * p + k + jdim * k] = s;
                else for (int k = 0; k < jdim; ++k) D[dl * p + k] = s;  // OpA = IDEN(P0) broadcast
            }
            ttype = TENSOR_GENERAL;
        }
    }
    afb_form fm{oa.op, oa.fem, oa.vec, ob.op, ob.fem, ob.vec, order, ttype, layout, AFB_HOST, D.data(), 1.0, 0, 0};
    if (face_num < 0) check(ctx, afb_fem3dtet_batched(ctx, &fm, f, XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, A.data, AFB_HOST));
    else {
        std::vector<int32_t> faces(f, face_num);
        check(ctx, afb_fem3dface_batched(ctx, &fm, f, faces.data(), XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, A.data, AFB_HOST));
    }
}
}  // namespace b200

/// Elemental matrix of int_T (D OpA(u)) . OpB(v) dx for XYZ.fusion tetrahedra; A is nfB x (nfA*fusion) col-major.
/// Dfnc: TensorType(const std::array<double,3>& x, double* Dmem, TensorDims Ddims, void* user_data, int iTet) with
/// Dmem a col-major (Ddims.first x Ddims.second) = (Dim(OpB) x Dim(OpA)) matrix (fem/operations/int_tet.h:31-47).
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
typename std::enable_if<!(OpA::composite || OpB::composite)>::type fem3Dtet(const Tetras<const double>& XYZ, const Functor& Dfnc, DenseMatrix<double>& A,
                                                                             int order = 5, void* user_data = nullptr) {
    b200::fem3Dtet_core<FuncTraits>(ApplyOpBase(OpA::op, OpA::fem, OpA::vec), ApplyOpBase(OpB::op, OpB::fem, OpB::vec), XYZ, Dfnc, A,
                        order, user_data);
}

template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dtet(const DenseMatrix<double>& XY0, const DenseMatrix<double>& XY1, const DenseMatrix<double>& XY2, const DenseMatrix<double>& XY3,
              const Functor& Dfnc, DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    fem3Dtet<OpA, OpB, FuncTraits>(make_tetras(XY0.data, XY1.data, XY2.data, XY3.data, static_cast<int>(XY0.nCol)), Dfnc, A, order, user_data);
}

/// reference overloads with caller memory (int_tet.h:17-29,111-122; dyn_ops.h:18-20): the memory argument is not used
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor, typename ScalarType, typename IndexType>
void fem3Dtet(const Tetras<const double>& XYZ, const Functor& Dfnc, DenseMatrix<double>& A, PlainMemory<ScalarType, IndexType>, int order = 5,
              void* user_data = nullptr) {
    fem3Dtet<OpA, OpB, FuncTraits>(XYZ, Dfnc, A, order, user_data);
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor, typename ScalarType, typename IndexType>
void fem3Dtet(const DenseMatrix<double>& XY0, const DenseMatrix<double>& XY1, const DenseMatrix<double>& XY2, const DenseMatrix<double>& XY3,
              const Functor& Dfnc, DenseMatrix<double>& A, PlainMemory<ScalarType, IndexType>, int order = 5, void* user_data = nullptr) {
    fem3Dtet<OpA, OpB, FuncTraits>(make_tetras(XY0.data, XY1.data, XY2.data, XY3.data, static_cast<int>(XY0.nCol)), Dfnc, A, order, user_data);
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor, typename ScalarType, typename IndexType>
void fem3Dtet(const Tetras<const double>& XYZ, const Functor& Dfnc, DenseMatrix<double>& A, DynMem<ScalarType, IndexType>&, int order = 5,
              void* user_data = nullptr) {
    fem3Dtet<OpA, OpB, FuncTraits>(XYZ, Dfnc, A, order, user_data);
}
/// runtime operators (dyn_ops.h:21-23, int_tet.h:144-160)
template <typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dtet(const Tetras<const double>& XYZ, const ApplyOpBase& applyOpU, const ApplyOpBase& applyOpV, const Functor& Dfnc, DenseMatrix<double>& A,
              DynMem<>& /*wmem*/, int order = 5, void* user_data = nullptr) {
    b200::fem3Dtet_core<FuncTraits>(applyOpU, applyOpV, XYZ, Dfnc, A, order, user_data);
}
template <typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dtet(const Tetras<const double>& XYZ, const ApplyOpBase& applyOpU, const ApplyOpBase& applyOpV, const Functor& Dfnc, DenseMatrix<double>& A,
              PlainMemoryX<> /*mem*/, int order = 5, void* user_data = nullptr) {
    b200::fem3Dtet_core<FuncTraits>(applyOpU, applyOpV, XYZ, Dfnc, A, order, user_data);
}
/// Elemental matrix of the surface integral int_f (D OpA(u)) . OpB(v) over face face_num of every tet: face k = vertices
/// {k, k+1, k+2 mod 4} (fem/operations/int_face.h:15-29,49-66); the callback sees the points of the triangle rule on the face.
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dface(const Tetras<const double>& XYZ, int face_num, const Functor& Dfnc, DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    if (face_num < 0) throw std::runtime_error("Wrong face index");
    b200::fem3Dtet_core<FuncTraits>(ApplyOpBase(OpA::op, OpA::fem, OpA::vec), ApplyOpBase(OpB::op, OpB::fem, OpB::vec), XYZ, Dfnc, A,
                        order, user_data, face_num);
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dface(const DenseMatrix<double>& XY0, const DenseMatrix<double>& XY1, const DenseMatrix<double>& XY2, const DenseMatrix<double>& XY3,
               int face_num, const Functor& Dfnc, DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    fem3Dface<OpA, OpB, FuncTraits>(make_tetras(XY0.data, XY1.data, XY2.data, XY3.data, static_cast<int>(XY0.nCol)), face_num, Dfnc, A, order, user_data);
}
/// caller-memory and runtime-operator overloads (int_face.h:22-29,90-93; dyn_ops.h:26-32): the memory argument is not used
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor, typename ScalarType, typename IndexType>
void fem3Dface(const Tetras<const double>& XYZ, int face_num, const Functor& Dfnc, DenseMatrix<double>& A, PlainMemory<ScalarType, IndexType>,
               int order = 5, void* user_data = nullptr) {
    fem3Dface<OpA, OpB, FuncTraits>(XYZ, face_num, Dfnc, A, order, user_data);
}
template <typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dface(const Tetras<const double>& XYZ, int face_num, const ApplyOpBase& applyOpU, const ApplyOpBase& applyOpV, const Functor& Dfnc,
               DenseMatrix<double>& A, PlainMemoryX<> /*mem*/, int order = 5, void* user_data = nullptr) {
    if (face_num < 0) throw std::runtime_error("Wrong face index");
    b200::fem3Dtet_core<FuncTraits>(applyOpU, applyOpV, XYZ, Dfnc, A, order, user_data, face_num);
}
template <typename OpA, typename OpB, typename ScalarType = double, typename IndexType = int>
PlainMemory<ScalarType, IndexType> fem3Dface_memory_requirements(int /*order*/, int /*fusion*/ = 1) { return PlainMemory<ScalarType, IndexType>(); }

/// no host scratch is needed: both sizes are zero (int_tet.h:154-160)
template <typename OpA, typename OpB, typename ScalarType = double, typename IndexType = int>
PlainMemory<ScalarType, IndexType> fem3Dtet_memory_requirements(int /*order*/, int /*fusion*/ = 1) { return PlainMemory<ScalarType, IndexType>(); }
template <typename FuncTraits = DfuncTraits<>>
PlainMemoryX<> fem3Dtet_memory_requirements(const ApplyOpBase&, const ApplyOpBase&, int /*order*/, int /*fusion*/ = 1) { return PlainMemoryX<>(); }

}  // namespace Ani

#include "composite.hpp"   // FemVecT / FemCom operators (needs Operator, Tetras, DenseMatrix, eval_tensor_points)

namespace Ani {
/// fem3Dtet with a composite space on either side (fem/operators.h:157-259): block by block on the scalar parts, every block one
/// batched call of the element kernel with the general sub-tensor of the block
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
typename std::enable_if<(OpA::composite || OpB::composite)>::type fem3Dtet(const Tetras<const double>& XYZ, const Functor& Dfnc, DenseMatrix<double>& A,
                                                                            int order = 5, void* user_data = nullptr) {
    const b200::CompositeOp ca = b200::Describe<OpA>::get(), cb = b200::Describe<OpB>::get();
    const bool is_constant = FuncTraits::IsConstant::value && FuncTraits::AggregateType::value == OnePointTensor;
    const int f = XYZ.fusion;
    afb_ctx* ctx = b200::default_context();
    b200::fem3Dtet_composite<FuncTraits>(ca, cb, XYZ, Dfnc, A, order, user_data,
        [&](int opA, int femA, int opB, int femB, const std::vector<double>& Dsub, std::vector<double>& Ablk) {
            afb_form fm{opA, femA, 1, opB, femB, 1, order, TENSOR_GENERAL, is_constant ? AFB_COEF_CONST : AFB_COEF_PER_POINT, AFB_HOST, Dsub.data(), 1.0, 0, 0};
            b200::check(ctx, afb_fem3dtet_batched(ctx, &fm, f, XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, Ablk.data(), AFB_HOST));
        });
}
}  // namespace Ani

namespace Ani {
/// runtime-operator form on composite spaces (dyn_ops.h:26-32 with ComplexFemSpace::getOP)
template <typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3Dtet(const Tetras<const double>& XYZ, const ApplyOpComposite& applyOpU, const ApplyOpComposite& applyOpV, const Functor& Dfnc, DenseMatrix<double>& A,
              int order = 5, void* user_data = nullptr) {
    const bool is_constant = FuncTraits::IsConstant::value && FuncTraits::AggregateType::value == OnePointTensor;
    const int f = XYZ.fusion;
    afb_ctx* ctx = b200::default_context();
    b200::fem3Dtet_composite<FuncTraits>(applyOpU.c, applyOpV.c, XYZ, Dfnc, A, order, user_data,
        [&](int opA, int femA, int opB, int femB, const std::vector<double>& Dsub, std::vector<double>& Ablk) {
            afb_form fm{opA, femA, 1, opB, femB, 1, order, TENSOR_GENERAL, is_constant ? AFB_COEF_CONST : AFB_COEF_PER_POINT, AFB_HOST, Dsub.data(), 1.0, 0, 0};
            b200::check(ctx, afb_fem3dtet_batched(ctx, &fm, f, XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, Ablk.data(), AFB_HOST));
        });
}
}  // namespace Ani

#include "face_normal.hpp"   // fem3DfaceN: contraction of the tensor with the face normal (needs Tetras, eval_tensor_points)

namespace Ani {
/// int_f ((D OpA(u)) . N) . OpB(v) over face face_num of every tet (fem/operations/int_face.h:32-47, 97-133); A is nfB x (nfA*fusion).
/// The callback fills a col-major (3 Dim(OpB) x Dim(OpA)) tensor, normal component fastest.
template <typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3DfaceN(const Tetras<const double>& XYZ, int face_num, const ApplyOpBase& applyOpU, const ApplyOpBase& applyOpV, const Functor& Dfnc,
                DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    const int f = XYZ.fusion, nfa = static_cast<int>(applyOpU.Nfa()), nfb = static_cast<int>(applyOpV.Nfa());
    if (A.size < static_cast<std::size_t>(nfa) * nfb * f)
        throw std::runtime_error("Expected dimensions of A is " + std::to_string(nfb) + "x" + std::to_string(nfa * f) + ", but A has size = " + std::to_string(A.size));
    A.nRow = nfb; A.nCol = static_cast<std::size_t>(nfa) * f;
    afb_ctx* ctx = b200::default_context();
    b200::fem3DfaceN_contract<FuncTraits>(static_cast<int>(applyOpU.Dim()), static_cast<int>(applyOpV.Dim()), XYZ, face_num, Dfnc, order, user_data,
        [&](const std::vector<double>& DN, std::size_t per_tet) {
            // one record per tet (constant tensor: the normal still differs from tet to tet) or one per point of the triangle rule
            afb_form fm{applyOpU.op, applyOpU.fem, applyOpU.vec, applyOpV.op, applyOpV.fem, applyOpV.vec, order, TENSOR_GENERAL,
                        per_tet == 1 ? AFB_COEF_PER_TET : AFB_COEF_PER_POINT, AFB_HOST, DN.data(), 1.0, 0, 0};
            std::vector<int32_t> faces(f, face_num);
            b200::check(ctx, afb_fem3dface_batched(ctx, &fm, f, faces.data(), XYZ.XY0, XYZ.XY1, XYZ.XY2, XYZ.XY3, A.data, AFB_HOST));
        });
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3DfaceN(const Tetras<const double>& XYZ, int face_num, const Functor& Dfnc, DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    fem3DfaceN<FuncTraits>(XYZ, face_num, ApplyOpBase(OpA::op, OpA::fem, OpA::vec), ApplyOpBase(OpB::op, OpB::fem, OpB::vec), Dfnc, A, order, user_data);
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor>
void fem3DfaceN(const DenseMatrix<double>& XY0, const DenseMatrix<double>& XY1, const DenseMatrix<double>& XY2, const DenseMatrix<double>& XY3,
                int face_num, const Functor& Dfnc, DenseMatrix<double>& A, int order = 5, void* user_data = nullptr) {
    fem3DfaceN<OpA, OpB, FuncTraits>(make_tetras(XY0.data, XY1.data, XY2.data, XY3.data, static_cast<int>(XY0.nCol)), face_num, Dfnc, A, order, user_data);
}
template <typename OpA, typename OpB, typename FuncTraits = DfuncTraits<>, typename Functor, typename ScalarType, typename IndexType>
void fem3DfaceN(const Tetras<const double>& XYZ, int face_num, const Functor& Dfnc, DenseMatrix<double>& A, PlainMemory<ScalarType, IndexType>,
                int order = 5, void* user_data = nullptr) {
    fem3DfaceN<OpA, OpB, FuncTraits>(XYZ, face_num, Dfnc, A, order, user_data);
}
template <typename OpA, typename OpB, typename ScalarType = double, typename IndexType = int>
PlainMemory<ScalarType, IndexType> fem3DfaceN_memory_requirements(int /*order*/, int /*fusion*/ = 1) { return PlainMemory<ScalarType, IndexType>(); }
}  // namespace Ani

#include "dc_on_dof.hpp"   // applyDir / applyVectorDir helpers of a local assembler (needs DenseMatrix, ArrayView)
