// Composite element spaces (anifem_b200/composite.hpp: FemVecT, FemCom) on the CPU: the composition code of the product --
// flattening into scalar parts, sub-tensors per block, placement of the blocks -- runs with the REFERENCE build's scalar fem3Dtet
// as block evaluator (oracle/_ref/libanifem_ref.so, test infrastructure) and is compared with the reference's own FemCom / FemVecT
// operators on the same seeded tetrahedra and tensors.  In the product the block evaluator is afb_fem3dtet_batched.
#include <cmath>
#include <cstdio>
#include <random>

#include "anifem_b200/fem.hpp"

extern "C" {
int ref_fem3dtet(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D, long f, const double* XY0,
                 const double* XY1, const double* XY2, const double* XY3, double* A, int mode, int fuse, int nthreads);
int ref_fem3dtet_composite(int which, int order, int ttype, int layout, const double* D, long f, const double* XY0, const double* XY1, const double* XY2,
                           const double* XY3, double* A);
int ref_fem3dface(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D, long f, const int* face,
                  const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A);
int ref_fem3dfaceN(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D, long f, const int* face,
                   const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A);
const char* ref_last_error();
}

using namespace Ani;
static int fails = 0;
#ifdef GPU_FRONT_END
static const double TOL = 1e-12;   // GPU kernels against the reference build
#else
static const double TOL = 1e-13;   // same evaluator on both sides
#endif
#define EXPECT(c)                                                                  \
    do {                                                                           \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } \
    } while (0)

struct Data {   // tensor data in the user-callback layout, read back by the callbacks of both sides
    int ttype, layout, q;
    const double* D;
    std::size_t len;
    mutable long calls = 0;
};

template <typename OpA, typename OpB>
static void run_case(int which, int order, int ttype, int layout, unsigned seed) {
    const int f = 3;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> XY[4];
    for (int k = 0; k < 4; ++k) XY[k].resize(3 * f);
    for (int r = 0; r < f; ++r) {
        const double base[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int k = 0; k < 4; ++k) for (int d = 0; d < 3; ++d) XY[k][d + 3 * r] = base[k][d] + 0.2 * U(rng) + r;
    }
    const int q = afb_tet_quadrature(order, nullptr, nullptr, 0);
    const int dimA = OpA::Dim::value, dimB = OpB::Dim::value, nfa = OpA::Nfa::value, nfb = OpB::Nfa::value;
    const std::size_t len = ttype == TENSOR_NULL ? 0 : (ttype == TENSOR_SCALAR ? 1 : static_cast<std::size_t>(dimA) * dimB);
    const std::size_t nrec = layout == 0 ? 1 : (layout == 1 ? f : static_cast<std::size_t>(q) * f);
    std::vector<double> D(std::max<std::size_t>(1, len * nrec));
    for (auto& x : D) x = U(rng);
    if (ttype == TENSOR_SYMMETRIC)   // the promise of the flag: symmetric records (the reference reads them without transposition)
        for (std::size_t p = 0; p < nrec; ++p)
            for (int k = 0; k < dimB; ++k)
                for (int l = 0; l < k; ++l) D[len * p + l + static_cast<std::size_t>(dimB) * k] = D[len * p + k + static_cast<std::size_t>(dimB) * l];
    std::vector<double> want(static_cast<std::size_t>(nfa) * nfb * f, 0.0), got(want.size(), -7.0);
    if (ref_fem3dtet_composite(which, order, ttype, layout, D.data(), f, XY[0].data(), XY[1].data(), XY[2].data(), XY[3].data(), want.data()) != 0) {
        std::printf("reference composite case %d failed: %s\n", which, ref_last_error());
        ++fails;
        return;
    }
    Data dat{ttype, layout, q, D.data(), len};
    auto Dfnc = [&dat](const std::array<double, 3>&, double* Dm, TensorDims, void*, int iTet) {
        const long n = dat.layout == 2 ? dat.calls % dat.q : 0;
        dat.calls++;
        const double* src = dat.D;
        if (dat.layout == 1) src += dat.len * iTet;
        if (dat.layout == 2) src += dat.len * (n + static_cast<std::size_t>(dat.q) * iTet);
        for (std::size_t i = 0; i < dat.len; ++i) Dm[i] = src[i];
        return static_cast<TensorType>(dat.ttype);
    };
    DenseMatrix<> A(got.data(), nfb, static_cast<std::size_t>(nfa) * f);
    auto T = make_tetras(XY[0].data(), XY[1].data(), XY[2].data(), XY[3].data(), f);
    const b200::CompositeOp ca = b200::Describe<OpA>::get(), cb = b200::Describe<OpB>::get();
    EXPECT(ca.nfa == nfa && ca.dim == dimA && cb.nfa == nfb && cb.dim == dimB);
    int blocks = 0;
#ifdef GPU_FRONT_END
    // the product's own front end: the blocks are evaluated by afb_fem3dtet_batched on the GPU
    fem3Dtet<OpA, OpB, DfuncTraits<>>(T, Dfnc, A, order);
    blocks = -1;
    if (which == 0 && ttype == TENSOR_GENERAL) {   // runtime twins give the same matrix
        std::vector<double> rt(got.size(), -7.0);
        DenseMatrix<> Art(rt.data(), nfb, static_cast<std::size_t>(nfa) * f);
        const ComplexFemSpace UP = (FemSpace(FEM_P2) ^ 3) * FemSpace(FEM_P1);
        dat.calls = 0;
        fem3Dtet(T, UP.getOP(IDEN), UP.getOP(IDEN), Dfnc, Art, order);
        EXPECT(rt == got);
    }
#else
    b200::fem3Dtet_composite<DfuncTraits<>>(ca, cb, T, Dfnc, A, order, nullptr,
        [&](int opA, int femA, int opB, int femB, const std::vector<double>& Dsub, std::vector<double>& Ablk) {
            ++blocks;
            // per-point general sub-tensor, evaluated by the reference's scalar fem3Dtet (runtime operators)
            const int rc = ref_fem3dtet(opA, femA, 1, opB, femB, 1, order, TENSOR_GENERAL, 2, Dsub.data(), f, XY[0].data(), XY[1].data(), XY[2].data(),
                                        XY[3].data(), Ablk.data(), 0, 1, 1);
            if (rc) { std::printf("reference block failed: %s\n", ref_last_error()); ++fails; }
        });
#endif
    double scale = 0, err = 0;
    for (std::size_t k = 0; k < want.size(); ++k) { scale = std::fmax(scale, std::fabs(want[k])); err = std::fmax(err, std::fabs(want[k] - got[k])); }
    std::printf("case %d (order %d, tensor %d, layout %d): %d x %d, %d blocks evaluated of %zu, max |dA| / |A| = %.2e\n", which, order, ttype, layout, nfb, nfa,
                blocks, ca.parts.size() * cb.parts.size(), err / scale);
    EXPECT(scale > 0 && err <= TOL * scale);
}

// fem3DfaceN: the product's contraction of the tensor with the face normal (face_normal.hpp) + the reference's fem3Dface as
// evaluator, against the reference's own fem3DfaceN
static void run_faceN(int opA, int femA, int opB, int femB, int vecB, int order, int ttype, bool constant, unsigned seed) {
    const int f = 3;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> XY[4];
    for (int k = 0; k < 4; ++k) XY[k].resize(3 * f);
    for (int r = 0; r < f; ++r) {
        const double base[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int k = 0; k < 4; ++k) for (int d = 0; d < 3; ++d) XY[k][d + 3 * r] = base[k][d] + 0.25 * U(rng) - r;
    }
    if (seed & 1) for (int d = 0; d < 3; ++d) std::swap(XY[2][d], XY[3][d]);   // a negatively oriented tet among them
    const ApplyOpBase oa(opA, femA, 1), ob(opB, femB, vecB);
    const int dimA = static_cast<int>(oa.Dim()), dimB = static_cast<int>(ob.Dim()), nfa = static_cast<int>(oa.Nfa()), nfb = static_cast<int>(ob.Nfa());
    const int q = afb_tri_quadrature(order, nullptr, nullptr, 0);
    const std::size_t len = ttype == TENSOR_SCALAR ? 1 : static_cast<std::size_t>(3) * dimB * dimA;
    const int layout = constant ? 0 : 2;
    std::vector<double> D(len * (constant ? 1 : static_cast<std::size_t>(q) * f));
    for (auto& x : D) x = U(rng);
    for (int face = 0; face < 4; ++face) {
        std::vector<int> faces(f, face);
        std::vector<double> want(static_cast<std::size_t>(nfa) * nfb * f, 0.0), got(want.size(), -7.0);
        if (ref_fem3dfaceN(opA, femA, 1, opB, femB, vecB, order, ttype, layout, D.data(), f, faces.data(), XY[0].data(), XY[1].data(), XY[2].data(), XY[3].data(),
                           want.data()) != 0) { std::printf("reference fem3DfaceN failed: %s\n", ref_last_error()); ++fails; return; }
        Data dat{ttype, layout, q, D.data(), len};
        auto Dfnc = [&dat](const std::array<double, 3>&, double* Dm, TensorDims, void*, int iTet) {
            const long n = dat.layout == 2 ? dat.calls % dat.q : 0;
            dat.calls++;
            const double* src = dat.D + (dat.layout == 2 ? dat.len * (n + static_cast<std::size_t>(dat.q) * iTet) : 0);
            for (std::size_t i = 0; i < dat.len; ++i) Dm[i] = src[i];
            return static_cast<TensorType>(dat.ttype);
        };
        auto T = make_tetras(XY[0].data(), XY[1].data(), XY[2].data(), XY[3].data(), f);
        auto eval = [&](const std::vector<double>& DN, std::size_t per_tet) {
            const int rc = ref_fem3dface(opA, femA, 1, opB, femB, vecB, order, TENSOR_GENERAL, per_tet == 1 ? 1 : 2, DN.data(), f, faces.data(), XY[0].data(),
                                         XY[1].data(), XY[2].data(), XY[3].data(), got.data());
            if (rc) { std::printf("reference fem3Dface failed: %s\n", ref_last_error()); ++fails; }
        };
#ifdef GPU_FRONT_END
        (void)eval;
        DenseMatrix<> Ag(got.data(), nfb, static_cast<std::size_t>(nfa) * f);
        if (constant) fem3DfaceN<DfuncTraits<PerPoint, true>>(T, face, oa, ob, Dfnc, Ag, order);
        else fem3DfaceN<DfuncTraits<>>(T, face, oa, ob, Dfnc, Ag, order);
#else
        if (constant) b200::fem3DfaceN_contract<DfuncTraits<PerPoint, true>>(dimA, dimB, T, face, Dfnc, order, nullptr, eval);
        else b200::fem3DfaceN_contract<DfuncTraits<>>(dimA, dimB, T, face, Dfnc, order, nullptr, eval);
#endif
        double scale = 0, err = 0;
        for (std::size_t k = 0; k < want.size(); ++k) { scale = std::fmax(scale, std::fabs(want[k])); err = std::fmax(err, std::fabs(want[k] - got[k])); }
        if (!(scale > 0 && err <= TOL * scale))
            std::printf("fem3DfaceN op %d fem %d -> op %d fem %d vec %d, face %d: max |dA| / |A| = %.2e\n", opA, femA, opB, femB, vecB, face, err / scale);
        EXPECT(scale > 0 && err <= TOL * scale);
    }
}

int main() {
    run_faceN(GRAD, FEM_P2, IDEN, FEM_P2, 1, 4, TENSOR_GENERAL, false, 11);    // (K grad u) . N v, K per point
    run_faceN(GRAD, FEM_P1, IDEN, FEM_P1, 1, 2, TENSOR_SCALAR, false, 12);     // s du/dn v
    run_faceN(IDEN, FEM_P1, IDEN, FEM_P2, 1, 3, TENSOR_GENERAL, true, 13);     // (b u) . N v, constant b (3 x 1): per-tet records after the contraction
    run_faceN(GRAD, FEM_P1, IDEN, FEM_P1, 3, 3, TENSOR_GENERAL, false, 14);    // vector test space: 9 x 3 tensor
    run_faceN(GRAD, FEM_P3, IDEN, FEM_P0, 1, 5, TENSOR_GENERAL, false, 15);
    std::printf("fem3DfaceN: 5 operator pairs x 4 faces done\n");
    using Stokes = FemCom<FemVec<3, FEM_P2>, FemFix<FEM_P1>>;
    using P1x2 = FemVecT<2, FemFix<FEM_P1>>;
    using P1P1 = FemCom<FemFix<FEM_P1>, FemFix<FEM_P1>>;
    using P2P0 = FemCom<FemFix<FEM_P2>, FemFix<FEM_P0>>;
    using P1x3 = FemVecT<3, FemFix<FEM_P1>>;
    using Mixed = FemCom<FemVecT<2, FemFix<FEM_P2>>, FemFix<FEM_P1>>;
    using Tri = FemCom<FemFix<FEM_P1>, FemFix<FEM_P1>, FemFix<FEM_P0>>;
    static_assert(Operator<IDEN, Stokes>::Nfa::value == 34 && Operator<IDEN, Stokes>::Dim::value == 4, "Taylor-Hood sizes");
    static_assert(Operator<GRAD, Mixed>::Nfa::value == 24 && Operator<GRAD, Mixed>::Dim::value == 9, "mixed sizes");
    // per-point layout everywhere: the composition hands per-point sub-tensors to the block evaluator
    run_case<Operator<IDEN, Stokes>, Operator<IDEN, Stokes>>(0, 4, TENSOR_GENERAL, 2, 1);
    run_case<Operator<IDEN, Stokes>, Operator<IDEN, Stokes>>(0, 3, TENSOR_SCALAR, 2, 2);    // block-diagonal: 4 of 16 blocks
    run_case<Operator<GRAD, P1x2>, Operator<GRAD, P1P1>>(1, 2, TENSOR_GENERAL, 2, 3);
    run_case<Operator<IDEN, P2P0>, Operator<GRAD, FemFix<FEM_P1>>>(2, 3, TENSOR_GENERAL, 2, 4);
    run_case<Operator<IDEN, P1x3>, Operator<IDEN, P1x3>>(3, 2, TENSOR_NULL, 2, 5);
    run_case<Operator<IDEN, P1x3>, Operator<IDEN, P1x3>>(3, 2, TENSOR_SYMMETRIC, 2, 6);
    run_case<Operator<GRAD, Mixed>, Operator<IDEN, Tri>>(4, 3, TENSOR_GENERAL, 2, 7);
    {   // runtime twins: (P2 ^ 3) * P1 describes the same operator as the template FemCom<FemVec<3,P2>, FemFix<P1>>
        const ComplexFemSpace UP = (FemSpace(FEM_P2) ^ 3) * FemSpace(FEM_P1);
        const b200::CompositeOp a = UP.getOP(GRAD).c, b = b200::Describe<Operator<GRAD, Stokes>>::get();
        EXPECT(UP.dofMapSize() == 34 && a.nfa == b.nfa && a.dim == b.dim && a.parts.size() == b.parts.size());
        for (std::size_t k = 0; k < a.parts.size() && k < b.parts.size(); ++k)
            EXPECT(a.parts[k].fem == b.parts[k].fem && a.parts[k].nfa_off == b.parts[k].nfa_off && a.parts[k].comp == b.parts[k].comp);
        EXPECT(UP.dofMap().NumDofOnTet() == 34 && UP.dofMap() == FemSpace(FEM_P2, 3).dofMap() * FemSpace(FEM_P1).dofMap());
        const ComplexFemSpace fused = FemSpace(FEM_P1) * FemSpace(FEM_P1) * FemSpace(FEM_P0);   // P1 * P1 fuses into P1 ^ 2
        EXPECT(fused.parts.size() == 2 && fused.parts[0].vec == 2 && fused.getOP(IDEN).Nfa() == 9 && fused.getOP(IDEN).Dim() == 3);
        bool thrown = false;
        try { UP.getOP(DIV); } catch (std::runtime_error&) { thrown = true; }
        EXPECT(thrown);
    }
    if (fails) { std::printf("test_composite: %d FAILED\n", fails); return 1; }
    std::printf("test_composite: all passed\n");
    return 0;
}
