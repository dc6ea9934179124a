// GPU test of the C++ mirror of the reference API (anifem_b200/fem.hpp, assembler.hpp).  Reads like the reference's own
// tests/fem/operations/int_tet_test.cpp: the same calls, the same golden tables.  Exit code 0 = pass.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <functional>

#include "anifem_b200/assembler.hpp"

using namespace Ani;

static double norm_diff(double a, const DenseMatrix<>& A, double b, const double* B) {
    double s = 0;
    for (std::size_t i = 0; i < A.nRow * A.nCol; ++i) { double d = a * A.data[i] + b * B[i]; s += d * d; }
    return std::sqrt(s);
}
static double norm(const DenseMatrix<>& A) { double s = 0; for (std::size_t i = 0; i < A.nRow * A.nCol; ++i) s += A.data[i] * A.data[i]; return std::sqrt(s); }

#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

// compile check of the composite-space front end (FemVecT / FemCom, composite.hpp); the composition itself is verified on the CPU
// by tests/cxx/test_composite.cpp -- this function is not called
void compile_check_composite(const Tetras<const double>& T, DenseMatrix<double>& A) {
    using Stokes = FemCom<FemVec<3, FEM_P2>, FemFix<FEM_P1>>;
    auto D = [](const std::array<double, 3>&, double* Dm, TensorDims d, void*, int) { for (std::size_t i = 0; i < d.first * d.second; ++i) Dm[i] = 0; return TENSOR_GENERAL; };
    fem3Dtet<Operator<IDEN, Stokes>, Operator<IDEN, Stokes>>(T, D, A, 3);
    fem3Dtet<Operator<GRAD, FemVecT<2, FemFix<FEM_P1>>>, Operator<GRAD, FemFix<FEM_P1>>, DfuncTraits<TENSOR_GENERAL, true>>(T, D, A, 2);
    const ComplexFemSpace UP = (FemSpace(FEM_P2) ^ 3) * FemSpace(FEM_P1);
    fem3Dtet(T, UP.getOP(IDEN), UP.getOP(IDEN), D, A, 3);
    fem3DfaceN<Operator<GRAD, FemFix<FEM_P2>>, Operator<IDEN, FemFix<FEM_P2>>>(T, 1, D, A, 4);
}

int main() {
    int fails = 0;
    double XY1p[] = {0, 0, 0}, XY2p[] = {2, 1, 1}, XY3p[] = {1, 2, 1}, XY4p[] = {2, 1, 2};
    DenseMatrix<> XY1(XY1p, 3, 1), XY2(XY2p, 3, 1), XY3(XY3p, 3, 1), XY4(XY4p, 3, 1);
    // --- GRAD(P1^3) x GRAD(P1^3), identity tensor: 12x12 table /1440 (int_tet_test.cpp:332-347) for trait variants
    {
        double A_exp[144] = {0};
        const double blk[16] = {160, -240, -80, 160, -240, 1440, -240, -960, -80, -240, 400, -80, 160, -960, -80, 880};
        for (int c = 0; c < 3; ++c) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) A_exp[(4 * c + i) + 12 * (4 * c + j)] = blk[4 * i + j] / 1440.0;
        double Ad[144];
        DenseMatrix<> A(Ad, 12, 12);
        using Op = Operator<GRAD, FemVec<3, FEM_P1>>;
        auto make_const = [](TensorType t) {
            return [t](const std::array<double, 3>&, double* Dmem, TensorDims d, void*, int) {
                for (std::size_t i = 0; i < d.first; ++i) for (std::size_t j = 0; j < d.second; ++j) Dmem[i + d.first * j] = (i == j);
                return t;
            };
        };
        fem3Dtet<Op, Op, DfuncTraits<TENSOR_GENERAL, true>>(XY1, XY2, XY3, XY4, make_const(TENSOR_GENERAL), A, 5);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
        fem3Dtet<Op, Op, DfuncTraits<PerPoint, false>>(XY1, XY2, XY3, XY4, make_const(TENSOR_SYMMETRIC), A, 5);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
        fem3Dtet<Op, Op, DfuncTraits<TENSOR_SCALAR, false>>(XY1, XY2, XY3, XY4, make_const(TENSOR_SCALAR), A, 5);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
        fem3Dtet<Op, Op, DfuncTraits<TENSOR_NULL, true>>(XY1, XY2, XY3, XY4, make_const(TENSOR_NULL), A, 5);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
    }
    // --- rhs trick: int (mu,mu,mu).phi_i with OpA = IDEN(P0) on P2^3 (int_tet_test.cpp:448-501)
    {
        double Bp[30] = {-1, -1, -1, -1, 4, 4, 4, 4, 4, 4, -1, -1, -1, -1, 4, 4, 4, 4, 4, 4, -1, -1, -1, -1, 4, 4, 4, 4, 4, 4};
        double Ap[30];
        DenseMatrix<> A(Ap, 30, 1);
        const double mu = 40;
        auto scal = [mu](const std::array<double, 3>&, double* Dmem, TensorDims d, void*, int) {
            for (std::size_t i = 0; i < d.first * d.second; ++i) Dmem[i] = mu;
            return TENSOR_SCALAR;
        };
        fem3Dtet<Operator<IDEN, FemFix<FEM_P0>>, Operator<IDEN, FemVec<3, FEM_P2>>, DfuncTraits<PerPoint, false>>(XY1, XY2, XY3, XY4, scal, A, 2);
        EXPECT(norm_diff(1, A, -1, Bp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
    }
    // --- reference overloads with caller memory and the runtime twin (dyn_ops.h:18-23, int_tet.h:111-160): same numbers
    {
        using Op = Operator<GRAD, FemFix<FEM_P2>>;
        auto Kfn = [](const std::array<double, 3>& x, double* Dmem, TensorDims d, void*, int) {
            for (std::size_t i = 0; i < d.first; ++i) for (std::size_t j = 0; j < d.second; ++j) Dmem[i + d.first * j] = (i == j) * (1 + x[0]) + 0.1 * (i + j);
            return TENSOR_SYMMETRIC;
        };
        double A0[100], A1[100], A2[100], A3[100];
        DenseMatrix<> M0(A0, 10, 10), M1(A1, 10, 10), M2(A2, 10, 10), M3(A3, 10, 10);
        fem3Dtet<Op, Op>(XY1, XY2, XY3, XY4, Kfn, M0, 4);
        auto req = fem3Dtet_memory_requirements<Op, Op>(4, 1);
        EXPECT(req.dSize == 0 && req.iSize == 0);
        fem3Dtet<Op, Op>(XY1, XY2, XY3, XY4, Kfn, M1, req, 4);
        DynMem<> wmem;
        fem3Dtet<Op, Op>(make_tetras(XY1p, XY2p, XY3p, XY4p, 1), Kfn, M2, wmem, 4);
        FemSpace P2s{FEM_P2};
        fem3Dtet(make_tetras(XY1p, XY2p, XY3p, XY4p, 1), P2s.getOP(GRAD), P2s.getOP(GRAD), Kfn, M3, wmem, 4);
        EXPECT(P2s.getOP(GRAD).Nfa() == 10 && P2s.getOP(GRAD).Dim() == 3 && (P2s ^ 3).getOP(GRAD).Dim() == 9 && (P2s ^ 3).dofMapSize() == 30);
        double d = 0;
        for (int i = 0; i < 100; ++i) d = std::fmax(d, std::fmax(std::fabs(A1[i] - A0[i]), std::fmax(std::fabs(A2[i] - A0[i]), std::fabs(A3[i] - A0[i]))));
        EXPECT(d == 0.0 && norm(M0) > 0);
        bool thrown = false;
        try { ApplyOpBase bad(DIV, FEM_P1, 1); (void)bad; } catch (std::runtime_error&) { thrown = true; }
        EXPECT(thrown);
    }
    // --- fusive mass matrix with D = x^2 per point (int_tet_test.cpp:395-446): fusion of 2 equals two single calls
    {
        double X0[6] = {0, 0, 0, 1, 1, 1}, X1[6] = {2, 1, 1, 2, 1, 1}, X2[6] = {1, 2, 1, 1, 2, 1}, X3[6] = {2, 1, 2, 1, 1, 2};
        auto xsq = [](const std::array<double, 3>& x, double* Dmem, TensorDims, void*, int) { Dmem[0] = x[0] * x[0]; return TENSOR_SCALAR; };
        using Op = Operator<IDEN, FemFix<FEM_P1>>;
        double A2[32], A1[32];
        DenseMatrix<> Af(A2, 4, 8), As(A1, 4, 4);
        fem3Dtet<Op, Op>(make_tetras(X0, X1, X2, X3, 2), xsq, Af, 4);
        for (int r = 0; r < 2; ++r) {
            DenseMatrix<> Ar(A1 + 16 * r, 4, 4);
            fem3Dtet<Op, Op>(make_tetras(X0 + 3 * r, X1 + 3 * r, X2 + 3 * r, X3 + 3 * r, 1), xsq, Ar, 4);
        }
        (void)As;
        double d = 0;
        for (int i = 0; i < 32; ++i) d = std::fmax(d, std::fabs(A2[i] - A1[i]));
        EXPECT(d <= 1e-15);
    }
    // --- error behaviour: incompatible identity tensor throws (diff_tensor.h:315-317); small A throws (int_tet.inl:8-9)
    {
        double Ad[16];
        DenseMatrix<> A(Ad, 4, 4);
        auto nul = [](const std::array<double, 3>&, double*, TensorDims, void*, int) { return TENSOR_NULL; };
        bool thrown = false;
        try { fem3Dtet<Operator<GRAD, FemFix<FEM_P1>>, Operator<IDEN, FemFix<FEM_P1>>, DfuncTraits<TENSOR_NULL, true>>(XY1, XY2, XY3, XY4, nul, A, 2); }
        catch (std::runtime_error&) { thrown = true; }
        EXPECT(thrown);
        DenseMatrix<> As(Ad, 2, 2);
        thrown = false;
        try { fem3Dtet<Operator<IDEN, FemFix<FEM_P1>>, Operator<IDEN, FemFix<FEM_P1>>>(XY1, XY2, XY3, XY4, nul, As, 2); }
        catch (std::runtime_error&) { thrown = true; }
        EXPECT(thrown);
    }
    // --- Assembler: P1 Poisson + mass + load on a 4^3 cube (the ex1.cpp / Fem/Ani/diffusion.cpp flow)
    {
        Assembler discr;
        bool thrown = false;
        try { discr.PrepareProblem(); } catch (std::runtime_error&) { thrown = true; }   // "Mesh was not specified"
        EXPECT(thrown);
        discr.SetCubeMesh(4, 4, 4).SetProbDescr({{FEM_P1, 1}});
        const double K[9] = {1, -1, 0, -1, 1, 0, 0, 0, 1}, one = 1.0;
        using G1 = Operator<GRAD, FemFix<FEM_P1>>; using I1 = Operator<IDEN, FemFix<FEM_P1>>;
        discr.AddMatForm<G1, G1>(0, 0, 2, TENSOR_SYMMETRIC, AFB_COEF_CONST, K).AddMatForm<I1, I1>(0, 0, 2, TENSOR_SCALAR, AFB_COEF_CONST, &one);
        discr.AddRhsForm<I1>(0, 2, TENSOR_SCALAR, AFB_COEF_CONST, &one);
        discr.PrepareProblem();
        EXPECT(discr.getBegInd() == 0 && discr.getEndInd() == 125);
        CsrMatrix A; std::vector<double> b;
        EXPECT(discr.Assemble(A, b) == 0);
        double sa = 0, sb = 0;
        for (double v : A.val) sa += v;
        for (double v : b) sb += v;
        EXPECT(std::fabs(sa - 1.0) < 1e-12);   // 1^T K_stiff 1 = 0, 1^T M 1 = |Omega| = 1
        EXPECT(std::fabs(sb - 1.0) < 1e-12);
        for (int64_t r = 0; r + 1 < (int64_t)A.rowptr.size(); ++r) {
            bool diag = false;
            for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) {
                if (k > A.rowptr[r]) EXPECT(A.colind[k] > A.colind[k - 1]);
                diag = diag || A.colind[k] == r;
            }
            EXPECT(diag);
        }
        std::vector<double> v1 = A.val;
        EXPECT(discr.Assemble(A, b) == 0);   // Assemble accumulates (assembler.inl:305-306)
        {   // Dirichlet dof 0 with value 2 (applyDir, dc_on_dof.h:27-45): row 0 = deg * e_0, rhs_0 = deg * 2, column 0 empty elsewhere
            std::vector<unsigned char> flag(125, 0); std::vector<double> bc(125, 0.0);
            flag[0] = 1; bc[0] = 2.0;
            discr.SetDirichlet(flag, bc);
            CsrMatrix Ad; std::vector<double> bd;
            EXPECT(discr.Assemble(Ad, bd) == 0);
            double diag = 0, off = 0, col0 = 0;
            for (int64_t k = Ad.rowptr[0]; k < Ad.rowptr[1]; ++k) (Ad.colind[k] == 0 ? diag : off) += std::fabs(Ad.val[k]);
            for (int64_t r = 1; r + 1 < (int64_t)Ad.rowptr.size(); ++r)
                for (int64_t k = Ad.rowptr[r]; k < Ad.rowptr[r + 1]; ++k) if (Ad.colind[k] == 0) col0 += std::fabs(Ad.val[k]);
            EXPECT(diag >= 1.0 && off == 0.0 && col0 == 0.0 && std::fabs(bd[0] - 2.0 * diag) < 1e-13);
            discr.SetDirichlet({}, {});
        }
        double d = 0;
        for (std::size_t k = 0; k < v1.size(); ++k) d = std::fmax(d, std::fabs(A.val[k] - 2 * v1[k]));
        EXPECT(d == 0.0);
    }
    // --- Assembler with surface terms: Robin matrix (alpha = 1) + Neumann load (g = 2) on the whole boundary of the unit cube, P2:
    //     1^T A_face 1 = |boundary| = 6, sum of the face load = 2 * 6 (the fem3Dface calls of Fem/Ani/diffusion.cpp:215-245)
    {
        Assembler discr;
        discr.SetCubeMesh(3, 3, 3).SetProbDescr({{FEM_P2, 1}});
        using G2 = Operator<GRAD, FemFix<FEM_P2>>; using I2 = Operator<IDEN, FemFix<FEM_P2>>;
        const double one = 1.0, two = 2.0;
        discr.AddMatForm<G2, G2>(0, 0, 2, TENSOR_NULL, AFB_COEF_CONST, nullptr);
        discr.PrepareProblem();
        int64_t nnode = 0, ntet = 0;
        EXPECT(afb_mesh_get(discr.context(), &nnode, &ntet, nullptr, nullptr, AFB_HOST) == 0);
        std::vector<double> xyz(3 * nnode); std::vector<int32_t> v(4 * ntet);
        EXPECT(afb_mesh_get(discr.context(), &nnode, &ntet, xyz.data(), v.data(), AFB_HOST) == 0);
        std::vector<int32_t> fc, fn;
        auto on_side = [&](int node, int side) { const double c = xyz[(side % 3) * nnode + node]; return std::fabs(c - (side < 3 ? 0.0 : 1.0)) < 1e-12; };
        for (int64_t e = 0; e < ntet; ++e)
            for (int k = 0; k < 4; ++k)
                for (int side = 0; side < 6; ++side)
                    if (on_side(v[((k + 0) % 4) * ntet + e], side) && on_side(v[((k + 1) % 4) * ntet + e], side) && on_side(v[((k + 2) % 4) * ntet + e], side)) {
                        fc.push_back((int32_t)e); fn.push_back(k);
                    }
        EXPECT(fc.size() == 6 * 9 * 2);
        CsrMatrix A0; std::vector<double> b0;
        EXPECT(discr.Assemble(A0, b0) == 0);
        discr.SetBoundaryFaces(fc, fn);
        discr.AddFaceMatForm<I2, I2>(0, 0, 4, TENSOR_SCALAR, AFB_COEF_CONST, &one).AddFaceRhsForm<I2>(0, 4, TENSOR_SCALAR, AFB_COEF_CONST, &two);
        CsrMatrix A; std::vector<double> b;
        EXPECT(discr.Assemble(A, b) == 0);
        double sa = 0, sb = 0;
        for (std::size_t k = 0; k < A.val.size(); ++k) sa += A.val[k] - A0.val[k];
        for (std::size_t k = 0; k < b.size(); ++k) sb += b[k] - b0[k];
        EXPECT(std::fabs(sa - 6.0) < 1e-11);
        EXPECT(std::fabs(sb - 12.0) < 1e-11);
    }
    // --- fem3Dface: GRAD(P2) x IDEN(P1^3) over face 1, normal-weighted tensor, 12x10 table of tests/fem/operations/int_face_test.cpp:29-95
    {
        double P1[] = {1, 1, 1}, P2[] = {2, 1, 1}, P3[] = {1, 2, 1}, P4[] = {1, 1, 2};
        DenseMatrix<> F1(P1, 3, 1), F2(P2, 3, 1), F3(P3, 3, 1), F4(P4, 3, 1);
        using OP1 = Operator<GRAD, FemFix<FEM_P2>>;
        using OP2 = Operator<IDEN, FemVec<3, FEM_P1>>;
        const double s3 = 1.0 / std::sqrt(3.0);   // outward normal of face 1 = (1,1,1)/sqrt(3)
        auto ddotn = [s3](const std::array<double, 3>& x, double* Dmem, TensorDims dims, void*, int) {
            const int d1 = 3, d2 = 3;
            if ((int)dims.first != d1 || (int)dims.second != d2) throw std::runtime_error("Error in expected tensor sizes");
            for (int j = 0; j < d2; ++j)
                for (int i = 0; i < d1; ++i) {
                    double v = 0;
                    for (int k = 0; k < 3; ++k) v += (x[j % 3] + (3 * i + k) / 10.0) * s3;
                    Dmem[i + d1 * j] = v;
                }
            return TENSOR_GENERAL;
        };
        const double A_exp[120] = {
            0.000, 2.150, 2.150, 2.150, 0.000, 2.600, 2.600, 2.600, 0.000, 3.050, 3.050, 3.050,
            0.000, 0.900, 0.075, 0.075, 0.000, 1.050, 0.075, 0.075, 0.000, 1.200, 0.075, 0.075,
            0.000, 0.075, 0.900, 0.075, 0.000, 0.075, 1.050, 0.075, 0.000, 0.075, 1.200, 0.075,
            0.000, 0.075, 0.075, 0.900, 0.000, 0.075, 0.075, 1.050, 0.000, 0.075, 0.075, 1.200,
            0.000,-4.300,-2.150,-2.150, 0.000,-5.200,-2.600,-2.600, 0.000,-6.100,-3.050,-3.050,
            0.000,-2.150,-4.300,-2.150, 0.000,-2.600,-5.200,-2.600, 0.000,-3.050,-6.100,-3.050,
            0.000,-2.150,-2.150,-4.300, 0.000,-2.600,-2.600,-5.200, 0.000,-3.050,-3.050,-6.100,
            0.000, 2.050, 2.050, 1.300, 0.000, 2.500, 2.500, 1.600, 0.000, 2.950, 2.950, 1.900,
            0.000, 2.050, 1.300, 2.050, 0.000, 2.500, 1.600, 2.500, 0.000, 2.950, 1.900, 2.950,
            0.000, 1.300, 2.050, 2.050, 0.000, 1.600, 2.500, 2.500, 0.000, 1.900, 2.950, 2.950};
        double Ad[120];
        DenseMatrix<> A(Ad, 12, 10);
        fem3Dface<OP1, OP2>(F1, F2, F3, F4, 1, ddotn, A, 3);
        EXPECT(A.nRow == 12 && A.nCol == 10);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
        fem3Dface<DfuncTraits<>>(make_tetras(P1, P2, P3, P4, 1), 1, FemSpace(FEM_P2).getOP(GRAD), FemSpace(FEM_P1, 3).getOP(IDEN), ddotn, A, PlainMemoryX<>(), 3);
        EXPECT(norm_diff(1, A, -1, A_exp) <= 100 * (1 + norm(A)) * DBL_EPSILON);
        bool thrown = false;
        try { fem3Dface<OP1, OP2>(F1, F2, F3, F4, 4, ddotn, A, 3); } catch (const std::runtime_error&) { thrown = true; }
        EXPECT(thrown);
    }
    // --- DfuncTraitsFusive (one callback invocation fills the tensors of all points, diff_tensor.h:29-53,101-135) and PerSelection
    //     (one tensor type per call) give the same matrices as the per-point callback
    {
        using OP = Operator<GRAD, FemFix<FEM_P2>>;
        double XY[4][6] = {{0, 0, 0, 1, 1, 1}, {1, 0.1, 0, 2.2, 1, 1}, {0, 1.3, 0.2, 1, 2, 1.5}, {0.1, 0.2, 0.9, 1, 1.1, 2.5}};
        const int f = 2, order = 3;
        auto Kpt = [](const std::array<double, 3>& X, double* D, TensorDims dims, void*, int) {
            for (std::size_t i = 0; i < dims.first * dims.second; ++i) D[i] = 0;
            for (int k = 0; k < 3; ++k) D[k + 3 * k] = 1 + X[0] + 0.5 * k;
            D[0 + 3 * 1] = D[1 + 3 * 0] = 0.25 * X[1];
            return TENSOR_SYMMETRIC;
        };
        int calls = 0;
        auto Kfus = [&](ArrayView<double> X, ArrayView<double> D, TensorDims dims, void*, const AniMemory<double, int>& mem) {
            ++calls;
            const std::size_t dl = dims.first * dims.second;
            EXPECT(X.size == 3 * mem.q * mem.f && D.size == dl * mem.q * mem.f && mem.XYG.data == X.data && mem.WG.size == mem.q);
            for (std::size_t p = 0; p < mem.q * mem.f; ++p) {
                double* Dp = D.data + dl * p;
                for (std::size_t i = 0; i < dl; ++i) Dp[i] = 0;
                for (int k = 0; k < 3; ++k) Dp[k + 3 * k] = 1 + X[3 * p] + 0.5 * k;
                Dp[0 + 3 * 1] = Dp[1 + 3 * 0] = 0.25 * X[3 * p + 1];
            }
            return TENSOR_SYMMETRIC;
        };
        std::vector<double> a1(100 * f), a2(100 * f), a3(100 * f);
        DenseMatrix<> A1(a1.data(), 10, 10 * f), A2(a2.data(), 10, 10 * f), A3(a3.data(), 10, 10 * f);
        auto T = make_tetras(XY[0], XY[1], XY[2], XY[3], f);
        fem3Dtet<OP, OP, DfuncTraits<>>(T, Kpt, A1, order);
        fem3Dtet<OP, OP, DfuncTraitsFusive<>>(T, Kfus, A2, order);
        fem3Dtet<OP, OP, DfuncTraits<PerSelection>>(T, Kpt, A3, order);
        EXPECT(calls == 1);
        EXPECT(a1 == a2 && a1 == a3);
    }
    // --- the MatFuncWrap plug-in point (func_wrap.h:310-347, assembler.h:326-328): the local assembler lambda of
    //     examples/tutorials/ex1.cpp:83-106 unchanged (P1, D(x) = 1 + x^2 per quadrature point, F = 1, Dirichlet u = 0 on the
    //     boundary through applyDir), installed with SetMatRHSFunc(GenerateElemMatRhs(...)) + a data gatherer; the result equals
    //     the declarative description (AddMatForm with the coefficient at the quadrature points + SetDirichlet)
    {
        using UFem = FemFix<FEM_P1>;
        constexpr int UNF = 4;
        const int order = 2;
        auto D_tensor = [](const std::array<double, 3>& X, double* D, TensorDims, void*, int) { D[0] = 1 + X[0] * X[0]; return TENSOR_SCALAR; };
        auto F_tensor = [](const std::array<double, 3>&, double* F, TensorDims, void*, int) { F[0] = 1; return TENSOR_SCALAR; };
        struct ProbLocData { std::array<int, 4> nlbl = {{0, 0, 0, 0}}; };
        std::function<void(const double**, double*, double*, void*)> local_assembler =
            [&D_tensor, &F_tensor, order](const double** XY, double* Adat, double* Fdat, void* user_data) -> void {
            DenseMatrix<> A(Adat, UNF, UNF), F(Fdat, UNF, 1);
            A.SetZero(); F.SetZero();
            DenseMatrix<> X0(const_cast<double*>(XY[0]), 3, 1), X1(const_cast<double*>(XY[1]), 3, 1), X2(const_cast<double*>(XY[2]), 3, 1), X3(const_cast<double*>(XY[3]), 3, 1);
            fem3Dtet<Operator<GRAD, UFem>, Operator<GRAD, UFem>, DfuncTraits<TENSOR_SCALAR, false>>(X0, X1, X2, X3, D_tensor, A, order);
            fem3Dtet<Operator<IDEN, FemFix<FEM_P0>>, Operator<IDEN, UFem>, DfuncTraits<TENSOR_SCALAR, true>>(X0, X1, X2, X3, F_tensor, F, order);
            auto& dat = *static_cast<ProbLocData*>(user_data);
            for (int i = 0; i < 4; ++i)
                if (dat.nlbl[i] > 0) applyDir(A, F, i, 0.0);
        };
        Assembler discr;
        discr.SetCubeMesh(3, 3, 2).SetProbDescr({{FEM_P1, 1}});
        discr.PrepareProblem();
        int64_t nnode = 0, ntet = 0;
        EXPECT(afb_mesh_get(discr.context(), &nnode, &ntet, nullptr, nullptr, AFB_HOST) == 0);
        std::vector<double> xyz(3 * nnode); std::vector<int32_t> v(4 * ntet);
        EXPECT(afb_mesh_get(discr.context(), &nnode, &ntet, xyz.data(), v.data(), AFB_HOST) == 0);
        std::vector<int> label(nnode, 0);
        for (int64_t n = 0; n < nnode; ++n)
            for (int k = 0; k < 3; ++k) {
                if (std::fabs(xyz[k * nnode + n] - 0) < 1e-12) label[n] |= 1 << (2 * k);
                if (std::fabs(xyz[k * nnode + n] - 1) < 1e-12) label[n] |= 1 << (2 * k + 1);
            }
        auto local_data_gatherer = [&label](ElementalAssembler& p) -> void {
            double* nn_p = p.get_nodes();
            const double* args[] = {nn_p, nn_p + 3, nn_p + 6, nn_p + 9};
            ProbLocData data;
            for (unsigned i = 0; i < data.nlbl.size(); ++i) data.nlbl[i] = label[p.node_ids[i]];
            p.compute(args, &data);
        };
        discr.SetMatRHSFunc(GenerateElemMatRhs(local_assembler, UNF, UNF));
        discr.SetDataGatherer(local_data_gatherer);
        CsrMatrix A1, A4; std::vector<double> b1, b4;
        discr.m_assm_traits.num_threads = 1;
        EXPECT(discr.Assemble(A1, b1) == 0);
        discr.m_assm_traits.num_threads = 4;
        EXPECT(discr.Assemble(A4, b4) == 0);
        EXPECT(A1.val == A4.val && b1 == b4);   // the scatter is deterministic whatever the number of host threads
        // the same problem described declaratively
        Assembler decl;
        decl.SetCubeMesh(3, 3, 2).SetProbDescr({{FEM_P1, 1}});
        const int q = afb_tet_quadrature(order, nullptr, nullptr, 0);
        std::vector<double> X[4];
        for (int k = 0; k < 4; ++k) {
            X[k].resize(3 * ntet);
            for (int64_t e = 0; e < ntet; ++e) for (int d = 0; d < 3; ++d) X[k][3 * e + d] = xyz[d * nnode + v[k * ntet + e]];
        }
        std::vector<double> XYG(3 * q * ntet), Dq(q * ntet);
        EXPECT(afb_quad_points(decl.context(), order, ntet, X[0].data(), X[1].data(), X[2].data(), X[3].data(), XYG.data(), AFB_HOST) == q);
        for (int64_t k = 0; k < q * ntet; ++k) Dq[k] = 1 + XYG[3 * k] * XYG[3 * k];
        const double one = 1.0;
        using G1 = Operator<GRAD, UFem>; using I1 = Operator<IDEN, UFem>;
        decl.AddMatForm<G1, G1>(0, 0, order, TENSOR_SCALAR, AFB_COEF_PER_POINT, Dq.data()).AddRhsForm<I1>(0, order, TENSOR_SCALAR, AFB_COEF_CONST, &one);
        decl.PrepareProblem();
        std::vector<unsigned char> flag(nnode, 0); std::vector<double> bc(nnode, 0.0);
        for (int64_t n = 0; n < nnode; ++n) flag[n] = label[n] > 0;
        decl.SetDirichlet(flag, bc);
        CsrMatrix Ad; std::vector<double> bd;
        EXPECT(decl.Assemble(Ad, bd) == 0);
        EXPECT(Ad.colind == A1.colind && Ad.rowptr == A1.rowptr);
        double scale = 0, err = 0, errb = 0;
        for (double x : Ad.val) scale = std::fmax(scale, std::fabs(x));
        for (std::size_t k = 0; k < Ad.val.size(); ++k) err = std::fmax(err, std::fabs(Ad.val[k] - A1.val[k]));
        for (std::size_t k = 0; k < bd.size(); ++k) errb = std::fmax(errb, std::fabs(bd[k] - b1[k]));
        EXPECT(err <= 1e-12 * scale && errb <= 1e-13);
        // matrix only / rhs only evaluators and the accumulate semantics of Assemble (assembler.inl:305-306)
        Assembler sep;
        sep.SetCubeMesh(3, 3, 2).SetProbDescr({{FEM_P1, 1}});
        sep.SetMatFunc(GenerateElemMat([&](const double** XY, double* Adat, void* ud) { double Ftmp[UNF]; local_assembler(XY, Adat, Ftmp, ud); }, UNF, UNF));
        sep.SetRHSFunc(GenerateElemRhs([&](const double** XY, double* Fdat, void* ud) { double Atmp[UNF * UNF]; local_assembler(XY, Atmp, Fdat, ud); }, UNF));
        sep.SetDataGatherer(local_data_gatherer);
        CsrMatrix As; std::vector<double> bs;
        EXPECT(sep.AssembleMatrix(As) == 0);
        EXPECT(sep.AssembleRHS(bs) == 0);
        EXPECT(As.val == A1.val && bs == b1);
        EXPECT(sep.Assemble(As, bs) == 0);
        double d2 = 0;
        for (std::size_t k = 0; k < As.val.size(); ++k) d2 = std::fmax(d2, std::fabs(As.val[k] - 2 * A1.val[k]));
        EXPECT(d2 == 0.0);
        Assembler none;
        none.SetCubeMesh(2, 2, 2).SetProbDescr({{FEM_P1, 1}});
        bool thrown = false;
        try { CsrMatrix An; std::vector<double> bn; none.Assemble(An, bn); } catch (std::runtime_error&) { thrown = true; }   // "System local evaluator is not specified"
        EXPECT(thrown);
    }
    // --- SetEnumerator(ANITYPE ... ETDIMBLOCKS) (assembler.h:333-336, global_enumerator.h:393-401): the matrix assembled under every
    //     numbering is the NATURAL one with rows / columns permuted by the map between the two dof tables (Taylor-Hood P2^3 x P1)
    {
        using GU = Operator<GRAD, FemVec<3, FEM_P2>>; using DU = Operator<DIV, FemVec<3, FEM_P2>>; using IP = Operator<IDEN, FemFix<FEM_P1>>;
        using IU = Operator<IDEN, FemVec<3, FEM_P2>>;
        const double fz[3] = {0.0, 0.0, -1.0};
        auto build = [&](Assembler& d, ASSEMBLING_TYPE t) {
            d.SetCubeMesh(3, 2, 2).SetProbDescr({{FEM_P2, 3}, {FEM_P1, 1}}).SetEnumerator(t);
            d.AddMatForm<GU, GU>(0, 0, 2, TENSOR_NULL, AFB_COEF_CONST, nullptr).AddMatForm<IP, DU>(1, 0, 2, TENSOR_NULL, AFB_COEF_CONST, nullptr, -1.0);
            d.AddMatForm<DU, IP>(0, 1, 2, TENSOR_NULL, AFB_COEF_CONST, nullptr, -1.0).AddRhsForm<IU>(0, 2, TENSOR_GENERAL, AFB_COEF_CONST, fz);
            d.PrepareProblem();
        };
        Assembler nat;
        build(nat, NATURAL);
        CsrMatrix An; std::vector<double> bn;
        EXPECT(nat.Assemble(An, bn) == 0);
        int64_t nnode = 0, ntet = 0;
        EXPECT(afb_mesh_get(nat.context(), &nnode, &ntet, nullptr, nullptr, AFB_HOST) == 0);
        std::vector<int32_t> v(4 * ntet);
        EXPECT(afb_mesh_get(nat.context(), &nnode, &ntet, nullptr, v.data(), AFB_HOST) == 0);
        const std::vector<EnumVar> ev = {{FEM_P2, 3}, {FEM_P1, 1}};
        const DofEnumeration en0 = enumerate_dofs(NATURAL, nnode, ntet, v.data(), v.data() + ntet, v.data() + 2 * ntet, v.data() + 3 * ntet, ev);
        {   // the host NATURAL table is the one the device numbering produced
            std::vector<int64_t> rc((std::size_t)ntet * en0.nloc), cc(rc.size());
            int nr, nc; int64_t rb, re, ng;
            EXPECT(afb_dofmap_get(nat.context(), &nr, &nc, &rb, &re, &ng, rc.data(), cc.data(), AFB_HOST) == 0);
            bool same = nr == en0.nloc;
            for (std::size_t k = 0; same && k < rc.size(); ++k) same = cc[k] == en0.elem2dof[k] + 1;
            EXPECT(same);
        }
        const ASSEMBLING_TYPE types[5] = {ANITYPE, MINIBLOCKS, DIMUNION, BYELEMTYPE, ETDIMBLOCKS};
        for (ASSEMBLING_TYPE t : types) {
            Assembler d;
            build(d, t);
            CsrMatrix A; std::vector<double> b;
            EXPECT(d.Assemble(A, b) == 0);
            EXPECT(A.val.size() == An.val.size() && b.size() == bn.size());
            const DofEnumeration en = enumerate_dofs(t, nnode, ntet, v.data(), v.data() + ntet, v.data() + 2 * ntet, v.data() + 3 * ntet, ev);
            std::vector<int64_t> perm(en0.nrows, -1);   // NATURAL id -> id under t
            for (std::size_t k = 0; k < en.elem2dof.size(); ++k) perm[en0.elem2dof[k]] = en.elem2dof[k];
            double scale = 0, err = 0, errb = 0;
            for (double x : An.val) scale = std::fmax(scale, std::fabs(x));
            bool found_all = true;
            for (int64_t r = 0; r + 1 < (int64_t)An.rowptr.size(); ++r) {
                const int64_t pr = perm[r];
                errb = std::fmax(errb, std::fabs(b[pr] - bn[r]));
                for (int64_t k = An.rowptr[r]; k < An.rowptr[r + 1]; ++k) {
                    const int32_t pc = (int32_t)perm[An.colind[k]];
                    const int32_t* lo = std::lower_bound(A.colind.data() + A.rowptr[pr], A.colind.data() + A.rowptr[pr + 1], pc);
                    if (lo == A.colind.data() + A.rowptr[pr + 1] || *lo != pc) { found_all = false; continue; }
                    err = std::fmax(err, std::fabs(A.val[lo - A.colind.data()] - An.val[k]));
                }
                if (A.rowptr[pr + 1] - A.rowptr[pr] != An.rowptr[r + 1] - An.rowptr[r]) found_all = false;
            }
            EXPECT(found_all);
            EXPECT(err <= 1e-12 * scale && errb <= 1e-13);
        }
    }
    std::printf(fails ? "test_shim: %d FAILED\n" : "test_shim: all passed\n", fails);
    return fails ? 1 : 0;
}
