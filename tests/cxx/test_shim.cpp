// GPU test of the C++ mirror of the reference API (anifem_b200/fem.hpp, assembler.hpp).  Reads like the reference's own
// tests/fem/operations/int_tet_test.cpp: the same calls, the same golden tables.  Exit code 0 = pass.
#include <cfloat>
#include <cmath>
#include <cstdio>

#include "anifem_b200/assembler.hpp"

using namespace Ani;

static double norm_diff(double a, const DenseMatrix<>& A, double b, const double* B) {
    double s = 0;
    for (std::size_t i = 0; i < A.nRow * A.nCol; ++i) { double d = a * A.data[i] + b * B[i]; s += d * d; }
    return std::sqrt(s);
}
static double norm(const DenseMatrix<>& A) { double s = 0; for (std::size_t i = 0; i < A.nRow * A.nCol; ++i) s += A.data[i] * A.data[i]; return std::sqrt(s); }

#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

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
    std::printf(fails ? "test_shim: %d FAILED\n" : "test_shim: all passed\n", fails);
    return fails ? 1 : 0;
}
