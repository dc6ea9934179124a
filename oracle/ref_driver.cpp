// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// ref_driver.cpp: a thin extern "C" shell around the UNMODIFIED reference element code
// (AniFem++ `fem/`, compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libanifem_ref.so).  Nothing from the reference is copied here: this file only
// *calls* the reference's public API:
//   - Ani::fem3Dtet runtime overload           (anifem++/fem/operations/int_tet.inl:3-27)
//   - Ani::fem3Dtet template+PlainMemory       (anifem++/fem/operations/int_tet.inl:30-57)
//   - tetrahedron_quadrature_formulas(order)   (anifem++/fem/quadrature_formulas.cpp:526-1503)
//   - Ani::Operator<OP,FEM>::apply             (anifem++/fem/spaces/poly_*.h, operators.h)
// It is used (a) to validate the C restatement in oracle/fem_oracle.c, (b) to generate the
// committed golden fixtures in tests/golden/, (c) as the "reference" CPU baseline of bench.py.
#include "anifem++/fem/operations/operations.h"
#include "anifem++/fem/operators.h"
#include "anifem++/fem/spaces/spaces.h"
#include "anifem++/fem/quadrature_formulas.h"

#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include <stdexcept>
#include <string>

using namespace Ani;

namespace {

thread_local std::string g_err;

// Tensor data in the *user callback layout*: a col-major (jdim x idim) matrix K(k,j) at
// D[k + jdim*j] (k = dim index of OpB(v), j = dim index of OpA(u)), exactly what a reference
// user lambda writes into Dmem (anifem++/fem/operations/core.h:27-40, diff_tensor.h:73-96).
struct TensorData {
    int ttype;        // Ani::TensorType
    int layout;       // 0 CONST, 1 PER_TET, 2 PER_POINT
    const double* D;  // CONST: [len]; PER_TET: [len*f]; PER_POINT: [len*q*f]
    long q;
    long base_tet;    // index of tetra 0 of the current fusion block inside D
    mutable long calls;  // running call counter (reference calls r-major, n-minor)
};

struct TensorFunctor {
    TensorData* td;
    TensorType operator()(const std::array<double, 3>& X, double* Dmem, TensorDims dims, void* user_data, int iTet) const {
        (void)X; (void)user_data;
        const long len = (td->ttype == TENSOR_SCALAR) ? 1 : (td->ttype == TENSOR_NULL ? 0 : (long)(dims.first * dims.second));
        long n = 0;
        if (td->layout == 2) { n = td->calls % td->q; }
        td->calls++;
        const double* src = td->D;
        if (td->layout == 1) src += len * (td->base_tet + iTet);
        if (td->layout == 2) src += len * (n + td->q * (td->base_tet + iTet));
        for (long i = 0; i < len; ++i) Dmem[i] = src[i];
        return static_cast<TensorType>(td->ttype);
    }
};

template <int OP>
std::shared_ptr<ApplyOpBase> make_op_fem(int fem, int vec) {
#define CASE(F)                                                                       \
    case F:                                                                           \
        if (vec == 1) return std::make_shared<ApplyOpFromTemplate<OP, FemFix<F>>>(); \
        if (vec == 3) return std::make_shared<ApplyOpFromTemplate<OP, FemVec<3, F>>>(); \
        break;
    switch (fem) {
        CASE(FEM_P0)
        CASE(FEM_P1)
        CASE(FEM_P2)
        CASE(FEM_P3)
    }
#undef CASE
    return nullptr;
}

std::shared_ptr<ApplyOpBase> make_div(int fem, int vec) {
    if (vec != 3) return nullptr;
    switch (fem) {
        case FEM_P1: return std::make_shared<ApplyOpFromTemplate<DIV, FemVec<3, FEM_P1>>>();
        case FEM_P2: return std::make_shared<ApplyOpFromTemplate<DIV, FemVec<3, FEM_P2>>>();
        case FEM_P3: return std::make_shared<ApplyOpFromTemplate<DIV, FemVec<3, FEM_P3>>>();
    }
    return nullptr;
}

std::shared_ptr<ApplyOpBase> make_op(int op, int fem, int vec) {
    switch (op) {
        case IDEN: return make_op_fem<IDEN>(fem, vec);
        case GRAD: return make_op_fem<GRAD>(fem, vec);
        case DIV: return make_div(fem, vec);
    }
    return nullptr;
}

// runtime-operator path, any (op,fem,vec) pair; processes tets in blocks of `fuse`
int run_runtime(int opA, int femA, int vecA, int opB, int femB, int vecB, int order,
                int ttype, int layout, const double* D, long f,
                const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                double* A, long t0, long t1, int fuse) {
    auto oa = make_op(opA, femA, vecA), ob = make_op(opB, femB, vecB);
    if (!oa || !ob) { g_err = "unsupported operator/space"; return -3; }
    (void)f;
    const long nfa = oa->Nfa(), nfb = ob->Nfa();
    auto formula = tetrahedron_quadrature_formulas(order);
    TensorData td{ttype, layout, D, formula.GetNumPoints(), 0, 0};
    TensorFunctor fn{&td};
    std::vector<char> raw;
    PlainMemoryX<> req;
    auto alloc = [&](int ff) {
        if (layout == 0)
            req = fem3Dtet_memory_requirements<DfuncTraits<PerPoint, true>>(*oa, *ob, order, ff);
        else
            req = fem3Dtet_memory_requirements<DfuncTraits<PerPoint, false>>(*oa, *ob, order, ff);
        raw.resize(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
    };
    int cur = -1;
    for (long t = t0; t < t1; t += fuse) {
        int ff = (int)std::min<long>(fuse, t1 - t);
        if (ff != cur) { alloc(ff); cur = ff; }
        td.base_tet = t; td.calls = 0;
        auto XYZ = make_tetras(XY0 + 3 * t, XY1 + 3 * t, XY2 + 3 * t, XY3 + 3 * t, ff);
        DenseMatrix<> Am(A + nfa * nfb * t, nfb, nfa * ff, nfa * nfb * ff);
        if (layout == 0)
            fem3Dtet<DfuncTraits<PerPoint, true>>(XYZ, *oa, *ob, fn, Am, req, order, nullptr);
        else
            fem3Dtet<DfuncTraits<PerPoint, false>>(XYZ, *oa, *ob, fn, Am, req, order, nullptr);
    }
    return 0;
}

// template + PlainMemory path (the CPU baseline named in BASELINE.md section 3)
template <typename OpA, typename OpB, typename Traits>
int run_template(int order, int ttype, int layout, const double* D,
                 const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                 double* A, long t0, long t1, int fuse) {
    constexpr long nfa = OpA::Nfa::value, nfb = OpB::Nfa::value;
    auto formula = tetrahedron_quadrature_formulas(order);
    TensorData td{ttype, layout, D, formula.GetNumPoints(), 0, 0};
    TensorFunctor fn{&td};
    std::vector<char> raw;
    PlainMemory<> req;
    int cur = -1;
    for (long t = t0; t < t1; t += fuse) {
        int ff = (int)std::min<long>(fuse, t1 - t);
        if (ff != cur) {
            req = fem3Dtet_memory_requirements<OpA, OpB>(order, ff);
            raw.resize(req.enoughRawSize());
            req.allocateFromRaw(raw.data(), raw.size());
            cur = ff;
        }
        td.base_tet = t; td.calls = 0;
        DenseMatrix<> x0(const_cast<double*>(XY0 + 3 * t), 3, ff), x1(const_cast<double*>(XY1 + 3 * t), 3, ff),
            x2(const_cast<double*>(XY2 + 3 * t), 3, ff), x3(const_cast<double*>(XY3 + 3 * t), 3, ff);
        DenseMatrix<> Am(A + nfa * nfb * t, nfb, nfa * ff, nfa * nfb * ff);
        fem3Dtet<OpA, OpB, Traits>(x0, x1, x2, x3, fn, Am, req, order, nullptr);
    }
    return 0;
}

using GP1 = Operator<GRAD, FemFix<FEM_P1>>;  using IP1 = Operator<IDEN, FemFix<FEM_P1>>;
using GP2 = Operator<GRAD, FemFix<FEM_P2>>;  using IP2 = Operator<IDEN, FemFix<FEM_P2>>;
using GP3 = Operator<GRAD, FemFix<FEM_P3>>;  using IP3 = Operator<IDEN, FemFix<FEM_P3>>;
using GP2v = Operator<GRAD, FemVec<3, FEM_P2>>; using IP2v = Operator<IDEN, FemVec<3, FEM_P2>>;
using DP2v = Operator<DIV, FemVec<3, FEM_P2>>;
using IP0 = Operator<IDEN, FemFix<FEM_P0>>;

// returns 1 if a template instantiation exists for this form and was run
int try_template(int opA, int femA, int vecA, int opB, int femB, int vecB, int order,
                 int ttype, int layout, const double* D,
                 const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                 double* A, long t0, long t1, int fuse, int* rc) {
    const long key = (((((long)opA * 8 + femA) * 4 + vecA) * 8 + opB) * 8 + femB) * 4 + vecB;
    auto K = [](int oa, int fa, int va, int ob, int fb, int vb) {
        return (((((long)oa * 8 + fa) * 4 + va) * 8 + ob) * 8 + fb) * 4 + vb;
    };
#define RUN(OA, OB)                                                                                          \
    {                                                                                                        \
        if (layout == 0) *rc = run_template<OA, OB, DfuncTraits<PerPoint, true>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, t0, t1, fuse); \
        else if (ttype == TENSOR_SYMMETRIC) *rc = run_template<OA, OB, DfuncTraits<TENSOR_SYMMETRIC, false>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, t0, t1, fuse); \
        else if (ttype == TENSOR_SCALAR) *rc = run_template<OA, OB, DfuncTraits<TENSOR_SCALAR, false>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, t0, t1, fuse); \
        else *rc = run_template<OA, OB, DfuncTraits<PerPoint, false>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, t0, t1, fuse); \
        return 1;                                                                                            \
    }
    if (key == K(GRAD, FEM_P1, 1, GRAD, FEM_P1, 1)) RUN(GP1, GP1)
    if (key == K(IDEN, FEM_P1, 1, IDEN, FEM_P1, 1)) RUN(IP1, IP1)
    if (key == K(GRAD, FEM_P2, 1, GRAD, FEM_P2, 1)) RUN(GP2, GP2)
    if (key == K(IDEN, FEM_P2, 1, IDEN, FEM_P2, 1)) RUN(IP2, IP2)
    if (key == K(GRAD, FEM_P3, 1, GRAD, FEM_P3, 1)) RUN(GP3, GP3)
    if (key == K(IDEN, FEM_P3, 1, IDEN, FEM_P3, 1)) RUN(IP3, IP3)
    if (key == K(GRAD, FEM_P2, 3, GRAD, FEM_P2, 3)) RUN(GP2v, GP2v)
    if (key == K(IDEN, FEM_P1, 1, DIV, FEM_P2, 3)) RUN(IP1, DP2v)
    if (key == K(DIV, FEM_P2, 3, IDEN, FEM_P1, 1)) RUN(DP2v, IP1)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P1, 1)) RUN(IP0, IP1)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P2, 1)) RUN(IP0, IP2)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P3, 1)) RUN(IP0, IP3)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P2, 3)) RUN(IP0, IP2v)
#undef RUN
    return 0;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// quadrature rule of the reference: p[4*q] barycentric, w[q]; returns q (or -1)
int ref_tet_quadrature(int order, double* p, double* w, int cap) {
    try {
        auto f = tetrahedron_quadrature_formulas(order);
        int q = f.GetNumPoints();
        if (p && w) {
            if (cap < q) return -1;
            std::memcpy(p, f.p, sizeof(double) * 4 * q);
            std::memcpy(w, f.w, sizeof(double) * q);
        }
        return q;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int ref_op_dims(int op, int fem, int vec, int* nfa, int* dim) {
    auto o = make_op(op, fem, vec);
    if (!o) return -3;
    *nfa = o->Nfa(); *dim = o->Dim();
    return 0;
}

// Batched element matrices through the reference. XYk: 3 x f col-major. A: nfB x (nfA*f) col-major.
// mode 0: runtime-operator overload; mode 1: template+PlainMemory overload when instantiated
// (falls back to runtime otherwise). nthreads: contiguous tet ranges over std::thread, mirroring
// ThreadPar::ParallelFor<STD> (anifem++/fem/mutex_type.h:110-131).
int ref_fem3dtet(int opA, int femA, int vecA, int opB, int femB, int vecB, int order,
                 int ttype, int layout, const double* D, long f,
                 const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                 double* A, int mode, int fuse, int nthreads) {
    if (fuse < 1) fuse = 1;
    if (nthreads < 1) nthreads = 1;
    std::vector<int> rcs(nthreads, 0);
    std::vector<std::string> errs(nthreads);
    auto work = [&](int th) {
        long t0 = f * th / nthreads, t1 = f * (th + 1) / nthreads;
        try {
            int rc = 0;
            if (mode == 1 && try_template(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, D, XY0, XY1, XY2, XY3, A, t0, t1, fuse, &rc)) {
                rcs[th] = rc;
            } else {
                rcs[th] = run_runtime(opA, femA, vecA, opB, femB, vecB, order, ttype, layout, D, f, XY0, XY1, XY2, XY3, A, t0, t1, fuse);
                if (rcs[th]) errs[th] = g_err;
            }
        } catch (std::exception& e) { errs[th] = e.what(); rcs[th] = -4; }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> ths;
        for (int i = 0; i < nthreads; ++i) ths.emplace_back(work, i);
        for (auto& t : ths) t.join();
    }
    for (int i = 0; i < nthreads; ++i) if (rcs[i]) { g_err = errs[i]; return rcs[i]; }
    return 0;
}

// Operator<OP,FEM>::apply table U[k + dim*(n + q*(i + nfa*r))] on f tets with an explicit rule
// (XYL 4*q barycentric) -- used to regenerate the golden U tables of
// tests/fem/spaces/predefined_spaces_test.cpp:11-643. Vector spaces return the dense expansion.
int ref_operator_apply(int op, int fem, int vec, int q, const double* XYL, const double* W, long f,
                       const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* Uout) {
    try {
        auto o = make_op(op, fem, vec);
        if (!o) { g_err = "unsupported operator/space"; return -3; }
        auto oreq = o->getMemoryRequirements(q, f);
        // build an AniMemoryX by hand the way the reference test does (predefined_spaces_test.cpp:28-50)
        std::vector<double> XYP(12 * f), PSI(9 * f), DET(f), MES(f), XYG(3 * q * f), Ud(oreq.Usz + 8), eR(oreq.extraRsz + 8);
        std::vector<int> eI(oreq.extraIsz + 8);
        std::vector<DenseMatrix<>> mtx(2 * (oreq.mtx_parts + 2));
        std::vector<int> mrow(4 * (oreq.mtx_parts + 2)), mcol(4 * (oreq.mtx_parts + 2));
        for (long r = 0; r < f; ++r) {
            const double* X[4] = {XY0 + 3 * r, XY1 + 3 * r, XY2 + 3 * r, XY3 + 3 * r};
            for (int l = 0; l < 4; ++l) for (int k = 0; k < 3; ++k) XYP[12 * r + 3 * l + k] = X[l][k] - X[0][k];
            DET[r] = inverse3x3(XYP.data() + 12 * r + 3, PSI.data() + 9 * r);
            MES[r] = std::abs(DET[r]) / 6;
        }
        AniMemoryX<> mem;
        mem.XYP.Init(XYP.data(), 12 * f); mem.PSI.Init(PSI.data(), 9 * f); mem.DET.Init(DET.data(), f); mem.MES.Init(MES.data(), f);
        mem.XYG.Init(XYG.data(), 3 * q * f);
        mem.XYL.Init(const_cast<double*>(XYL), 4 * q); mem.WG.Init(const_cast<double*>(W), q);
        mem.U.Init(Ud.data(), oreq.Usz); mem.extraR.Init(eR.data(), oreq.extraRsz); mem.extraI.Init(eI.data(), oreq.extraIsz);
        mem.MTX.Init(mtx.data(), mtx.size()); mem.MTXI_ROW.Init(mrow.data(), mrow.size()); mem.MTXI_COL.Init(mcol.data(), mcol.size());
        mem.busy_mtx_parts = 0;
        mem.q = q; mem.f = f;
        BandDenseMatrixX<> U = (*o)(mem, mem.U);
        const long dim = o->Dim(), nfa = o->Nfa();
        std::fill(Uout, Uout + dim * q * nfa * f, 0.0);
        for (std::size_t d1 = 0; d1 < U.nparts; ++d1) {
            long isz = U.stCol[d1 + 1] - U.stCol[d1], ksz = U.stRow[d1 + 1] - U.stRow[d1];
            for (long r = 0; r < f; ++r)
                for (long i = U.stCol[d1]; i < U.stCol[d1 + 1]; ++i)
                    for (long n = 0; n < q; ++n)
                        for (long k = U.stRow[d1]; k < U.stRow[d1 + 1]; ++k)
                            Uout[k + dim * (n + q * (i + nfa * r))] =
                                U.data[d1].data[(k - U.stRow[d1]) + n * ksz + U.data[d1].nRow * ((i - U.stCol[d1]) + isz * r)];
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -4; }
}

}  // extern "C"
