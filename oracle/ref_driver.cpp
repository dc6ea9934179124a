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
#include "anifem++/fem/tetdofmap.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <algorithm>
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


// ---- per-thread single-cell evaluators for the CPU-baseline assembler ---------------------------
struct CellRunner {
    virtual ~CellRunner() {}
    virtual void run(const double* X0, const double* X1, const double* X2, const double* X3, long e, double* A) = 0;
    long nfa = 0, nfb = 0;
};

template <typename OpA, typename OpB, typename Traits>
struct TemplateRunner : CellRunner {
    TensorData td;
    int order;
    std::vector<char> raw;
    PlainMemory<> req;
    TemplateRunner(int order_, int ttype, int layout, const double* D) : order(order_) {
        nfa = OpA::Nfa::value; nfb = OpB::Nfa::value;
        td = TensorData{ttype, layout, D, tetrahedron_quadrature_formulas(order).GetNumPoints(), 0, 0};
        req = fem3Dtet_memory_requirements<OpA, OpB>(order, 1);
        raw.resize(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
    }
    void run(const double* X0, const double* X1, const double* X2, const double* X3, long e, double* A) override {
        td.base_tet = e; td.calls = 0;
        TensorFunctor fn{&td};
        DenseMatrix<> x0(const_cast<double*>(X0), 3, 1), x1(const_cast<double*>(X1), 3, 1), x2(const_cast<double*>(X2), 3, 1), x3(const_cast<double*>(X3), 3, 1);
        DenseMatrix<> Am(A, nfb, nfa, nfa * nfb);
        fem3Dtet<OpA, OpB, Traits>(x0, x1, x2, x3, fn, Am, req, order, nullptr);
    }
};

struct RuntimeRunner : CellRunner {
    std::shared_ptr<ApplyOpBase> oa, ob;
    TensorData td;
    int order, layout;
    std::vector<char> raw;
    PlainMemoryX<> req;
    RuntimeRunner(std::shared_ptr<ApplyOpBase> a, std::shared_ptr<ApplyOpBase> b, int order_, int ttype, int layout_, const double* D)
        : oa(a), ob(b), order(order_), layout(layout_) {
        nfa = oa->Nfa(); nfb = ob->Nfa();
        td = TensorData{ttype, layout, D, tetrahedron_quadrature_formulas(order).GetNumPoints(), 0, 0};
        if (layout == 0) req = fem3Dtet_memory_requirements<DfuncTraits<PerPoint, true>>(*oa, *ob, order, 1);
        else req = fem3Dtet_memory_requirements<DfuncTraits<PerPoint, false>>(*oa, *ob, order, 1);
        raw.resize(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
    }
    void run(const double* X0, const double* X1, const double* X2, const double* X3, long e, double* A) override {
        td.base_tet = e; td.calls = 0;
        TensorFunctor fn{&td};
        auto XYZ = make_tetras(X0, X1, X2, X3, 1);
        DenseMatrix<> Am(A, nfb, nfa, nfa * nfb);
        if (layout == 0) fem3Dtet<DfuncTraits<PerPoint, true>>(XYZ, *oa, *ob, fn, Am, req, order, nullptr);
        else fem3Dtet<DfuncTraits<PerPoint, false>>(XYZ, *oa, *ob, fn, Am, req, order, nullptr);
    }
};

std::unique_ptr<CellRunner> make_runner(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D) {
    auto K = [](int oa, int fa, int va, int ob, int fb, int vb) { return (((((long)oa * 8 + fa) * 4 + va) * 8 + ob) * 8 + fb) * 4 + vb; };
    const long key = K(opA, femA, vecA, opB, femB, vecB);
#define MK(OA, OB)                                                                                                   \
    {                                                                                                                \
        if (layout == 0) return std::unique_ptr<CellRunner>(new TemplateRunner<OA, OB, DfuncTraits<PerPoint, true>>(order, ttype, layout, D)); \
        if (ttype == TENSOR_SYMMETRIC) return std::unique_ptr<CellRunner>(new TemplateRunner<OA, OB, DfuncTraits<TENSOR_SYMMETRIC, false>>(order, ttype, layout, D)); \
        if (ttype == TENSOR_SCALAR) return std::unique_ptr<CellRunner>(new TemplateRunner<OA, OB, DfuncTraits<TENSOR_SCALAR, false>>(order, ttype, layout, D)); \
        return std::unique_ptr<CellRunner>(new TemplateRunner<OA, OB, DfuncTraits<PerPoint, false>>(order, ttype, layout, D)); \
    }
    if (key == K(GRAD, FEM_P1, 1, GRAD, FEM_P1, 1)) MK(GP1, GP1)
    if (key == K(IDEN, FEM_P1, 1, IDEN, FEM_P1, 1)) MK(IP1, IP1)
    if (key == K(GRAD, FEM_P2, 1, GRAD, FEM_P2, 1)) MK(GP2, GP2)
    if (key == K(IDEN, FEM_P2, 1, IDEN, FEM_P2, 1)) MK(IP2, IP2)
    if (key == K(GRAD, FEM_P3, 1, GRAD, FEM_P3, 1)) MK(GP3, GP3)
    if (key == K(IDEN, FEM_P3, 1, IDEN, FEM_P3, 1)) MK(IP3, IP3)
    if (key == K(GRAD, FEM_P2, 3, GRAD, FEM_P2, 3)) MK(GP2v, GP2v)
    if (key == K(IDEN, FEM_P1, 1, DIV, FEM_P2, 3)) MK(IP1, DP2v)
    if (key == K(DIV, FEM_P2, 3, IDEN, FEM_P1, 1)) MK(DP2v, IP1)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P1, 1)) MK(IP0, IP1)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P2, 1)) MK(IP0, IP2)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P3, 1)) MK(IP0, IP3)
    if (key == K(IDEN, FEM_P0, 1, IDEN, FEM_P2, 3)) MK(IP0, IP2v)
#undef MK
    auto oa = make_op(opA, femA, vecA), ob = make_op(opB, femB, vecB);
    if (!oa || !ob) return nullptr;
    return std::unique_ptr<CellRunner>(new RuntimeRunner(oa, ob, order, ttype, layout, D));
}

}  // namespace

template <typename Op>
static int ref_apply_impl(int mode, long f, int q, const double* pts, const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                          const double* dofs, double* out) {
    constexpr int nfa = Op::Nfa::value, dim = Op::Dim::value;
    if (mode == 0) {
        auto req = fem3DapplyL_memory_requirements<Op>(q, (int)f);
        // slack behind the planned block: for multi-part (vector) operators the reference's evaluation stages dim*q*f values in its
        // last scratch region, which the planner sizes nfa*f (core.inl:384, eval.inl:61)
        req.dSize += (std::size_t)dim * q * f + (std::size_t)nfa * f + 64;
        std::vector<char> raw(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
        DenseMatrix<> d(const_cast<double*>(dofs), nfa, f), o(out, (std::size_t)dim * q, f);
        fem3DapplyL<Op>(make_tetras(XY0, XY1, XY2, XY3, (int)f), ArrayView<>(const_cast<double*>(pts), 4 * q), d, o, req);
    } else {
        auto req = fem3DapplyX_memory_requirements<Op>(q, 1);
        req.dSize += (std::size_t)dim * q + (std::size_t)nfa + 64;
        std::vector<char> raw(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
        Tetra<const double> T(XY0, XY1, XY2, XY3);
        fem3DapplyX<Op>(T, ArrayView<const double>(pts, 3 * q), ArrayView<>(const_cast<double*>(dofs), nfa), ArrayView<>(out, (std::size_t)dim * q), req);
    }
    return 0;
}

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

// triangle rule of the reference: p[3*q] barycentric, w[q]; returns q (or -1)
int ref_tri_quadrature(int order, double* p, double* w, int cap) {
    try {
        auto f = triangle_quadrature_formulas(order);
        int q = f.GetNumPoints();
        if (p && w) {
            if (cap < q) return -1;
            std::memcpy(p, f.p, sizeof(double) * 3 * q);
            std::memcpy(w, f.w, sizeof(double) * q);
        }
        return q;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Surface-integral element matrices through the reference's runtime fem3Dface (fem/operations/int_face.inl:160-199), one
// tet per call with its own face number face[r]. Layouts as ref_fem3dtet; PER_POINT coefficients are indexed by the points of
// the triangle rule.
int ref_fem3dface(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D,
                  long f, const int* face, const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A) {
    try {
        auto oa = make_op(opA, femA, vecA), ob = make_op(opB, femB, vecB);
        if (!oa || !ob) { g_err = "unsupported operator/space"; return -3; }
        const long nfa = oa->Nfa(), nfb = ob->Nfa();
        auto formula = triangle_quadrature_formulas(order);
        TensorData td{ttype, layout, D, formula.GetNumPoints(), 0, 0};
        TensorFunctor fn{&td};
        std::vector<char> raw;
        PlainMemoryX<> req;
        if (layout == 0) req = fem3Dface_memory_requirements<DfuncTraits<PerPoint, true>>(*oa, *ob, order, 1);
        else req = fem3Dface_memory_requirements<DfuncTraits<PerPoint, false>>(*oa, *ob, order, 1);
        raw.resize(req.enoughRawSize());
        req.allocateFromRaw(raw.data(), raw.size());
        for (long t = 0; t < f; ++t) {
            td.base_tet = t; td.calls = 0;
            auto XYZ = make_tetras(XY0 + 3 * t, XY1 + 3 * t, XY2 + 3 * t, XY3 + 3 * t, 1);
            DenseMatrix<> Am(A + nfa * nfb * t, nfb, nfa, nfa * nfb);
            if (layout == 0) fem3Dface<DfuncTraits<PerPoint, true>>(XYZ, face[t], *oa, *ob, fn, Am, req, order, nullptr);
            else fem3Dface<DfuncTraits<PerPoint, false>>(XYZ, face[t], *oa, *ob, fn, Am, req, order, nullptr);
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -4; }
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

// One form of the CPU-baseline assembler (mirrors struct afb_form; is_rhs: OpA = IDEN(P0) trick)
struct RefForm {
    int opA, femA, vecA, opB, femB, vecB, order, ttype, layout, is_rhs;
    const double* D;
    double alpha;
    int row_off, col_off;
};

// CPU baseline = the reference's element code (unmodified, one cell per call exactly like
// AssemblerT::Assemble, assembler.inl:353-381) + the restated scatter of assembler.inl:397-481 into a
// pre-built sorted CSR (binary search = the is_mtx_include_template branch :428-438), cell ranges split
// over std::thread like ThreadPar::ParallelFor<STD> (fem/mutex_type.h:110-131), row updates serialised by
// a per-row spin lock standing in for INMOST::Sparse::LockService (assembler.inl:339-340,406,474).
// NOTE: this is FASTER than the true reference scatter (INMOST Row linear find-or-append + reallocs).
int ref_assemble_csr(int nforms, const RefForm* forms, long ntet, const double* coords /*nnode x 3*/, const long* tets /*ntet x 4*/,
                     int nloc, const long* codes /*ntet x nloc, sign*(id+1)*/, long row_begin, long nrows,
                     const long* rowptr, const int* colind, double* val, double* rhs, double drop_val, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::atomic_flag> locks(nthreads > 1 ? nrows : 0);
    for (auto& l : locks) l.clear();
    std::vector<int> status(nthreads, 0);
    std::vector<std::string> errs(nthreads);
    auto work = [&](int th) {
        try {
            std::vector<std::unique_ptr<CellRunner>> runners;
            for (int k = 0; k < nforms; ++k) {
                const RefForm& f = forms[k];
                runners.push_back(make_runner(f.opA, f.femA, f.vecA, f.opB, f.femB, f.vecB, f.order, f.ttype, f.layout, f.D));
                if (!runners.back()) { status[th] = -3; errs[th] = "unsupported operator/space"; return; }
            }
            std::vector<double> A((size_t)nloc * nloc), F(nloc), blk((size_t)nloc * nloc);
            const long t0 = ntet * th / nthreads, t1 = ntet * (th + 1) / nthreads;
            for (long e = t0; e < t1; ++e) {
                const long* nd = tets + 4 * e;
                const double *X0 = coords + 3 * nd[0], *X1 = coords + 3 * nd[1], *X2 = coords + 3 * nd[2], *X3 = coords + 3 * nd[3];
                std::fill(A.begin(), A.end(), 0.0);  // ElementalAssembler::update (elemental_assembler.cpp:105-117)
                std::fill(F.begin(), F.end(), 0.0);
                for (int k = 0; k < nforms; ++k) {
                    const RefForm& f = forms[k];
                    CellRunner& r = *runners[k];
                    r.run(X0, X1, X2, X3, e, blk.data());
                    if (f.is_rhs) for (long ib = 0; ib < r.nfb; ++ib) F[f.row_off + ib] += f.alpha * blk[ib];
                    else
                        for (long ia = 0; ia < r.nfa; ++ia)
                            for (long ib = 0; ib < r.nfb; ++ib) A[(f.row_off + ib) + (size_t)nloc * (f.col_off + ia)] += f.alpha * blk[ib + r.nfb * ia];
                }
                const long* cd = codes + (size_t)nloc * e;
                for (int i = 0; i < nloc; ++i) {
                    if (cd[i] == 0) continue;
                    const long rid = std::labs(cd[i]) - 1 - row_begin;
                    if (rid < 0 || rid >= nrows) continue;  // ghost row
                    const int rs = cd[i] < 0 ? -1 : 1;
                    if (nthreads > 1) while (locks[rid].test_and_set(std::memory_order_acquire)) {}
                    if (rhs) rhs[rid] += rs * F[i];
                    if (val) {
                        const long b = rowptr[rid], en = rowptr[rid + 1];
                        for (int j = 0; j < nloc; ++j) {
                            const double a = A[i + (size_t)nloc * j];
                            if (!std::isfinite(a)) { status[th] = -1; continue; }
                            if (!(std::fabs(a) > drop_val)) continue;
                            const long cid = std::labs(cd[j]) - 1;
                            const int cs = cd[j] < 0 ? -1 : 1;
                            const int* it = std::lower_bound(colind + b, colind + en, (int)cid);
                            if (it != colind + en && *it == cid) val[it - colind] += rs * cs * a;
                        }
                    }
                    if (nthreads > 1) locks[rid].clear(std::memory_order_release);
                    if (!std::isfinite(F[i])) status[th] = -1;
                }
            }
        } catch (std::exception& ex) { status[th] = -4; errs[th] = ex.what(); }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> ths;
        for (int i = 0; i < nthreads; ++i) ths.emplace_back(work, i);
        for (auto& t : ths) t.join();
    }
    int st = 0;
    for (int i = 0; i < nthreads; ++i) if (status[i] < st) { st = status[i]; g_err = errs[i]; }
    return st;
}


// ---- essential conditions on a local matrix (fem/operations/dc_on_dof.h): the reference's own helpers on caller data.
// A: n x n column-major, F: n; dof_id[d], Vorth d x d column-major, bc[ndc], dc_orth[ndc] or NULL.
// what: 0 applyVectorDir(A, F), 1 applyVectorDirMatrix(A), 2 applyVectorDirResidual(F), 3 applyDir(A, F, dof_id[0], bc[0]),
//       4 applyVectorDirMatrixExtCol(A as a column of n entries), 5 applyVectorDirMatrixExtRow(A as a row of n entries)
int ref_dirichlet_local(int what, int n, double* A, double* F, int d, const unsigned* dof_id, const double* Vorth, int ndc, const double* bc,
                        const unsigned* dc_orth) {
    using namespace Ani;
    try {
        DenseMatrix<double> Am(A, n, n), Fm(F, n, 1), V(const_cast<double*>(Vorth), d, d);
        std::vector<double> work((size_t)d * n + 16);
        ArrayView<double> mem(work.data(), work.size()), bcv(const_cast<double*>(bc), ndc);
        switch (what) {
            case 0: applyVectorDir<double>(Am, Fm, dof_id, V, bcv, mem, (uint)ndc, dc_orth); break;
            case 1: applyVectorDirMatrix<double>(Am, dof_id, V, mem, (uint)ndc, dc_orth); break;
            case 2: applyVectorDirResidual<double>(Fm, dof_id, V, mem, (uint)ndc, dc_orth); break;
            case 3: applyDir<double>(Am, Fm, (int)dof_id[0], bc[0]); break;
            case 4: { ArrayView<double> col(A, n); applyVectorDirMatrixExtCol<double>(col, dof_id, V, mem, (uint)ndc, dc_orth); break; }
            case 5: { ArrayView<double> row(A, n); applyVectorDirMatrixExtRow<double>(row, dof_id, V, (uint)ndc, dc_orth); break; }
            default: return -7;
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}


// ---- local dof maps (fem/tetdofmap.h): the reference's own classes built from a flat description.
// spec: 1 n0..n5 = UniteDofMap | 2 dim <map> = VectorDofMap | 3 k <map>*k = ComplexDofMap | 4 k <map>*k = product with simplifications
// (operator*) | 5 k <map> = operator^.  out[3*gid + {0,1,2}] = {etype, nelem, leid}; sel (or NULL): bits of the chosen nodes / edges /
// faces / cell -> by_sp[0] = count, by_sp[1..] = tet indices of the dofs on the selection (ascending).  Returns NumDofOnTet or < 0.
static Ani::DofT::DofMap ref_build_dofmap(const int*& p) {
    using namespace Ani::DofT;
    const int kind = *p++;
    if (kind == 1) { std::array<uint, NGEOM_TYPES> n; for (int t = 0; t < NGEOM_TYPES; ++t) n[t] = (uint)*p++; return DofMap(std::make_shared<UniteDofMap>(n)); }
    if (kind == 2) { const int dim = *p++; DofMap b = ref_build_dofmap(p); return DofMap(std::make_shared<VectorDofMap>(dim, b.base())); }
    if (kind == 3) { const int k = *p++; std::vector<DofMap> v; for (int i = 0; i < k; ++i) v.push_back(ref_build_dofmap(p)); return merge(v); }
    if (kind == 4) { const int k = *p++; std::vector<DofMap> v; for (int i = 0; i < k; ++i) v.push_back(ref_build_dofmap(p)); return merge_with_simplifications(v); }
    if (kind == 5) { const int k = *p++; DofMap b = ref_build_dofmap(p); return b ^ (uint)k; }
    throw std::runtime_error("bad dof map description");
}
int ref_dofmap_table(const int* spec, int* out, int cap, const int* sel, int* by_sp) {
    using namespace Ani::DofT;
    try {
        const int* p = spec;
        DofMap m = ref_build_dofmap(p);
        const int n = (int)m.NumDofOnTet();
        if (n > cap) return -2;
        for (int g = 0; g < n; ++g) {
            LocalOrder lo = m.LocalOrderOnTet(TetOrder((uint)g));
            out[3 * g] = lo.etype; out[3 * g + 1] = lo.nelem; out[3 * g + 2] = (int)lo.leid;
        }
        if (sel && by_sp) {
            TetGeomSparsity sp;
            for (int d = 0; d < 4; ++d) for (int i = 0; i < 6; ++i) if ((sel[d] >> i) & 1) sp.set((uchar)d, i, false);
            std::vector<int> ids;
            for (auto it = m.beginBySparsity(sp, false); it != m.endBySparsity(); ++it) ids.push_back((int)(*it).gid);
            std::sort(ids.begin(), ids.end());
            by_sp[0] = (int)ids.size();
            for (size_t k = 0; k < ids.size(); ++k) by_sp[1 + k] = ids[k];
        }
        return n;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// structural equality of two descriptions (DofMap::operator==)
int ref_dofmap_equal(const int* spec_a, const int* spec_b) {
    try {
        const int *pa = spec_a, *pb = spec_b;
        return ref_build_dofmap(pa) == ref_build_dofmap(pb) ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// closure of a selection: in/out bits of nodes / edges / faces / cell after set(dim, i, with_closure) then unset(udim, ui, with_closure)
int ref_sparsity_ops(int dim, int i, int closure, int udim, int ui, int uclosure, int* bits) {
    using namespace Ani::DofT;
    try {
        TetGeomSparsity sp;
        sp.set((uchar)dim, i, closure != 0);
        if (udim >= 0) sp.unset((uchar)udim, ui, uclosure != 0);
        for (int d = 0; d < 4; ++d) { auto ids = sp.getElemsIds((uchar)d); bits[d] = 0; for (int k = 0; k < ids.second; ++k) bits[d] |= 1 << ids.first[k]; }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}


// ---- composite spaces (fem/operators.h:71-74, 157-259): the reference's own FemVecT / FemCom operators on caller data.
// which: 0  IDEN(FemCom<FemVec<3,P2>, FemFix<P1>>) x same            (34 x 34, tensor 4 x 4)
//        1  GRAD(FemVecT<2, FemFix<P1>>) x GRAD(FemCom<FemFix<P1>, FemFix<P1>>)   (8 x 8, tensor 6 x 6)
//        2  IDEN(FemCom<FemFix<P2>, FemFix<P0>>) -> GRAD(FemFix<P1>)   (4 x 11, tensor 3 x 2)
//        3  IDEN(FemVecT<3, FemFix<P1>>) x same                       (12 x 12, tensor 3 x 3)
//        4  GRAD(FemCom<FemVecT<2, FemFix<P2>>, FemFix<P1>>) x IDEN(FemCom<FemFix<P1>, FemFix<P1>, FemFix<P0>>)  (9 x 24, tensor 3 x 9)
// D in the user-callback layout (col-major Dim(OpB) x Dim(OpA)); A: nfB x (nfA * f) col-major.  Returns 0, or -1 (message in ref_last_error).
int ref_fem3dtet_composite(int which, int order, int ttype, int layout, const double* D, long f, const double* XY0, const double* XY1,
                           const double* XY2, const double* XY3, double* A) {
    try {
        using Stokes = FemCom<FemVec<3, FEM_P2>, FemFix<FEM_P1>>;
        using P1x2 = FemVecT<2, FemFix<FEM_P1>>;
        using P1P1 = FemCom<FemFix<FEM_P1>, FemFix<FEM_P1>>;
        using P2P0 = FemCom<FemFix<FEM_P2>, FemFix<FEM_P0>>;
        using P1x3 = FemVecT<3, FemFix<FEM_P1>>;
        using Mixed = FemCom<FemVecT<2, FemFix<FEM_P2>>, FemFix<FEM_P1>>;
        using Tri = FemCom<FemFix<FEM_P1>, FemFix<FEM_P1>, FemFix<FEM_P0>>;
        switch (which) {
            case 0: return run_template<Operator<IDEN, Stokes>, Operator<IDEN, Stokes>, DfuncTraits<>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, 0, f, (int)f);
            case 1: return run_template<Operator<GRAD, P1x2>, Operator<GRAD, P1P1>, DfuncTraits<>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, 0, f, (int)f);
            case 2: return run_template<Operator<IDEN, P2P0>, Operator<GRAD, FemFix<FEM_P1>>, DfuncTraits<>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, 0, f, (int)f);
            case 3: return run_template<Operator<IDEN, P1x3>, Operator<IDEN, P1x3>, DfuncTraits<>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, 0, f, (int)f);
            case 4: return run_template<Operator<GRAD, Mixed>, Operator<IDEN, Tri>, DfuncTraits<>>(order, ttype, layout, D, XY0, XY1, XY2, XY3, A, 0, f, (int)f);
        }
        g_err = "unknown composite case";
        return -7;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}


// ---- fem3DfaceN (fem/operations/int_face.h:32-47, 97-133, dyn_ops.inl:44-50): int_f ((D OpA(u)) . N) . OpB(v), the callback
// fills a col-major (3 Dim(OpB) x Dim(OpA)) tensor, normal component fastest.  Runtime operators, one tet per call, layouts as
// ref_fem3dface.
int ref_fem3dfaceN(int opA, int femA, int vecA, int opB, int femB, int vecB, int order, int ttype, int layout, const double* D,
                   long f, const int* face, const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A) {
    try {
        auto oa = make_op(opA, femA, vecA), ob = make_op(opB, femB, vecB);
        if (!oa || !ob) { g_err = "unsupported operator/space"; return -3; }
        const long nfa = oa->Nfa(), nfb = ob->Nfa();
        auto formula = triangle_quadrature_formulas(order);
        TensorData td{ttype, layout, D, formula.GetNumPoints(), 0, 0};
        TensorFunctor fn{&td};
        for (long t = 0; t < f; ++t) {
            td.base_tet = t; td.calls = 0;
            auto XYZ = make_tetras(XY0 + 3 * t, XY1 + 3 * t, XY2 + 3 * t, XY3 + 3 * t, 1);
            DenseMatrix<> Am(A + nfa * nfb * t, nfb, nfa, nfa * nfb);
            DynMem<> wmem;
            if (layout == 0) fem3DfaceN<DfuncTraits<PerPoint, true>>(XYZ, face[t], *oa, *ob, fn, Am, wmem, order, nullptr);
            else fem3DfaceN<DfuncTraits<PerPoint, false>>(XYZ, face[t], *oa, *ob, fn, Am, wmem, order, nullptr);
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -4; }
}


// ---- FE-function evaluation (fem/operations/eval.h): the reference's own fem3DapplyL / fem3DapplyX for a few operators.
// which: 0 GRAD(P2), 1 IDEN(P3), 2 IDEN(FemVec<3,P1>), 3 GRAD(FemVec<3,P2>).  mode 0: fem3DapplyL with XYL[4q] on f tets;
// mode 1: fem3DapplyX with physical points pts[3q] on tet 0 (f must be 1).  dofs: nfa x f, out: (dim*q) x f.
int ref_fem3dapply(int which, int mode, long f, int q, const double* pts, const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                   const double* dofs, double* out) {
    try {
        switch (which) {
            case 0: return ref_apply_impl<Operator<GRAD, FemFix<FEM_P2>>>(mode, f, q, pts, XY0, XY1, XY2, XY3, dofs, out);
            case 1: return ref_apply_impl<Operator<IDEN, FemFix<FEM_P3>>>(mode, f, q, pts, XY0, XY1, XY2, XY3, dofs, out);
            case 2: return ref_apply_impl<Operator<IDEN, FemVec<3, FEM_P1>>>(mode, f, q, pts, XY0, XY1, XY2, XY3, dofs, out);
            case 3: return ref_apply_impl<Operator<GRAD, FemVec<3, FEM_P2>>>(mode, f, q, pts, XY0, XY1, XY2, XY3, dofs, out);
        }
        g_err = "unknown operator case";
        return -7;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
