// FE-function evaluation at points: Op(u_h)(x_n) on every tetrahedron.
//
// Replaces (reference): fem3DapplyL (fem/operations/eval.h:13-120) -> internalFem3DApply (fem/operations/core.inl:369-404):
//     opU[k + dim*(n + q*r)] = sum_i U[k + dim*(n + q*(i + nfa*r))] * dofs[i + nfa*r]
// i.e. the operator table of the space contracted with the cell's dof vector.  The output layout is the FusiveTensor
// layout of a PER_POINT coefficient (fem/diff_tensor.h:112-124), so the result of afb_eval_quadrature feeds afb_assemble
// directly: this is the building block of re-assembly in nonlinear problems (coefficients depending on u_h, grad u_h).
// One thread per (tetrahedron, point); tables phi / G^ at the points are built on the host (afb_tables.cpp).
#include <algorithm>
#include <vector>

#include "afb_internal.h"

using namespace afb;

namespace {

struct EvalP {
    int op, vec, nf, q, dim;
    const double* phi;   // [q][nf]
    const double* grd;   // [q][nf][3]
    // geometry: SoA mesh (x != NULL) or 4 blocks 3 x f
    const double *x, *y, *z;
    const int32_t *v0, *v1, *v2, *v3;
    const double* XY[4];
    // dofs: batched [nfa x f] (u == NULL) or gathered from the global vector u through the dof table
    const double* dofs;
    const double* u;
    const int32_t* e2c;
    int col_off;
    long long ntet;
};

__global__ void __launch_bounds__(256) k_eval(EvalP P, double* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= P.ntet * P.q) return;
    const long long e = t / P.q;
    const int n = (int)(t - e * P.q);
    double PSI[9];
    if (P.op != AFB_IDEN) {
        double Pt[4][3];
        if (P.x) {
            const int nn[4] = {__ldg(P.v0 + e), __ldg(P.v1 + e), __ldg(P.v2 + e), __ldg(P.v3 + e)};
            for (int k = 0; k < 4; ++k) { Pt[k][0] = __ldg(P.x + nn[k]); Pt[k][1] = __ldg(P.y + nn[k]); Pt[k][2] = __ldg(P.z + nn[k]); }
        } else {
            for (int k = 0; k < 4; ++k)
                for (int d = 0; d < 3; ++d) Pt[k][d] = __ldg(P.XY[k] + 3 * e + d);
        }
        double m[9];
        for (int c = 0; c < 3; ++c)
            for (int i = 0; i < 3; ++i) m[i + 3 * c] = Pt[c + 1][i] - Pt[0][i];
        const double c00 = m[4] * m[8] - m[7] * m[5], c01 = m[7] * m[2] - m[1] * m[8], c02 = m[1] * m[5] - m[4] * m[2];
        const double id = 1.0 / (m[0] * c00 + m[3] * c01 + m[6] * c02);
        PSI[0] = c00 * id; PSI[1] = c01 * id; PSI[2] = c02 * id;
        PSI[3] = (m[6] * m[5] - m[3] * m[8]) * id; PSI[4] = (m[0] * m[8] - m[6] * m[2]) * id; PSI[5] = (m[3] * m[2] - m[0] * m[5]) * id;
        PSI[6] = (m[3] * m[7] - m[6] * m[4]) * id; PSI[7] = (m[6] * m[1] - m[0] * m[7]) * id; PSI[8] = (m[0] * m[4] - m[3] * m[1]) * id;
    }
    double* o = out + (size_t)P.dim * t;
    double div = 0.0;
    for (int c = 0; c < P.vec; ++c) {
        double s = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
        for (int i = 0; i < P.nf; ++i) {
            double d;
            if (P.u) {
                const int code = __ldg(P.e2c + (long long)(P.col_off + c * P.nf + i) * P.ntet + e);
                d = code == 0 ? 0.0 : (code > 0 ? __ldg(P.u + code - 1) : -__ldg(P.u - code - 1));
            } else d = __ldg(P.dofs + (size_t)(P.vec * P.nf) * e + c * P.nf + i);
            if (P.op == AFB_IDEN) s = fma(__ldg(P.phi + (size_t)n * P.nf + i), d, s);
            else {
                const double* G = P.grd + ((size_t)n * P.nf + i) * 3;
                g0 = fma(__ldg(G), d, g0); g1 = fma(__ldg(G + 1), d, g1); g2 = fma(__ldg(G + 2), d, g2);
            }
        }
        if (P.op == AFB_IDEN) o[c] = s;
        else {
            // physical gradient component k = sum_a PSI[a + 3k] * (reference gradient a)
            const double p0 = PSI[0] * g0 + PSI[1] * g1 + PSI[2] * g2, p1 = PSI[3] * g0 + PSI[4] * g1 + PSI[5] * g2,
                         p2 = PSI[6] * g0 + PSI[7] * g1 + PSI[8] * g2;
            if (P.op == AFB_GRAD) { o[3 * c] = p0; o[3 * c + 1] = p1; o[3 * c + 2] = p2; }
            else div += c == 0 ? p0 : (c == 1 ? p1 : p2);
        }
    }
    if (P.op == AFB_DIV) o[0] = div;
}

int eval_impl(afb_ctx* ctx, int op, int fem, int vec, int q, const double* XYL, int64_t f, const double* const* XY, const double* dofs,
              const double* u, int col_off, double* out, int mem_space) {
    OpInfo o;
    if (resolve_op(op, fem, vec, &o)) { set_error(ctx, "unsupported operator/space"); return -3; }
    if (f <= 0 || q <= 0) return 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const int nf = o.nf_base;
    std::vector<double> h((size_t)q * nf * 4);
    basis_values(fem, q, XYL, h.data());
    basis_ref_grads(fem, q, XYL, h.data() + (size_t)q * nf);
    AFB_CUDA(ctx, ctx->tables.reserve(h.size() * sizeof(double)));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tables.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    EvalP P;
    P.op = op; P.vec = vec; P.nf = nf; P.q = q; P.dim = o.dim;
    P.phi = ctx->tables.as<double>(); P.grd = P.phi + (size_t)q * nf;
    P.ntet = f; P.col_off = col_off; P.u = nullptr; P.dofs = nullptr; P.e2c = nullptr;
    P.x = P.y = P.z = nullptr; P.v0 = P.v1 = P.v2 = P.v3 = nullptr;
    for (int k = 0; k < 4; ++k) P.XY[k] = nullptr;
    const size_t nout = (size_t)o.dim * q * f;
    double* dout = out;
    if (XY) {  // batched form: coordinates and dofs from the caller
        const size_t ndofs = (size_t)o.nfa * f;
        if (mem_space == AFB_HOST) {
            AFB_CUDA(ctx, ctx->xy.reserve((size_t)12 * f * sizeof(double)));
            AFB_CUDA(ctx, ctx->tmp1.reserve(ndofs * sizeof(double)));
            AFB_CUDA(ctx, ctx->tmp3.reserve(nout * sizeof(double)));
            for (int k = 0; k < 4; ++k) {
                AFB_CUDA(ctx, cudaMemcpyAsync(ctx->xy.as<double>() + (size_t)3 * f * k, XY[k], (size_t)3 * f * sizeof(double), cudaMemcpyHostToDevice, st));
                P.XY[k] = ctx->xy.as<double>() + (size_t)3 * f * k;
            }
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp1.p, dofs, ndofs * sizeof(double), cudaMemcpyHostToDevice, st));
            P.dofs = ctx->tmp1.as<double>();
            dout = ctx->tmp3.as<double>();
        } else {
            for (int k = 0; k < 4; ++k) P.XY[k] = XY[k];
            P.dofs = dofs;
        }
    } else {   // mesh form: geometry and dof table of the context, dofs gathered from the global vector
        P.x = ctx->x.as<double>(); P.y = ctx->y.as<double>(); P.z = ctx->z.as<double>();
        P.v0 = ctx->v[0].as<int32_t>(); P.v1 = ctx->v[1].as<int32_t>(); P.v2 = ctx->v[2].as<int32_t>(); P.v3 = ctx->v[3].as<int32_t>();
        P.e2c = ctx->e2c.as<int32_t>();
        if (mem_space == AFB_HOST) {
            AFB_CUDA(ctx, ctx->tmp1.reserve((size_t)ctx->ncols_global * sizeof(double)));
            AFB_CUDA(ctx, ctx->tmp3.reserve(nout * sizeof(double)));
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp1.p, u, (size_t)ctx->ncols_global * sizeof(double), cudaMemcpyHostToDevice, st));
            P.u = ctx->tmp1.as<double>();
            dout = ctx->tmp3.as<double>();
        } else P.u = u;
    }
    const long long nthr = f * q;
    k_eval<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(P, dout);
    ctx->launches++;
    AFB_CUDA(ctx, cudaGetLastError());
    if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(out, dout, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

extern "C" {

int afb_fem3dapply_batched(afb_ctx* ctx, int op, int fem, int vec, int q, const double* XYL, int64_t f, const double* XY0, const double* XY1,
                           const double* XY2, const double* XY3, const double* dofs, double* opU, int mem_space) {
    if (!ctx) return -7;
    if (f <= 0 || q <= 0) return 0;
    if (!XYL || !XY0 || !XY1 || !XY2 || !XY3 || !dofs || !opU) { set_error(ctx, "afb_fem3dapply_batched: null buffer"); return -7; }
    const double* XY[4] = {XY0, XY1, XY2, XY3};
    return eval_impl(ctx, op, fem, vec, q, XYL, f, XY, dofs, nullptr, 0, opU, mem_space);
}

int afb_eval_quadrature(afb_ctx* ctx, int op, int fem, int vec, int col_off, int order, const double* u, double* out, int mem_space) {
    if (!ctx) return -7;
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified"); return -6; }
    if (!u || !out) { set_error(ctx, "afb_eval_quadrature: null buffer"); return -7; }
    const double *p, *w;
    const int q = tet_rule(order, &p, &w);
    if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
    OpInfo o;
    if (resolve_op(op, fem, vec, &o)) { set_error(ctx, "unsupported operator/space"); return -3; }
    if (col_off < 0 || col_off + o.nfa > ctx->ncol_loc) { set_error(ctx, "afb_eval_quadrature: variable outside the element vector"); return -7; }
    const int rc = eval_impl(ctx, op, fem, vec, q, p, ctx->ntet, nullptr, nullptr, u, col_off, out, mem_space);
    return rc ? rc : q;
}

}  // extern "C"
