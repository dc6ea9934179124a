// Element-matrix kernels (K1): batched replacement of Ani::fem3Dtet / internalFem3Dtet
// (reference: fem/operations/int_tet.inl:30-57, core.inl:217-367; tensor application
// fem/diff_tensor.h:276-994; contraction core.inl:28-60; band structure of vector spaces
// fem/operators.h:127-155, DIV :320-353).
//
// k_element_generic: one warp per tetrahedron, any operator/space/tensor combination.
//   geometry   -> every lane redundantly: edge vectors, PSI = inverse Jacobian, |T| (registers)
//   tables     -> physical base gradients Ub[d][n][a] = sum_j PSI[j + 3d] * G^[n][a][j]   (shared)
//   DU         -> DU[k][n][ia] = w_n |T| sum_j K(k,j) U[j,n,ia], band of ia only           (shared)
//   A(ib,ia)   -> sum_n sum_d Vb[d][n][b] * DU[band(ib)+d][n][ia], entries strided over lanes,
//                 accumulators in registers, quadrature points processed in shared-memory chunks
// The reference tables phi / G^ are element independent and come from afb_tables.cpp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "afb_internal.h"

namespace {

struct GeomSrc {
    const double *x, *y, *z;
    const int32_t *v0, *v1, *v2, *v3;
    const double* XY[4];  // 3 x f col-major each, used when x == nullptr
};

__device__ __forceinline__ void load_tet(const GeomSrc& g, long long e, double P[4][3]) {
    if (g.x) {
        const int n0 = __ldg(g.v0 + e), n1 = __ldg(g.v1 + e), n2 = __ldg(g.v2 + e), n3 = __ldg(g.v3 + e);
        const int nn[4] = {n0, n1, n2, n3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            P[k][0] = __ldg(g.x + nn[k]);
            P[k][1] = __ldg(g.y + nn[k]);
            P[k][2] = __ldg(g.z + nn[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            P[k][0] = __ldg(g.XY[k] + 3 * e + 0);
            P[k][1] = __ldg(g.XY[k] + 3 * e + 1);
            P[k][2] = __ldg(g.XY[k] + 3 * e + 2);
        }
    }
}

// PSI[j + 3k] = (M^-1)(j,k), M columns = P1-P0, P2-P0, P3-P0; returns det  (cofactor inverse)
__device__ __forceinline__ double jacobian_inverse(const double P[4][3], double PSI[9]) {
    double m[9];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i + 3 * c] = P[c + 1][i] - P[0][i];
    const double c00 = m[4] * m[8] - m[7] * m[5];
    const double c01 = m[7] * m[2] - m[1] * m[8];
    const double c02 = m[1] * m[5] - m[4] * m[2];
    const double det = m[0] * c00 + m[3] * c01 + m[6] * c02;
    const double id = 1.0 / det;
    // PSI[j + 3k] = inv(j,k) = cof(k,j)/det  (adjugate; same entries as geometry.h:25-45)
    PSI[0] = c00 * id;
    PSI[1] = c01 * id;
    PSI[2] = c02 * id;
    PSI[3] = (m[6] * m[5] - m[3] * m[8]) * id;
    PSI[4] = (m[0] * m[8] - m[6] * m[2]) * id;
    PSI[5] = (m[3] * m[2] - m[0] * m[5]) * id;
    PSI[6] = (m[3] * m[7] - m[6] * m[4]) * id;
    PSI[7] = (m[6] * m[1] - m[0] * m[7]) * id;
    PSI[8] = (m[0] * m[4] - m[3] * m[1]) * id;
    return det;
}

template <int ACC>
__global__ void __launch_bounds__(128) k_element_generic(FormDev F, long long f, GeomSrc g, double* __restrict__ out,
                                                         int ia0, int nia, int words_per_warp) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (e >= f) return;
    double* sm = smem + (size_t)wib * words_per_warp;
    const bool gradA = F.opA != AFB_IDEN, gradB = F.opB != AFB_IDEN;
    double* UbA = sm;                                            // [dbA][qc][nfbA] if gradA
    double* UbB = UbA + (gradA ? 3 * F.qc * F.nfbA : 0);         // [dbB][qc][nfbB] if gradB && !same
    double* DU = UbB + ((gradB && !F.same) ? 3 * F.qc * F.nfbB : 0);  // [jdim][qc][nia]
    if (F.same) UbB = UbA;

    double P[4][3], PSI[9];
    load_tet(g, F.item_tet ? (long long)__ldg(F.item_tet + e) : e, P);
    const double det = jacobian_inverse(P, PSI);
    double vol = fabs(det) * (1.0 / 6.0);
    const double *phiA = F.phiA, *grdA = F.grdA, *phiB = F.phiB, *grdB = F.grdB;
    if (F.face) {
        // fem3Dface (int_face.inl:160-199): tables of the triangle rule lifted to face fc, measure = area of the face, computed
        // from the vertices relative to P0 like the reference's XYP (core.inl:231-240, geometry.h:67-72)
        const int fc = __ldg(F.face + e) & 3;
        phiA += (size_t)fc * F.q * F.nfbA; grdA += (size_t)3 * fc * F.q * F.nfbA;
        phiB += (size_t)fc * F.q * F.nfbB; grdB += (size_t)3 * fc * F.q * F.nfbB;
        const int i0 = fc, i1 = (fc + 1) & 3, i2 = (fc + 2) & 3;
        double a[3], b[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            // selects instead of P[i][d] with a run-time i: the coordinates stay in registers
            auto pick = [&](int i) { return i == 0 ? P[0][d] : (i == 1 ? P[1][d] : (i == 2 ? P[2][d] : P[3][d])); };
            const double x0 = pick(i0) - P[0][d], x1 = pick(i1) - P[0][d], x2 = pick(i2) - P[0][d];
            a[d] = x0 - x2; b[d] = x1 - x2;
        }
        const double c0 = a[1] * b[2] - a[2] * b[1], c1 = -a[0] * b[2] + a[2] * b[0], c2 = a[0] * b[1] - a[1] * b[0];
        vol = sqrt(c0 * c0 + c1 * c1 + c2 * c2) * 0.5;
    }

    double acc[ACC];
#pragma unroll
    for (int t = 0; t < ACC; ++t) acc[t] = 0.0;
    const int nent = F.nfb * nia;

    for (int n0 = 0; n0 < F.q; n0 += F.qc) {
        const int qn = min(F.qc, F.q - n0);
        // ---- physical base gradients
        if (gradA) {
            for (int it = lane; it < qn * F.nfbA; it += 32) {
                const int nl = it / F.nfbA, a = it - nl * F.nfbA;
                const double* G = grdA + ((size_t)(n0 + nl) * F.nfbA + a) * 3;
                const double g0 = __ldg(G), g1 = __ldg(G + 1), g2 = __ldg(G + 2);
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    UbA[(d * F.qc + nl) * F.nfbA + a] = PSI[0 + 3 * d] * g0 + PSI[1 + 3 * d] * g1 + PSI[2 + 3 * d] * g2;
            }
        }
        if (gradB && !F.same) {
            for (int it = lane; it < qn * F.nfbB; it += 32) {
                const int nl = it / F.nfbB, b = it - nl * F.nfbB;
                const double* G = grdB + ((size_t)(n0 + nl) * F.nfbB + b) * 3;
                const double g0 = __ldg(G), g1 = __ldg(G + 1), g2 = __ldg(G + 2);
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    UbB[(d * F.qc + nl) * F.nfbB + b] = PSI[0 + 3 * d] * g0 + PSI[1 + 3 * d] * g1 + PSI[2 + 3 * d] * g2;
            }
        }
        __syncwarp();
        // ---- DU[k][nl][ial] = w vol sum_j K(k,j) U[j]
        for (int it = lane; it < F.jdim * qn * nia; it += 32) {
            const int ial = it % nia;
            const int r2 = it / nia;
            const int nl = r2 % qn, k = r2 / qn;
            const int ia = ia0 + ial;
            const int c = ia / F.nfbA, a = ia - c * F.nfbA;
            const int n = n0 + nl;
            const double wv = __ldg(F.W + n) * vol;
            const double* Dn = F.D;
            if (F.layout == AFB_COEF_PER_TET) Dn += (size_t)F.dlen * e;
            else if (F.layout == AFB_COEF_PER_POINT) Dn += (size_t)F.dlen * (n + (size_t)F.q * e);
            // band of ia: U[j0 + d], d < nb
            int j0, nb;
            if (F.opA == AFB_IDEN) { j0 = c; nb = 1; }
            else if (F.opA == AFB_GRAD) { j0 = 3 * c; nb = 3; }
            else { j0 = 0; nb = 1; }
            double u[3];
            if (F.opA == AFB_IDEN) u[0] = __ldg(phiA + (size_t)n * F.nfbA + a);
            else if (F.opA == AFB_GRAD) {
#pragma unroll
                for (int d = 0; d < 3; ++d) u[d] = UbA[(d * F.qc + nl) * F.nfbA + a];
            } else u[0] = UbA[(c * F.qc + nl) * F.nfbA + a];
            double s;
            if (F.ttype >= AFB_TENSOR_SYMMETRIC) {
                s = 0.0;
                for (int d = 0; d < nb; ++d) s += __ldg(Dn + k + F.jdim * (j0 + d)) * u[d];
            } else {
                const double sc = (F.ttype == AFB_TENSOR_SCALAR) ? __ldg(Dn) : 1.0;
                if (F.jdim == F.idim) s = (k >= j0 && k < j0 + nb) ? sc * u[k - j0] : 0.0;
                else s = sc * u[0];  // OpA = IDEN(P0) broadcast (rhs trick)
            }
            DU[(k * F.qc + nl) * nia + ial] = wv * s;
        }
        __syncwarp();
        // ---- A(ib, ia) += sum_n sum_d Vb * DU
#pragma unroll
        for (int t = 0; t < ACC; ++t) {
            const int idx = lane + 32 * t;
            if (idx < nent) {
                const int ib = idx / nia, ial = idx - ib * nia;
                const int cb = ib / F.nfbB, b = ib - cb * F.nfbB;
                double s = acc[t];
                if (F.opB == AFB_IDEN) {
                    for (int nl = 0; nl < qn; ++nl)
                        s += __ldg(phiB + (size_t)(n0 + nl) * F.nfbB + b) * DU[(cb * F.qc + nl) * nia + ial];
                } else if (F.opB == AFB_GRAD) {
                    for (int nl = 0; nl < qn; ++nl)
#pragma unroll
                        for (int d = 0; d < 3; ++d)
                            s += UbB[(d * F.qc + nl) * F.nfbB + b] * DU[((3 * cb + d) * F.qc + nl) * nia + ial];
                } else {
                    for (int nl = 0; nl < qn; ++nl) s += UbB[(cb * F.qc + nl) * F.nfbB + b] * DU[nl * nia + ial];
                }
                acc[t] = s;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int t = 0; t < ACC; ++t) {
        const int idx = lane + 32 * t;
        if (idx < nent) {
            const int ib = idx / nia, ial = idx - ib * nia;
            double* o = out + e * F.s_e + (long long)(F.row_off + ib) * F.s_ib + (long long)(F.col_off + ia0 + ial) * F.s_ia;
            const double v = F.alpha * acc[t];
            if (F.add) *o += v; else *o = v;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// k_element_sq: register-tiled element kernel for the contraction-bound case: all forms of one Assemble that
// are square on one scalar space with the same operator on both sides,
//     A(i,j) = sum_forms sum_n sum_d V[d][n][i] * DU[d][n][j],   V = GRAD or IDEN of P1/P2/P3,  DU = w_n |T| alpha K_n V
// with any tensor kind and layout (CONST / PER_TET / PER_POINT).  Same arithmetic as internalFem3Dtet + fusive_AT_mul_B
// (fem/operations/core.inl:28-60,277-367), organised as a rank-1-update GEMM: one warp per tetrahedron, lane (li, lj) of an
// 8 x 4 grid owns the entries i = li + 8t, j = lj + 4u in registers (3 x 5 for P3), V and DU of a chunk of quadrature
// points sit in shared memory, and every quadrature point / direction costs TR + TC shared loads for TR*TC DFMAs per lane
// (the generic kernel needs two loads per DFMA).  All forms accumulate into the same registers: the staged element matrix
// is written once.
constexpr int SQ_MAXF = 4;
constexpr int SQ_QC = 8;  // quadrature points per shared-memory chunk
struct SqForm {
    int grad;            // 1: GRAD x GRAD, 0: IDEN x IDEN
    int q;
    const double *W, *phi, *grd;
    int ttype, layout, dlen;
    const double* D;
    double alpha;
};
struct SqParams { int nforms; SqForm f[SQ_MAXF]; };

template <int NF>
__global__ void __launch_bounds__(128) k_element_sq(SqParams P, long long ntet, GeomSrc g, double* __restrict__ out, long long s_e, int s_i, int s_j) {
    constexpr int TR = (NF + 7) / 8, TC = (NF + 3) / 4;
    __shared__ double sU[4][3 * SQ_QC * NF];
    __shared__ double sDU[4][3 * SQ_QC * NF];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * 4 + wib;
    if (e >= ntet) return;
    double* U = sU[wib];
    double* DU = sDU[wib];
    const int li = lane >> 2, lj = lane & 3;
    double Pt[4][3], PSI[9];
    load_tet(g, e, Pt);
    const double vol = fabs(jacobian_inverse(Pt, PSI)) * (1.0 / 6.0);
    double acc[TR][TC];
#pragma unroll
    for (int t = 0; t < TR; ++t)
#pragma unroll
        for (int u = 0; u < TC; ++u) acc[t][u] = 0.0;

    for (int fi = 0; fi < P.nforms; ++fi) {
        const SqForm& F = P.f[fi];
        const int nd = F.grad ? 3 : 1;
        for (int n0 = 0; n0 < F.q; n0 += SQ_QC) {
            const int qn = min(SQ_QC, F.q - n0);
            // ---- V and DU of this chunk
            for (int it = lane; it < qn * NF; it += 32) {
                const int nl = it / NF, a = it - nl * NF;
                const int n = n0 + nl;
                const double wv = __ldg(F.W + n) * vol * F.alpha;
                const double* Dn = F.D;
                if (F.layout == AFB_COEF_PER_TET) Dn += (size_t)F.dlen * e;
                else if (F.layout == AFB_COEF_PER_POINT) Dn += (size_t)F.dlen * (n + (size_t)F.q * e);
                if (F.grad) {
                    const double* G = F.grd + ((size_t)n * NF + a) * 3;
                    const double g0 = __ldg(G), g1 = __ldg(G + 1), g2 = __ldg(G + 2);
                    double u[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) u[d] = PSI[0 + 3 * d] * g0 + PSI[1 + 3 * d] * g1 + PSI[2 + 3 * d] * g2;
#pragma unroll
                    for (int d = 0; d < 3; ++d) U[(d * SQ_QC + nl) * NF + a] = u[d];
                    if (F.ttype >= AFB_TENSOR_SYMMETRIC) {
#pragma unroll
                        for (int k = 0; k < 3; ++k)  // DU[k] = sum_l K(k,l) u[l], K(k,l) at k + 3l
                            DU[(k * SQ_QC + nl) * NF + a] = wv * (__ldg(Dn + k) * u[0] + __ldg(Dn + k + 3) * u[1] + __ldg(Dn + k + 6) * u[2]);
                    } else {
                        const double c = wv * (F.ttype == AFB_TENSOR_SCALAR ? __ldg(Dn) : 1.0);
#pragma unroll
                        for (int d = 0; d < 3; ++d) DU[(d * SQ_QC + nl) * NF + a] = c * u[d];
                    }
                } else {
                    const double ph = __ldg(F.phi + (size_t)n * NF + a);
                    const double c = wv * (F.ttype >= AFB_TENSOR_SCALAR ? __ldg(Dn) : 1.0);
                    U[nl * NF + a] = ph;
                    DU[nl * NF + a] = c * ph;
                }
            }
            __syncwarp();
            // ---- rank-1 updates
            for (int d = 0; d < nd; ++d)
                for (int nl = 0; nl < qn; ++nl) {
                    const double* ur = U + (d * SQ_QC + nl) * NF;
                    const double* dr = DU + (d * SQ_QC + nl) * NF;
                    double vr[TR], dc[TC];
#pragma unroll
                    for (int t = 0; t < TR; ++t) vr[t] = (li + 8 * t < NF) ? ur[li + 8 * t] : 0.0;
#pragma unroll
                    for (int u = 0; u < TC; ++u) dc[u] = (lj + 4 * u < NF) ? dr[lj + 4 * u] : 0.0;
#pragma unroll
                    for (int t = 0; t < TR; ++t)
#pragma unroll
                        for (int u = 0; u < TC; ++u) acc[t][u] = fma(vr[t], dc[u], acc[t][u]);
                }
            __syncwarp();
        }
    }
    double* o = out + e * s_e;
#pragma unroll
    for (int t = 0; t < TR; ++t)
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const int i = li + 8 * t, j = lj + 4 * u;
            if (i < NF && j < NF) o[i * s_i + j * s_j] = acc[t][u];   // (NF, 1): staged row-major; (1, NF): fem3Dtet's column-major block
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_element_mma: the same contraction as k_element_sq on the FP64 tensor path (north_star: "DMMA ... only for the high-order or
// vector-valued spaces where ncu shows the contraction dominates").  k_element_sq was bound by the shared-memory operand
// loads of its rank-1 updates (LSU 73 %, FP64 pipe 35 %, profiles/r01d_c3_c4_kernels.md): with mma.sync.m8n8k4.f64 one operand
// pair per lane feeds 256 FMAs, and the operands never pass through shared memory at all:
//     A_e(i,j) = sum_k V[k][i] * DU[k][j],   k = (quadrature point n, direction d)
// is a sum of 8x8 tiles D += A(8x4) B(4x8) with A[i][k] = V[k][i] (row operand: lane (gid, tig) holds V[n0+tig][8ti+gid]) and
// B[k][j] = DU[k][j] (column operand: the same lane holds DU[n0+tig][8tj+gid]).  Both operands of a lane belong to the SAME
// (point, basis function) pairs, so the lane computes its physical gradients u = PSI^T grad_ref (3 coalesced table loads from the
// lane-ordered shared-memory copy of the reference table, 9 DFMA) and its DU = w_n |T| K_n u once per group of four points and
// feeds them to the tiles of 3 directions x NT x NT (symmetric forms: upper tiles only, mirrored at the store).
// One warp per tetrahedron, grid-stride; accumulators: 2 registers per tile.
template <int NF, bool SYM>
__global__ void __launch_bounds__(128) k_element_mma(SqParams P, long long ntet, GeomSrc g, double* __restrict__ out, long long s_e, int s_i, int s_j) {
    constexpr int NT = (NF + 7) / 8;
    extern __shared__ double smt[];   // per form: weights [groups][32], then grad [groups][NT][3][32] or iden [groups][NT][32]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    int tab_off[SQ_MAXF];
    {
        int off = 0;
        for (int fi = 0; fi < P.nforms; ++fi) {
            const SqForm& F = P.f[fi];
            const int groups = (F.q + 3) / 4;
            tab_off[fi] = off;
            const int per = F.grad ? NT * 3 : NT;
            for (int it = threadIdx.x; it < groups * 32; it += blockDim.x) {
                const int n = (it >> 5) * 4 + (it & 3);
                smt[off + it] = n < F.q ? __ldg(F.W + n) * F.alpha : 0.0;
            }
            double* tb = smt + off + groups * 32;
            for (int it = threadIdx.x; it < groups * per * 32; it += blockDim.x) {
                const int l32 = it & 31, r = it >> 5;
                const int gg = r / per, rr = r - gg * per;
                const int n = gg * 4 + (l32 & 3);
                double v = 0.0;
                if (F.grad) {
                    const int t = rr / 3, l = rr - 3 * t, i = 8 * t + (l32 >> 2);
                    if (n < F.q && i < NF) v = __ldg(F.grd + ((size_t)n * NF + i) * 3 + l);
                } else {
                    const int i = 8 * rr + (l32 >> 2);
                    if (n < F.q && i < NF) v = __ldg(F.phi + (size_t)n * NF + i);
                }
                tb[it] = v;
            }
            off += groups * 32 * (1 + per);
        }
    }
    __syncthreads();
    for (long long e = (long long)blockIdx.x * 4 + wib; e < ntet; e += (long long)gridDim.x * 4) {
        double Pt[4][3], PSI[9];
        load_tet(g, e, Pt);
        const double vol = fabs(jacobian_inverse(Pt, PSI)) * (1.0 / 6.0);
        double c0[NT][NT], c1[NT][NT];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) { c0[a][b] = 0.0; c1[a][b] = 0.0; }
        for (int fi = 0; fi < P.nforms; ++fi) {
            const SqForm& F = P.f[fi];
            const int groups = (F.q + 3) / 4;
            const double* wt = smt + tab_off[fi];
            const double* tb = wt + groups * 32;
            // Coefficients are fetched ahead of their use: the scalar coefficients of four groups of points at once, a tensor one
            // group ahead (double buffer).  [Loaded inside the group they were the top stall: 35 % of the samples waited for the
            // global load in front of the DMUL w_n |T| c_n, profiles/r02/r02d_element_mma.md.]
            const bool tens = F.ttype >= AFB_TENSOR_SYMMETRIC && F.grad;
            const bool scal = F.ttype >= AFB_TENSOR_SCALAR && !tens;
            const size_t pstride = F.layout == AFB_COEF_PER_POINT ? (size_t)F.dlen : 0;
            const double* Dbase = F.D;
            if (F.layout == AFB_COEF_PER_TET) Dbase += (size_t)F.dlen * e;
            else if (F.layout == AFB_COEF_PER_POINT) Dbase += (size_t)F.dlen * F.q * e;
            // padding points carry weight 0; their coefficient is read from the last point
            auto dptr = [&](int gg) { return Dbase + pstride * (size_t)min(gg * 4 + tig, F.q - 1); };
            double Kn[9];
            if (tens) {
                const double* Dn = dptr(0);
#pragma unroll
                for (int k = 0; k < 9; ++k) Kn[k] = __ldg(Dn + k);
            }
            for (int g0 = 0; g0 < groups; g0 += 4) {
                double cf[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) cf[k] = (scal && g0 + k < groups) ? __ldg(dptr(g0 + k)) : 1.0;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int gg = g0 + k4;
                    if (gg >= groups) break;
                    const double wv = wt[gg * 32 + lane] * vol;
                    if (F.grad) {
                        double u[NT][3], du[NT][3];
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            const double g0v = tb[((gg * NT + t) * 3 + 0) * 32 + lane], g1v = tb[((gg * NT + t) * 3 + 1) * 32 + lane],
                                         g2v = tb[((gg * NT + t) * 3 + 2) * 32 + lane];
#pragma unroll
                            for (int d = 0; d < 3; ++d) u[t][d] = PSI[0 + 3 * d] * g0v + PSI[1 + 3 * d] * g1v + PSI[2 + 3 * d] * g2v;
                        }
                        if (tens) {
                            double K[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) K[k] = wv * Kn[k];
                            if (gg + 1 < groups) {
                                const double* Dn = dptr(gg + 1);
#pragma unroll
                                for (int k = 0; k < 9; ++k) Kn[k] = __ldg(Dn + k);
                            }
#pragma unroll
                            for (int t = 0; t < NT; ++t)
#pragma unroll
                                for (int k = 0; k < 3; ++k) du[t][k] = K[k] * u[t][0] + K[k + 3] * u[t][1] + K[k + 6] * u[t][2];   // K(k,l) at k + 3l
                        } else {
                            const double c = wv * cf[k4];
#pragma unroll
                            for (int t = 0; t < NT; ++t)
#pragma unroll
                                for (int d = 0; d < 3; ++d) du[t][d] = c * u[t][d];
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int a = 0; a < NT; ++a)
#pragma unroll
                                for (int b = SYM ? a : 0; b < NT; ++b)
                                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                                 : "+d"(c0[a][b]), "+d"(c1[a][b]) : "d"(u[a][d]), "d"(du[b][d]));
                    } else {
                        const double c = wv * cf[k4];
                        double ph[NT], dp[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) { ph[t] = tb[(gg * NT + t) * 32 + lane]; dp[t] = c * ph[t]; }
#pragma unroll
                        for (int a = 0; a < NT; ++a)
#pragma unroll
                            for (int b = SYM ? a : 0; b < NT; ++b)
                                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                             : "+d"(c0[a][b]), "+d"(c1[a][b]) : "d"(ph[a]), "d"(dp[b]));
                    }
                }
            }
        }
        // tile (a, b): this lane holds (i, j) and (i, j + 1) with i = 8a + gid, j = 8b + 2 tig
        double* o = out + e * s_e;
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = SYM ? a : 0; b < NT; ++b) {
                const int i = 8 * a + gid, j = 8 * b + 2 * tig;
                if (i < NF && j < NF) {
                    o[i * s_i + j * s_j] = c0[a][b];
                    if (j + 1 < NF) o[i * s_i + (j + 1) * s_j] = c1[a][b];
                    if (SYM && a != b) {
                        o[j * s_i + i * s_j] = c0[a][b];
                        if (j + 1 < NF) o[(j + 1) * s_i + i * s_j] = c1[a][b];
                    }
                }
            }
    }
}

template <int ACC>
cudaError_t launch_generic(const FormDev& F, long long f, const GeomSrc& g, double* out, int ia0, int nia, int words, cudaStream_t st) {
    const int warps = 4;
    const size_t smem = (size_t)warps * words * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_element_generic<ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long blocks = (f + warps - 1) / warps;
    k_element_generic<ACC><<<(unsigned)blocks, warps * 32, smem, st>>>(F, f, g, out, ia0, nia, words);
    return cudaGetLastError();
}

}  // namespace

namespace afb {

int form_dlen(const afb_form& form, const OpInfo& A, const OpInfo& B) {
    if (form.tensor_type == AFB_TENSOR_NULL) return 0;
    if (form.tensor_type == AFB_TENSOR_SCALAR) return 1;
    return A.dim * B.dim;
}

int launch_form(afb_ctx* ctx, const afb_form& form, const OpInfo& A, const OpInfo& B, int64_t f,
                const double* x, const double* y, const double* z,
                const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                const double* XY, double* out, long long s_e, long long s_ib, long long s_ia, int add,
                const double* Ddev, const int32_t* face, const int32_t* item_tet) {
    if (f <= 0) return 0;
    const double *p, *w;
    const int q = face ? tri_rule(form.quad_order, &p, &w) : tet_rule(form.quad_order, &p, &w);
    if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
    const int tt = form.tensor_type;
    if (tt < AFB_TENSOR_NULL || tt > AFB_TENSOR_GENERAL) { set_error(ctx, "bad tensor_type"); return -7; }
    if ((tt == AFB_TENSOR_NULL || tt == AFB_TENSOR_SCALAR) && A.dim != B.dim && !(A.nfa == 1 && A.dim == 1)) {
        set_error(ctx, "Identity tensor defined only for compatible (with same dimensions) operators A and B");
        return -5;
    }
    if (tt == AFB_TENSOR_SYMMETRIC && A.dim != B.dim) { set_error(ctx, "TENSOR_SYMMETRIC should have equal dimensions"); return -5; }
    if (form.coef_layout < AFB_COEF_CONST || form.coef_layout > AFB_COEF_PER_POINT) { set_error(ctx, "bad coef_layout"); return -7; }
    // element-independent tables, cached on the device per (space, rule)
    const double *dW, *dphiA, *dgrdA, *dphiB, *dgrdB;
    int rc = get_tables(ctx, A.fem, form.quad_order, &dW, &dphiA, &dgrdA, face != nullptr);
    if (rc) return rc;
    rc = get_tables(ctx, B.fem, form.quad_order, &dW, &dphiB, &dgrdB, face != nullptr);
    if (rc) return rc;

    FormDev F;
    F.opA = A.op; F.vecA = A.vec; F.nfbA = A.nf_base; F.nfa = A.nfa; F.idim = A.dim; F.dbA = A.dim_base;
    F.opB = B.op; F.vecB = B.vec; F.nfbB = B.nf_base; F.nfb = B.nfa; F.jdim = B.dim; F.dbB = B.dim_base;
    F.same = (A.op == B.op && A.fem == B.fem && A.vec == B.vec);
    F.q = q;
    F.W = dW; F.phiA = dphiA; F.grdA = dgrdA; F.phiB = dphiB; F.grdB = dgrdB;
    F.ttype = tt; F.layout = form.coef_layout; F.dlen = form_dlen(form, A, B);
    F.D = Ddev;
    F.alpha = form.alpha;
    F.s_e = s_e; F.s_ib = s_ib; F.s_ia = s_ia; F.row_off = form.row_off; F.col_off = form.col_off; F.add = add;
    F.face = face; F.item_tet = item_tet;

    GeomSrc g;
    g.x = x; g.y = y; g.z = z; g.v0 = v0; g.v1 = v1; g.v2 = v2; g.v3 = v3;
    for (int k = 0; k < 4; ++k) g.XY[k] = XY ? XY + (size_t)3 * f * k : nullptr;

    // split the trial dofs so that nfb*nia fits the accumulator budget (29 per lane)
    const int max_ent = 32 * 29;
    int nia_max = A.nfa;
    while (B.nfa * nia_max > max_ent) nia_max = (nia_max + 1) / 2;
    if (nia_max < 1) { set_error(ctx, "element block too large"); return -3; }
    const int budget_words = 3072;  // 24 KiB of shared memory per warp
    for (int ia0 = 0; ia0 < A.nfa; ia0 += nia_max) {
        const int nia = std::min(nia_max, A.nfa - ia0);
        const int per_pt = (A.op != AFB_IDEN ? 3 * A.nf_base : 0) + ((B.op != AFB_IDEN && !F.same) ? 3 * B.nf_base : 0) + B.dim * nia;
        int qc = std::min(q, std::max(1, budget_words / per_pt));
        if (per_pt > budget_words) { set_error(ctx, "element block does not fit shared memory"); return -3; }
        F.qc = qc;
        const int words = per_pt * qc;
        const int nent = B.nfa * nia;
        const int acc = (nent + 31) / 32;
        cudaError_t e;
        if (acc <= 1) e = launch_generic<1>(F, f, g, out, ia0, nia, words, ctx->stream);
        else if (acc <= 2) e = launch_generic<2>(F, f, g, out, ia0, nia, words, ctx->stream);
        else if (acc <= 4) e = launch_generic<4>(F, f, g, out, ia0, nia, words, ctx->stream);
        else if (acc <= 7) e = launch_generic<7>(F, f, g, out, ia0, nia, words, ctx->stream);
        else if (acc <= 13) e = launch_generic<13>(F, f, g, out, ia0, nia, words, ctx->stream);
        else if (acc <= 19) e = launch_generic<19>(F, f, g, out, ia0, nia, words, ctx->stream);
        else e = launch_generic<29>(F, f, g, out, ia0, nia, words, ctx->stream);
        ctx->launches++;
        if (e != cudaSuccess) return cuda_fail(ctx, e, "k_element_generic launch");
    }
    return 0;
}

// Launches k_element_sq for the forms sel[] (all square on one scalar space, same operator on both sides, full local
// matrix) on the elements [e_lo, e_lo + ntet): out[(e - e_lo)*s_e + i*nf + j] = sum of the selected forms (store); Dd already points
// at the data of element e_lo.  Returns 0 ok, < 0 error.
int launch_forms_sq(afb_ctx* ctx, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa, const std::vector<const double*>& Dd,
                    const std::vector<int>& sel, int64_t e_lo, int64_t ntet, double* out, long long s_e, const double* XY, int colmajor) {
    if (sel.empty() || (int)sel.size() > SQ_MAXF) return -7;
    SqParams P;
    P.nforms = (int)sel.size();
    const int nf = oa[sel[0]].nf_base;
    for (size_t k = 0; k < sel.size(); ++k) {
        const afb_form& f = fm[sel[k]];
        const OpInfo& A = oa[sel[k]];
        SqForm& F = P.f[k];
        const double *p, *w;
        F.q = tet_rule(f.quad_order, &p, &w);
        if (F.q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
        int rc = get_tables(ctx, A.fem, f.quad_order, &F.W, &F.phi, &F.grd);
        if (rc) return rc;
        F.grad = A.op == AFB_GRAD;
        F.ttype = f.tensor_type; F.layout = f.coef_layout; F.dlen = form_dlen(f, A, A);
        F.D = Dd[sel[k]]; F.alpha = f.alpha;
    }
    GeomSrc g;
    g.x = ctx->x.as<double>(); g.y = ctx->y.as<double>(); g.z = ctx->z.as<double>();
    g.v0 = ctx->v[0].as<int32_t>() + e_lo; g.v1 = ctx->v[1].as<int32_t>() + e_lo; g.v2 = ctx->v[2].as<int32_t>() + e_lo; g.v3 = ctx->v[3].as<int32_t>() + e_lo;
    for (int k = 0; k < 4; ++k) g.XY[k] = nullptr;
    if (XY) {  // batched fem3Dtet call: 4 blocks 3 x f instead of the context's mesh
        g.x = g.y = g.z = nullptr;
        for (int k = 0; k < 4; ++k) g.XY[k] = XY + (size_t)3 * ntet * k;
    }
    const int s_i = colmajor ? 1 : nf, s_j = colmajor ? nf : 1;
    const unsigned blocks = (unsigned)((ntet + 3) / 4);
    // P2 / P3: FP64 tensor path (k_element_mma).  Symmetric instantiation when every form promises a symmetric product
    // (no tensor, scalar, or TENSOR_SYMMETRIC = "symmetric tensor, may be used to optimize calculations", diff_tensor.h:20).
    if ((nf == 10 || nf == 20) && !getenv("AFB_DISABLE_MMA_ELEMENT")) {
        bool sym = !getenv("AFB_MMA_NOSYM");
        size_t words = 0;
        for (int k = 0; k < P.nforms; ++k) {
            if (P.f[k].ttype == AFB_TENSOR_GENERAL) sym = false;
            const int groups = (P.f[k].q + 3) / 4, nt = (nf + 7) / 8;
            words += (size_t)groups * 32 * (1 + (P.f[k].grad ? 3 * nt : nt));
        }
        const size_t smem = words * sizeof(double);
        if (smem <= 200 * 1024) {
            const unsigned grid = (unsigned)std::min<long long>(blocks, 148LL * 16);
#define LAUNCH_MMA(NFV, S)                                                                                                         \
    do {                                                                                                                           \
        cudaError_t ea = cudaFuncSetAttribute(k_element_mma<NFV, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (ea != cudaSuccess) return cuda_fail(ctx, ea, "cudaFuncSetAttribute(k_element_mma)");                                   \
        k_element_mma<NFV, S><<<grid, 128, smem, ctx->stream>>>(P, ntet, g, out, s_e, s_i, s_j);                                   \
    } while (0)
            if (nf == 10) { if (sym) LAUNCH_MMA(10, true); else LAUNCH_MMA(10, false); }
            else { if (sym) LAUNCH_MMA(20, true); else LAUNCH_MMA(20, false); }
#undef LAUNCH_MMA
            ctx->launches++;
            ctx->k_elem = "k_element_mma";
            cudaError_t em = cudaGetLastError();
            if (em != cudaSuccess) return cuda_fail(ctx, em, "k_element_mma launch");
            return 0;
        }
    }
    if (nf == 4) k_element_sq<4><<<blocks, 128, 0, ctx->stream>>>(P, ntet, g, out, s_e, s_i, s_j);
    else if (nf == 10) k_element_sq<10><<<blocks, 128, 0, ctx->stream>>>(P, ntet, g, out, s_e, s_i, s_j);
    else if (nf == 20) k_element_sq<20><<<blocks, 128, 0, ctx->stream>>>(P, ntet, g, out, s_e, s_i, s_j);
    else return -7;
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_element_sq launch");
    ctx->k_elem = "k_element_sq";
    return 0;
}

}  // namespace afb
