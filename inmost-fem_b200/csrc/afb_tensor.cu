// Fused fast path of afb_assemble for affine elements with element-wise constant coefficients:
// "tensor representation" of the element matrix, expanded inside the row gather.
//
// Reference semantics being reproduced (AniFem++): fem3Dtet<Operator<GRAD|IDEN,FemFix<P0..P3>>, ...> with a
// CONST / per-tetrahedron tensor (fem/operations/core.inl:277-367, fem/diff_tensor.h:276-549) summed over the forms
// of the user's local assembler and scattered by AssemblerT::Assemble (inmost_interface/assembler.inl:397-481).
//
// Algebra.  With the coefficient K constant on a tet, physical gradients (grad phi)_k = sum_a PSI[a+3k] G^_a and
// |T| constant, the reference's quadrature sum factors exactly (for ANY rule, no exactness assumption):
//   A_e(i,j) = sum_n w_n |T| (K grad phi_j).grad phi_i = sum_{ab} M_e[ab] * S^{ab}_{ij},
//   M_e[ab]  = |T| sum_{kl} PSI[a+3k] K(k,l) PSI[b+3l]          (per element: 6 or 9 doubles)
//   S^{ab}_{ij} = sum_n w_n G^B_{ia}(n) G^A_{jb}(n)             (element independent, built once on the host)
// and likewise IDEN x IDEN (1 number per element), GRAD x IDEN / IDEN x GRAD (3 numbers), and the rhs trick.
// All forms of one Assemble are concatenated: A_e(i,j) = sum_alpha T[alpha][i][j] * g_e[alpha].
//
// Kernels.
//   k_geom          one thread per tet: coordinates + coefficient -> g_e[alpha] (coalesced 16-byte stores).  This
//                   replaces the 100-entry element matrix that the generic path stages through HBM by <= 8 doubles.
//   k_gather_tensor a group of G lanes owns a CSR row, walks the row's (element, local row) adjacency in ascending
//                   element order, lane j evaluates A_e(i,j) from g_e (broadcast load) and the shared-memory table T,
//                   and adds it into the shared-memory row image at the precomputed slot; the row is written once.
//                   Deterministic, no atomics.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "afb_internal.h"
#include "afb_ring_plan.h"

using namespace afb;

namespace {

constexpr int MAX_TFORMS = 8;

struct TFormDev {
    int kind;      // 0 GRADxGRAD, 1 IDENxIDEN, 2 GRAD(A)xIDEN(B), 3 IDEN(A)xGRAD(B)
    int full;      // kind 0: 1 = a 3x3 tensor K is given (9 values through kidx), 0 = c * identity (value kidx[0])
    int layout;    // CONST / PER_TET
    int dstride;   // doubles per element record of D (PER_TET)
    int kidx[9];   // position of every tensor value inside the record; -1 = the value 0, -2 = the value 1
    int goff, ng;  // components [goff, goff+ng) of g_e
    int bary;      // kind 0, ng 6: store the off-diagonal barycentric coefficients (G01,G23,G02,G13,G03,G12) of the ring kernel instead of M
    double alpha;
    const double* D;
};
struct GeomParams {
    int nforms;
    int ngpad;  // doubles per element in gbuf (even)
    int ngtot;  // components actually used; the pad slot is zeroed
    int zero_all;  // records have fixed slots some of which no form fills (ring kernel): zero the whole record first
    TFormDev f[MAX_TFORMS];
};

__device__ __forceinline__ double jac_inv(const double P[4][3], double PSI[9]) {
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
    PSI[0] = c00 * id; PSI[1] = c01 * id; PSI[2] = c02 * id;
    PSI[3] = (m[6] * m[5] - m[3] * m[8]) * id;
    PSI[4] = (m[0] * m[8] - m[6] * m[2]) * id;
    PSI[5] = (m[3] * m[2] - m[0] * m[5]) * id;
    PSI[6] = (m[3] * m[7] - m[6] * m[4]) * id;
    PSI[7] = (m[6] * m[1] - m[0] * m[7]) * id;
    PSI[8] = (m[0] * m[4] - m[3] * m[1]) * id;
    return det;
}

__device__ __forceinline__ double coef_value(const double* __restrict__ D, int idx) {
    return idx >= 0 ? __ldg(D + idx) : (idx == -2 ? 1.0 : 0.0);
}

__global__ void __launch_bounds__(256) k_geom(long long ntet, GeomParams gp, const double* __restrict__ x, const double* __restrict__ y,
                                              const double* __restrict__ z, const int32_t* __restrict__ v0, const int32_t* __restrict__ v1,
                                              const int32_t* __restrict__ v2, const int32_t* __restrict__ v3, double* __restrict__ gbuf,
                                              const unsigned* __restrict__ old2new /* NULL or the Morton id of every element */) {
    // records are built in shared memory and leave the CTA as 16-byte pieces, `parts` consecutive lanes per record: full
    // sectors per store instruction even though the records scatter (Morton order).
    // 64-byte records (SWZ): two records per 128-byte row, the 8-byte slot of component k XOR-ed with the row number, so that
    // the 32 lanes writing component k of their records hit 16 different bank pairs (natural layout: stride 64 bytes = 8-way
    // conflicts on every store, the LSU pipe was 97 % busy; profiles/r02b_rings.md)
    extern __shared__ double srec[];                       // [256][ngpad]
    __shared__ long long sdst[256];
    const long long e0 = blockIdx.x * (long long)blockDim.x;
    const long long e = e0 + threadIdx.x;
    const bool valid = e < ntet;
    const bool swz = gp.ngpad == 8;
    const int swf = swz ? ((threadIdx.x >> 1) & 7) : 0;
    double* g = swz ? srec + (size_t)(threadIdx.x >> 1) * 16 + ((threadIdx.x & 1) << 3) : srec + (size_t)threadIdx.x * gp.ngpad;
#define GG(idx) g[(idx) ^ swf]
    if (valid) {
    sdst[threadIdx.x] = old2new ? (long long)old2new[e] : e;
    const int nn[4] = {__ldg(v0 + e), __ldg(v1 + e), __ldg(v2 + e), __ldg(v3 + e)};
    double P[4][3], PSI[9];
#pragma unroll
    for (int k = 0; k < 4; ++k) { P[k][0] = __ldg(x + nn[k]); P[k][1] = __ldg(y + nn[k]); P[k][2] = __ldg(z + nn[k]); }
    const double vol = fabs(jac_inv(P, PSI)) * (1.0 / 6.0);
    if (gp.ngpad > gp.ngtot) GG(gp.ngtot) = 0.0;
    if (gp.zero_all) for (int k = 0; k < gp.ngpad; ++k) GG(k) = 0.0;
    for (int f = 0; f < gp.nforms; ++f) {
        const TFormDev& F = gp.f[f];
        const double* D = F.D;
        if (F.layout == AFB_COEF_PER_TET) D += (size_t)F.dstride * e;
        const double s = vol * F.alpha;
        const int ob = F.goff;
        if (F.kind == 0) {
            if (F.full) {
                double K[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) K[t] = coef_value(D, F.kidx[t]);   // K(k,l) at k + 3l
                // R[k][b] = sum_l K(k,l) PSI[b+3l];  M[a][b] = sum_k PSI[a+3k] R[k][b]
                double R[3][3];
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int b = 0; b < 3; ++b) R[k][b] = K[k] * PSI[b] + K[k + 3] * PSI[b + 3] + K[k + 6] * PSI[b + 6];
                double M[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) M[a][b] = s * (PSI[a] * R[0][b] + PSI[a + 3] * R[1][b] + PSI[a + 6] * R[2][b]);
                if (F.ng == 6) { GG(ob + (0)) = M[0][0]; GG(ob + (1)) = M[1][1]; GG(ob + (2)) = M[2][2]; GG(ob + (3)) = M[0][1]; GG(ob + (4)) = M[0][2]; GG(ob + (5)) = M[1][2]; }
                else {
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) GG(ob + (3 * a + b)) = M[a][b];
                }
            } else {
                const double c = s * coef_value(D, F.kidx[0]);
                // M = c * PSI PSI^T (rows a of the inverse Jacobian dotted)
                GG(ob + (0)) = c * (PSI[0] * PSI[0] + PSI[3] * PSI[3] + PSI[6] * PSI[6]);
                GG(ob + (1)) = c * (PSI[1] * PSI[1] + PSI[4] * PSI[4] + PSI[7] * PSI[7]);
                GG(ob + (2)) = c * (PSI[2] * PSI[2] + PSI[5] * PSI[5] + PSI[8] * PSI[8]);
                GG(ob + (3)) = c * (PSI[0] * PSI[1] + PSI[3] * PSI[4] + PSI[6] * PSI[7]);
                GG(ob + (4)) = c * (PSI[0] * PSI[2] + PSI[3] * PSI[5] + PSI[6] * PSI[8]);
                GG(ob + (5)) = c * (PSI[1] * PSI[2] + PSI[4] * PSI[5] + PSI[7] * PSI[8]);
            }
            if (F.bary) {
                // M = G restricted to the vertices 1..3 (reference gradients d/dxi_a = grad l_a, a = 1..3); the barycentric gradients
                // sum to zero, so G_0b = -(M_1b + M_2b + M_3b).  Record order of the ring kernel: (G01,G23), (G02,G13), (G03,G12)
                const double m00 = GG(ob + (0)), m11 = GG(ob + (1)), m22 = GG(ob + (2)), m01 = GG(ob + (3)), m02 = GG(ob + (4)), m12 = GG(ob + (5));
                GG(ob + (0)) = -(m00 + m01 + m02); GG(ob + (1)) = m12;
                GG(ob + (2)) = -(m01 + m11 + m12); GG(ob + (3)) = m02;
                GG(ob + (4)) = -(m02 + m12 + m22); GG(ob + (5)) = m01;
            }
        } else if (F.kind == 1) {
            GG(ob + (0)) = s * coef_value(D, F.kidx[0]);
        } else {
            // kind 2: K(0,l) = k_l -> GG(ob + (b)) = s sum_l k_l PSI[b+3l];  kind 3: K(k,0) = k_k -> GG(ob + (a)) = s sum_k PSI[a+3k] k_k
            const double k0 = coef_value(D, F.kidx[0]), k1 = coef_value(D, F.kidx[1]), k2 = coef_value(D, F.kidx[2]);
#pragma unroll
            for (int a = 0; a < 3; ++a) GG(ob + (a)) = s * (PSI[a] * k0 + PSI[a + 3] * k1 + PSI[a + 6] * k2);
        }
    }
    }  // valid
#undef GG
    __syncthreads();
    const int parts = gp.ngpad >> 1;
    const int nrec = (int)min((long long)blockDim.x, ntet - e0);
    const double2* src = reinterpret_cast<const double2*>(srec);
    if (swz) {
        for (int q = threadIdx.x; q < nrec * 4; q += blockDim.x) {
            const int rec = q >> 2, part = q & 3, f = (rec >> 1) & 7;
            double2 d = src[(size_t)(rec >> 1) * 8 + ((rec & 1) << 2) + (part ^ (f >> 1))];
            if (f & 1) { const double t = d.x; d.x = d.y; d.y = t; }
            reinterpret_cast<double2*>(gbuf + sdst[rec] * 8)[part] = d;
        }
    } else
    for (int q = threadIdx.x; q < nrec * parts; q += blockDim.x) {
        const int rec = q / parts, part = q - rec * parts;
        reinterpret_cast<double2*>(gbuf + sdst[rec] * gp.ngpad)[part] = src[q];
    }
}

struct GatherT {
    long long nrows, ntet;
    int nrow_loc, ncol_loc, max_len;
    int G, gpw, passes;        // lanes per row, rows per warp, column passes per visit
    int nga, ngf, ngpad;       // matrix / rhs components, stride of gbuf
    unsigned long long divM;   // ceil(2^40 / nrow_loc)
    const long long* rowptr;
    const long long* radj_ptr;
    const unsigned* radj;
    const void* pos;           // uint8 | uint16, adjacency order
    int pos_bytes;
    const double* gbuf;
    const double* TA;          // [nga][nrow_loc][ncol_loc]
    const double* TF;          // [ngf][nrow_loc]
    double* val;
    double* rhs;
    int accumulate;
    double drop_val;
    int* status;
};

template <int NGMAX>
__global__ void __launch_bounds__(256) k_gather_tensor(GatherT p) {
    extern __shared__ double sm[];
    const int tabA = p.nga * p.nrow_loc * p.ncol_loc, tabF = p.ngf * p.nrow_loc;
    double* sTA = sm;
    double* sTF = sTA + tabA;
    double* sacc = sTF + tabF;
    for (int t = threadIdx.x; t < tabA; t += blockDim.x) sTA[t] = p.TA[t];
    for (int t = threadIdx.x; t < tabF; t += blockDim.x) sTF[t] = p.TF[t];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int g = lane / p.G;           // group inside the warp
    if (g >= p.gpw) return;             // idle tail lanes
    const int gl = lane - g * p.G;
    const unsigned gmask = (p.G == 32) ? 0xffffffffu : (((1u << p.G) - 1u) << (g * p.G));
    double* acc = sacc + (size_t)(warp * p.gpw + g) * p.max_len;
    const int rpb = wpb * p.gpw;
    const bool doA = p.val != nullptr, doF = p.rhs != nullptr;
    bool bad = false;
    for (long long r = (long long)blockIdx.x * rpb + warp * p.gpw + g; r < p.nrows; r += (long long)gridDim.x * rpb) {
        const long long p0 = p.rowptr[r];
        const int len = (int)(p.rowptr[r + 1] - p0);
        const long long a0 = p.radj_ptr[r], a1 = p.radj_ptr[r + 1];
        if (doA) {
            for (int s = gl; s < len; s += p.G) acc[s] = 0.0;
            __syncwarp(gmask);
        }
        double fsum = 0.0;
        for (long long a = a0; a < a1; ++a) {
            const unsigned t = __ldg(p.radj + a);
            const unsigned e = (unsigned)(((unsigned long long)t * p.divM) >> 40);
            const int i = (int)(t - e * (unsigned)p.nrow_loc);
            const double2* ge = reinterpret_cast<const double2*>(p.gbuf + (size_t)e * p.ngpad);
            double gv[NGMAX];
#pragma unroll
            for (int c = 0; c < NGMAX / 2; ++c) {
                if (2 * c < p.nga + p.ngf) { const double2 d = __ldg(ge + c); gv[2 * c] = d.x; gv[2 * c + 1] = d.y; }
                else { gv[2 * c] = 0.0; gv[2 * c + 1] = 0.0; }
            }
            if (doF && gl == 0) {
                double f = 0.0;
#pragma unroll
                for (int c = 0; c < NGMAX; ++c)
                    if (c >= p.nga && c < p.nga + p.ngf) f += sTF[(c - p.nga) * p.nrow_loc + i] * gv[c];
                bad |= !isfinite(f);
                fsum += f;
            }
            if (doA) {
                const size_t base = (size_t)a * p.ncol_loc;
                for (int ps = 0; ps < p.passes; ++ps) {
                    const int j = gl + ps * p.G;
                    if (j < p.ncol_loc) {
                        const double* T = sTA + i * p.ncol_loc + j;
                        const int stride = p.nrow_loc * p.ncol_loc;
                        double v = 0.0;
#pragma unroll
                        for (int c = 0; c < NGMAX; ++c)
                            if (c < p.nga) v += T[c * stride] * gv[c];
                        const int sl = p.pos_bytes == 1 ? (int)static_cast<const unsigned char*>(p.pos)[base + j]
                                                        : (int)static_cast<const unsigned short*>(p.pos)[base + j];
                        bad |= !isfinite(v);
                        if (fabs(v) > p.drop_val) acc[sl] += v;
                    }
                }
                __syncwarp(gmask);
            }
        }
        if (doA) {
            if (p.accumulate) for (int s = gl; s < len; s += p.G) p.val[p0 + s] += acc[s];
            else for (int s = gl; s < len; s += p.G) p.val[p0 + s] = acc[s];
            __syncwarp(gmask);
        }
        if (doF && gl == 0) {
            if (p.accumulate) p.rhs[r] += fsum; else p.rhs[r] = fsum;
        }
    }
    if (bad) *p.status = 1;
}

// Specialisation for square element matrices with compile-time sizes (NLOC = 4, 10, 20), NGA matrix components and
// NGF rhs components: no predicates in the inner loop, shared memory addressed through 32-bit shared-window
// addresses with immediate offsets, the slot bytes of a visit read before the FMA chain, the adjacency entry of the
// next visit prefetched, one warp-uniform trip count (max degree of the warp's rows) so that the per-visit barrier is a
// plain full-mask __syncwarp, and NaN/Inf detection folded into one FMA per entry (v*0 accumulates NaN).
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

template <int NLOC, int G, int NGA, int NGF, typename PosT>
__global__ void __launch_bounds__(256) k_gather_tensor_sq(GatherT p) {
    // G = lanes per row
    constexpr int PASSES = (NLOC + G - 1) / G;    // columns per lane
    constexpr int GPW = 32 / G;                   // rows per warp
    constexpr int TAB = NLOC * NLOC;
    constexpr int NG = NGA + NGF;
    constexpr int NGP = (NG + 1) & ~1;
    static_assert(PASSES * G == NLOC, "columns must tile the lane group");
    extern __shared__ double sm[];
    double* sTA = sm;                   // [NGA][NLOC][NLOC]
    double* sTF = sTA + NGA * TAB;      // [NGF][NLOC]
    double* sacc = sTF + NGF * NLOC;
    for (int t = threadIdx.x; t < NGA * TAB; t += 256) sTA[t] = p.TA[t];
    for (int t = threadIdx.x; t < NGF * NLOC; t += 256) sTF[t] = p.TF[t];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = min(lane / G, GPW - 1);
    const bool lane_on = lane < GPW * G;          // the tail lanes of a warp idle but keep the warp converged
    const int gl = lane_on ? lane - g * G : 0;
    constexpr int RPB = 8 * GPW;
    const bool doA = p.val != nullptr, doF = p.rhs != nullptr;
    const PosT* __restrict__ pos = static_cast<const PosT*>(p.pos);
    const unsigned* __restrict__ radj = p.radj;
    const double* __restrict__ gbuf = p.gbuf;
    const double drop = p.drop_val;
    const unsigned sTA_a = (unsigned)__cvta_generic_to_shared(sTA) + gl * 8;
    const unsigned sTF_a = (unsigned)__cvta_generic_to_shared(sTF);
    const unsigned acc_a = (unsigned)__cvta_generic_to_shared(sacc) + (unsigned)(warp * GPW + g) * p.max_len * 8;
    double chk = 0.0;
    const long long nrows_pad = ((p.nrows + RPB - 1) / RPB) * RPB;
    for (long long r0 = (long long)blockIdx.x * RPB + warp * GPW; r0 < nrows_pad; r0 += (long long)gridDim.x * RPB) {
        const long long r = r0 + g;
        const bool row_on = lane_on && r < p.nrows;
        long long p0 = 0, a0 = 0;
        int len = 0, deg = 0;
        if (row_on) {
            p0 = p.rowptr[r];
            len = (int)(p.rowptr[r + 1] - p0);
            a0 = p.radj_ptr[r];
            deg = (int)(p.radj_ptr[r + 1] - a0);
        }
        const int degmax = __reduce_max_sync(0xffffffffu, deg);
        if (doA) {
            for (int s = gl; s < len; s += G) sts64(acc_a + s * 8, 0.0);
            __syncwarp();
        }
        double fsum = 0.0;
        const unsigned* ra = radj + a0;
        const PosT* pa = pos + a0 * NLOC + gl;
        unsigned tn = deg > 0 ? __ldg(ra) : 0u;
        for (int k = 0; k < degmax; ++k) {
            const bool on = k < deg;
            const unsigned t = tn;
            if (k + 1 < deg) tn = __ldg(ra + k + 1);
            const unsigned e = t / (unsigned)NLOC;
            const int i = (int)(t - e * (unsigned)NLOC);
            int sl[PASSES];
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) sl[ps] = on ? (int)pa[k * NLOC + ps * G] : 0;
            double gv[NGP];
            if (on) {
                const double2* ge = reinterpret_cast<const double2*>(gbuf + (size_t)e * NGP);
#pragma unroll
                for (int c = 0; c < NGP / 2; ++c) { const double2 d = __ldg(ge + c); gv[2 * c] = d.x; gv[2 * c + 1] = d.y; }
            } else {
#pragma unroll
                for (int c = 0; c < NGP; ++c) gv[c] = 0.0;
            }
            if (NGF > 0 && doF && gl == 0) {
                double f = 0.0;
#pragma unroll
                for (int c = 0; c < NGF; ++c) f = fma(lds64(sTF_a + (c * NLOC + i) * 8), gv[NGA + c], f);
                chk = fma(f, 0.0, chk);
                fsum += f;
            }
            if (NGA > 0 && doA) {
                const unsigned Ta = sTA_a + i * (NLOC * 8);
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NGA; ++c) v = fma(lds64(Ta + (c * TAB + ps * G) * 8), gv[c], v);
                    chk = fma(v, 0.0, chk);
                    if (on && fabs(v) > drop) {
                        const unsigned sa = acc_a + sl[ps] * 8;
                        sts64(sa, lds64(sa) + v);
                    }
                }
                __syncwarp();
            }
        }
        if (doA) {
            if (p.accumulate) for (int s = gl; s < len; s += G) p.val[p0 + s] += lds64(acc_a + s * 8);
            else for (int s = gl; s < len; s += G) p.val[p0 + s] = lds64(acc_a + s * 8);
            __syncwarp();
        }
        if (doF && row_on && gl == 0) {
            if (p.accumulate) p.rhs[r] += fsum; else p.rhs[r] = fsum;
        }
    }
    if (chk != chk) *p.status = 1;   // NaN <=> some local value was NaN or +-Inf
}

template <int NLOC, int G, int NGA, int NGF, typename PosT>
cudaError_t launch_sq(const GatherT& p, cudaStream_t st) {
    constexpr int RPB = 8 * (32 / G);
    const size_t smem = ((size_t)NGA * NLOC * NLOC + (size_t)NGF * NLOC + (size_t)RPB * p.max_len) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_gather_tensor_sq<NLOC, G, NGA, NGF, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((p.nrows + RPB - 1) / RPB, 148LL * 32));
    k_gather_tensor_sq<NLOC, G, NGA, NGF, PosT><<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NLOC, int G, typename PosT>
cudaError_t launch_sq_ng(const GatherT& p, int nga, int ngf, cudaStream_t st, bool* done) {
    *done = true;
#define SQ(A, F) if (nga == A && ngf == F) return launch_sq<NLOC, G, A, F, PosT>(p, st);
    SQ(6, 1) SQ(6, 0) SQ(7, 1) SQ(7, 0) SQ(1, 1) SQ(1, 0) SQ(0, 1) SQ(9, 1) SQ(9, 0) SQ(10, 1) SQ(10, 0) SQ(3, 0) SQ(3, 1)
#undef SQ
    *done = false;
    return cudaSuccess;
}

// S tables on the host
struct HostForm {
    int kind, ng;
    std::vector<double> T;  // [ng][nfb][nfa]
};

void build_form_table(const SForm& f, std::vector<double>& T) {
    const double *pq, *wq;
    const int q = tet_rule(f.quad_order, &pq, &wq);
    const int nfa = f.nfa, nfb = f.nfb, kind = f.kind, ng = f.ng;
    std::vector<double> phiA((size_t)q * nfa), phiB((size_t)q * nfb), GA((size_t)q * nfa * 3), GB((size_t)q * nfb * 3);
    basis_values(f.femA, q, pq, phiA.data());
    basis_values(f.femB, q, pq, phiB.data());
    basis_ref_grads(f.femA, q, pq, GA.data());
    basis_ref_grads(f.femB, q, pq, GB.data());
    T.assign((size_t)ng * nfb * nfa, 0.0);
    auto S = [&](int a, int b, int i, int j) {  // sum_n w_n GB[i][a] GA[j][b]
        double s = 0;
        for (int n = 0; n < q; ++n) s += wq[n] * GB[((size_t)n * nfb + i) * 3 + a] * GA[((size_t)n * nfa + j) * 3 + b];
        return s;
    };
    for (int i = 0; i < nfb; ++i)
        for (int j = 0; j < nfa; ++j) {
            auto at = [&](int c) -> double& { return T[((size_t)c * nfb + i) * nfa + j]; };
            if (kind == 0) {
                if (ng == 9) { for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) at(3 * a + b) = S(a, b, i, j); }
                else {
                    at(0) = S(0, 0, i, j); at(1) = S(1, 1, i, j); at(2) = S(2, 2, i, j);
                    at(3) = S(0, 1, i, j) + S(1, 0, i, j); at(4) = S(0, 2, i, j) + S(2, 0, i, j); at(5) = S(1, 2, i, j) + S(2, 1, i, j);
                }
            } else if (kind == 1) {
                double s = 0;
                for (int n = 0; n < q; ++n) s += wq[n] * phiB[(size_t)n * nfb + i] * phiA[(size_t)n * nfa + j];
                at(0) = s;
            } else if (kind == 2) {
                for (int b = 0; b < 3; ++b) {
                    double s = 0;
                    for (int n = 0; n < q; ++n) s += wq[n] * phiB[(size_t)n * nfb + i] * GA[((size_t)n * nfa + j) * 3 + b];
                    at(b) = s;
                }
            } else {
                for (int a = 0; a < 3; ++a) {
                    double s = 0;
                    for (int n = 0; n < q; ++n) s += wq[n] * GB[((size_t)n * nfb + i) * 3 + a] * phiA[(size_t)n * nfa + j];
                    at(a) = s;
                }
            }
        }
}

template <int NGMAX>
cudaError_t launch_gt(const GatherT& p, unsigned grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_gather_tensor<NGMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_gather_tensor<NGMAX><<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

namespace afb {

// The fused path for a group of scalar block forms that share one gather plan.  `ctx` owns the mesh and the work buffers,
// `plan` the dof map / pattern / gather plan the forms are assembled with (== ctx for single-field problems, a pair
// sub-context of afb_blocks.cu for the blocks of vector and mixed spaces; then p0_override gives the first CSR entry of every
// plan row inside the caller's matrix).  Returns 1 (lane-group gather) or 2 (cluster gather k_rows_cl), 0 if the group
// cannot be handled (nothing was launched), < 0 on error.
int fused_group(afb_ctx* ctx, afb_ctx* plan, const std::vector<SForm>& mat, const std::vector<SForm>& rhsf, double* dval, double* drhs,
                const long long* p0_override, int accumulate, double drop_val, int* status_flag, bool record_events, int phase,
                const int* tix, const unsigned short* rtab, const int* rdst) {
    const int nfA = (int)mat.size(), nfF = (int)rhsf.size(), nforms = nfA + nfF;
    if (nforms == 0 || nforms > MAX_TFORMS) return 0;
    const int nrl = plan->nrow_loc, ncl = plan->ncol_loc;
    int nga = 0, ngf = 0;
    for (const SForm& f : mat) nga += f.ng;
    for (const SForm& f : rhsf) ngf += f.ng;
    const int ngtot = nga + ngf;
    if (ngtot > 32) return 0;
    const size_t tabA = (size_t)nga * nrl * ncl, tabF = (size_t)ngf * nrl;
    if ((tabA + tabF) * 8 > 96 * 1024) return 0;
    const int ngpad = (ngtot + 1) & ~1;
    // the cluster gather reads the coefficients in the Morton order of its plan, the lane-group gather in mesh order
    const bool use_rows = rows_supports(plan, nga, ngf) && (nga == 0 || dval) && (ngf == 0 || drhs);
    if (!use_rows && (plan != ctx || p0_override)) return 0;

    // ---- host tables T[alpha][i][j] over the whole local matrix of the plan, rhs table TF[beta][i]
    std::vector<double> TA(tabA, 0.0), TF(tabF, 0.0);
    GeomParams gp;
    std::memset(&gp, 0, sizeof(gp));
    gp.nforms = nforms; gp.ngpad = ngpad; gp.ngtot = ngtot;
    int offA = 0, offF = nga;
    for (int k = 0; k < nforms; ++k) {
        const bool ismat = k < nfA;
        const SForm& f = ismat ? mat[k] : rhsf[k - nfA];
        std::vector<double> T;
        build_form_table(f, T);
        const int goff = ismat ? offA : offF;
        for (int c = 0; c < f.ng; ++c)
            for (int i = 0; i < f.nfb; ++i)
                for (int j = 0; j < f.nfa; ++j) {
                    const double v = T[((size_t)c * f.nfb + i) * f.nfa + j];
                    if (ismat) TA[((size_t)(goff + c) * nrl + f.row_off + i) * ncl + f.col_off + j] = v;
                    else TF[(size_t)(goff - nga + c) * nrl + f.row_off + i] = v;
                }
        TFormDev& d = gp.f[k];
        d.kind = f.kind; d.full = f.full; d.layout = f.layout; d.dstride = f.dstride;
        for (int t = 0; t < 9; ++t) d.kidx[t] = f.kidx[t];
        d.goff = goff; d.ng = f.ng; d.alpha = f.alpha; d.D = f.D;
        (ismat ? offA : offF) += f.ng;
    }
    cudaStream_t st = ctx->stream;
    // ---- ring-traversal kernel (afb_rings.cu): square P2 problems with one symmetric stiffness form (+ mass, + load)
    {
        int is = -1, im = -1, nother = 0;
        for (int k = 0; k < nfA; ++k) {
            const SForm& f = mat[k];
            const bool p2 = f.nfa == 10 && f.nfb == 10 && f.row_off == 0 && f.col_off == 0;
            if (p2 && f.kind == 0 && f.ng == 6 && is < 0) is = k;
            else if (p2 && f.kind == 1 && im < 0) im = k;
            else ++nother;
        }
        const bool load_ok = nfF == 0 || (nfF == 1 && rhsf[0].kind == 1 && rhsf[0].ng == 1 && rhsf[0].nfb == 10 && rhsf[0].row_off == 0);
        const bool eligible = plan == ctx && !p0_override && !tix && !rtab && !rdst && dval && is >= 0 && nother == 0 && load_ok && (nfF == 0 || drhs) &&
                              !getenv("AFB_DISABLE_RING_KERNEL");
        if (eligible) { const int rce = ensure_ring_plan(ctx); if (rce) return rce; }
        if (eligible && rings_supports(ctx, 1, im >= 0 ? 1 : 0, nfF)) {
            std::vector<double> TM, TG(600), Tm, Tf;
            build_form_table(mat[is], TM);
            ring_table_from_M(TM.data(), 10, TG.data());
            if (im >= 0) build_form_table(mat[im], Tm);
            if (nfF) build_form_table(rhsf[0], Tf);
            // the kernel relies on two symmetries of the tables: invariance under vertex relabelling (one frame for every ring tet)
            // and T[q][i][j] == T[q][j][i] (entries (a,ab) / (ab,a) etc. are evaluated once)
            double asym = 0.0, tmax = 0.0;
            for (int q = 0; q < 6; ++q)
                for (int i = 0; i < 10; ++i)
                    for (int j = 0; j < 10; ++j) {
                        asym = std::max(asym, std::fabs(TG[(q * 10 + i) * 10 + j] - TG[(q * 10 + j) * 10 + i]));
                        tmax = std::max(tmax, std::fabs(TG[(q * 10 + i) * 10 + j]));
                    }
            if (im >= 0)
                for (int i = 0; i < 10; ++i)
                    for (int j = 0; j < 10; ++j) asym = std::max(asym, std::fabs(Tm[i * 10 + j] - Tm[j * 10 + i]) * (tmax > 0 ? 1.0 : 0.0));
            if (ring_table_symmetry_defect(TG.data()) < 1e-12 && asym <= 1e-13 * tmax) {
                GeomParams gr;
                std::memset(&gr, 0, sizeof(gr));
                gr.ngpad = 8; gr.ngtot = 8; gr.zero_all = 1;
                auto add = [&](const SForm& f, int goff, int bary) {
                    TFormDev& d = gr.f[gr.nforms++];
                    d.kind = f.kind; d.full = f.full; d.layout = f.layout; d.dstride = f.dstride;
                    for (int t = 0; t < 9; ++t) d.kidx[t] = f.kidx[t];
                    d.goff = goff; d.ng = f.ng; d.bary = bary; d.alpha = f.alpha; d.D = f.D;
                };
                add(mat[is], 0, 1);
                if (im >= 0) add(mat[im], 6, 0);
                if (nfF) add(rhsf[0], 7, 0);
                AFB_CUDA(ctx, ctx->stageF.reserve((size_t)ctx->ntet * 8 * sizeof(double)));
                double* gb = ctx->stageF.as<double>();
                if (record_events) cudaEventRecord(ctx->ev[1], st);
                if (phase != 2) {
                    k_geom<<<(unsigned)((ctx->ntet + 255) / 256), 256, (size_t)256 * 8 * sizeof(double), st>>>(
                        ctx->ntet, gr, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(), ctx->v[0].as<int32_t>(), ctx->v[1].as<int32_t>(),
                        ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>(), gb, ctx->rp_old2new.as<unsigned>());
                    ctx->launches++;
                    AFB_CUDA(ctx, cudaGetLastError());
                }
                if (record_events) cudaEventRecord(ctx->ev[2], st);
                const int rc = launch_rings(ctx, TG.data(), im >= 0 ? Tm.data() : nullptr, nfF ? Tf.data() : nullptr, gb, dval, drhs, accumulate, drop_val,
                                            status_flag, phase);
                if (rc < 0) return rc;
                if (record_events) cudaEventRecord(ctx->ev[3], st);
                return 3;
            }
        }
    }
    AFB_CUDA(ctx, ctx->stageF.reserve((size_t)ctx->ntet * ngpad * sizeof(double)));  // g_e buffer
    double* gbuf = ctx->stageF.as<double>();

    if (record_events) cudaEventRecord(ctx->ev[1], st);
    const unsigned gridg = (unsigned)((ctx->ntet + 255) / 256);
    if (phase != 2)   // phase 2 of a phased assembly reuses the coefficients phase 1 left in the context
    k_geom<<<gridg, 256, (size_t)256 * ngpad * sizeof(double), st>>>(ctx->ntet, gp, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(), ctx->v[0].as<int32_t>(),
                                 ctx->v[1].as<int32_t>(), ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>(), gbuf,
                                 use_rows ? plan->rp_old2new.as<unsigned>() : nullptr);
    ctx->launches++;
    AFB_CUDA(ctx, cudaGetLastError());
    if (record_events) cudaEventRecord(ctx->ev[2], st);

    if (use_rows) {
        // cluster-tiled thread-per-row gather (afb_rows.cu)
        plan->stream = ctx->stream;
        const int rc = launch_rows(plan, nga, ngf, TA.data(), TF.data(), gbuf, dval, drhs, accumulate, drop_val, status_flag, p0_override, phase, tix, rtab, rdst);
        if (plan != ctx) { ctx->launches += plan->launches; plan->launches = 0; if (rc < 0) set_error(ctx, plan->err); }
        if (rc < 0) return rc;
        if (rc != 1) { set_error(ctx, "internal: cluster gather refused a case it advertised"); return -4; }
        if (record_events) cudaEventRecord(ctx->ev[3], st);
        return 2;
    }

    // ---- lane-group gather (single-field plans only)
    // tables: small, uploaded through a per-context device buffer; stream-ordered so reuse across calls is safe
    AFB_CUDA(ctx, ctx->tables.reserve((tabA + tabF + 2) * sizeof(double)));
    if (tabA) AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tables.p, TA.data(), tabA * sizeof(double), cudaMemcpyHostToDevice, st));
    if (tabF) AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tables.as<double>() + tabA, TF.data(), tabF * sizeof(double), cudaMemcpyHostToDevice, st));
    int bestG = 1; double bestU = -1;
    for (int G = 1; G <= 32; ++G) {
        const int passes = (ncl + G - 1) / G;
        if (passes > 4) continue;
        const double util = (double)ncl / (passes * G) * ((32 / G) * G / 32.0);
        if (util > bestU + 1e-9) { bestU = util; bestG = G; }
    }
    GatherT p;
    p.nrows = ctx->row_end - ctx->row_begin; p.ntet = ctx->ntet;
    p.nrow_loc = nrl; p.ncol_loc = ncl; p.max_len = std::max(1, ctx->max_row_len);
    p.G = bestG; p.gpw = 32 / bestG; p.passes = (ncl + bestG - 1) / bestG;
    p.nga = nga; p.ngf = ngf; p.ngpad = ngpad;
    p.divM = ((1ULL << 40) + nrl - 1) / nrl;
    p.rowptr = ctx->rowptr.as<long long>(); p.radj_ptr = ctx->radj_ptr.as<long long>(); p.radj = ctx->radj.as<unsigned>();
    p.pos = ctx->pos.p; p.pos_bytes = ctx->pos_bytes; p.gbuf = gbuf;
    p.TA = ctx->tables.as<double>(); p.TF = ctx->tables.as<double>() + tabA;
    p.val = dval; p.rhs = drhs; p.accumulate = accumulate; p.drop_val = drop_val; p.status = status_flag;
    const int rpb = 8 * p.gpw;
    const size_t smem = (tabA + tabF + (size_t)rpb * p.max_len) * sizeof(double);
    if (smem > 200 * 1024) { set_error(ctx, "afb_assemble: matrix rows too long for the shared-memory row image"); return -3; }
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((p.nrows + rpb - 1) / rpb, 148LL * 32));
    cudaError_t e = cudaSuccess;
    bool done = false;
    if (nrl == ncl && (nrl == 4 || nrl == 10 || nrl == 20) && !getenv("AFB_DISABLE_SQ_KERNEL")) {
        // lanes per row: fewer lanes -> more rows (independent load chains) per warp and less per-visit overhead
        int G = nrl == 4 ? 2 : 5;
        if (const char* gs = getenv("AFB_SQ_G")) G = atoi(gs);
        if (nrl == 4 && G != 2 && G != 4) G = 2;
        if (nrl != 4 && G != 5 && G != 10) G = 5;
        const size_t smem_sq = ((size_t)nga * nrl * nrl + (size_t)ngf * nrl + (size_t)8 * (32 / G) * p.max_len) * sizeof(double);
        if (smem_sq <= 200 * 1024) {
#define DISPATCH(NL, GG)                                                                                  \
    if (nrl == NL && G == GG) {                                                                           \
        if (ctx->pos_bytes == 1) e = launch_sq_ng<NL, GG, unsigned char>(p, nga, ngf, st, &done);         \
        else e = launch_sq_ng<NL, GG, unsigned short>(p, nga, ngf, st, &done);                            \
    }
            DISPATCH(4, 2) DISPATCH(4, 4) DISPATCH(10, 5) DISPATCH(10, 10) DISPATCH(20, 5) DISPATCH(20, 10)
#undef DISPATCH
        }
    }
    if (done) { /* specialised kernel launched */ }
    else if (ngtot <= 2) e = launch_gt<2>(p, grid, smem, st);
    else if (ngtot <= 6) e = launch_gt<6>(p, grid, smem, st);
    else if (ngtot <= 8) e = launch_gt<8>(p, grid, smem, st);
    else if (ngtot <= 12) e = launch_gt<12>(p, grid, smem, st);
    else if (ngtot <= 16) e = launch_gt<16>(p, grid, smem, st);
    else if (ngtot <= 24) e = launch_gt<24>(p, grid, smem, st);
    else e = launch_gt<32>(p, grid, smem, st);
    ctx->launches++;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_gather_tensor launch");
    if (record_events) cudaEventRecord(ctx->ev[3], st);
    return 1;
}

// Scalar form (vecA = vecB = 1, operators IDEN / GRAD, CONST or PER_TET coefficient) -> SForm.  false: not representable.
bool make_sform(const afb_form& f, const OpInfo& oa, const OpInfo& ob, const double* Ddev, SForm* out) {
    if (oa.vec != 1 || ob.vec != 1) return false;
    if (f.coef_layout == AFB_COEF_PER_POINT) return false;
    const bool gA = f.opA == AFB_GRAD, gB = f.opB == AFB_GRAD;
    if ((f.opA != AFB_IDEN && !gA) || (f.opB != AFB_IDEN && !gB)) return false;
    SForm s;
    std::memset(&s, 0, sizeof(s));
    s.kind = gA ? (gB ? 0 : 2) : (gB ? 3 : 1);
    const int tt = f.tensor_type;
    if (tt == AFB_TENSOR_SYMMETRIC && oa.dim != ob.dim) return false;
    const int dlen = form_dlen(f, oa, ob);
    s.layout = f.coef_layout; s.dstride = dlen; s.D = Ddev; s.alpha = f.alpha;
    for (int t = 0; t < 9; ++t) s.kidx[t] = -1;
    if (s.kind == 0) {
        s.full = tt >= AFB_TENSOR_SYMMETRIC;
        s.ng = tt == AFB_TENSOR_GENERAL ? 9 : 6;
        if (s.full) for (int t = 0; t < 9; ++t) s.kidx[t] = t;
        else s.kidx[0] = tt == AFB_TENSOR_SCALAR ? 0 : -2;
    } else if (s.kind == 1) {
        s.ng = 1;
        s.kidx[0] = tt >= AFB_TENSOR_SCALAR ? 0 : -2;
    } else {
        s.ng = 3;
        if (tt >= AFB_TENSOR_SYMMETRIC) { s.kidx[0] = 0; s.kidx[1] = 1; s.kidx[2] = 2; }
        else {
            // scalar / identity tensors are only legal here through the IDEN(P0) broadcast (diff_tensor.h:333-338)
            if (!(s.kind == 3 && oa.nfa == 1)) return false;
            s.kidx[0] = s.kidx[1] = s.kidx[2] = tt == AFB_TENSOR_SCALAR ? 0 : -2;
        }
    }
    s.femA = oa.fem; s.femB = ob.fem; s.nfa = oa.nf_base; s.nfb = ob.nf_base;
    s.quad_order = f.quad_order; s.row_off = f.row_off; s.col_off = f.col_off;
    *out = s;
    return true;
}

// Single-field entry: every form is a scalar form on the context's own plan.
int assemble_tensor_path(afb_ctx* ctx, int nfA, int nfF, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa,
                         const std::vector<OpInfo>& ob, const std::vector<const double*>& Dd, double* dval, double* drhs,
                         int accumulate, double drop_val, int* status_flag, int phase) {
    if (getenv("AFB_DISABLE_TENSOR_PATH")) return 0;
    if (ctx->has_signs) return 0;
    std::vector<SForm> mat, rhsf;
    for (int k = 0; k < nfA + nfF; ++k) {
        SForm s;
        if (!make_sform(fm[k], oa[k], ob[k], Dd[k], &s)) return 0;
        (k < nfA ? mat : rhsf).push_back(s);
    }
    return fused_group(ctx, ctx, mat, rhsf, dval, drhs, nullptr, accumulate, drop_val, status_flag, true, phase);
}

}  // namespace afb
