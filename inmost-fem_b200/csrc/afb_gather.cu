// Deterministic, atomic-free value scatter (K3) and the assembly driver.
//
// Replaces (reference): the scatter loops of AssemblerT::Assemble / AssembleMatrix / AssembleRHS
// (inmost_interface/assembler.inl:397-481, :780-858, :555-573):
//     rhs[r] += s_r F[i];  matrix[r][c] += s_r s_c A(i,j)  if |A(i,j)| > drop_val;  NaN/Inf -> status -1
// The reference serialises row updates with INMOST's LockService and a find-or-insert per entry, so the summation
// order depends on thread timing.  Here the scatter is turned inside out: a group of lanes owns one matrix row,
// walks the row's adjacency list (element,local row) in ascending element order (plan built in afb_pattern.cu),
// reads that row of the staged element matrix with coalesced loads and adds it into a shared-memory image of the
// CSR row at the precomputed slots; the finished row is written once, coalesced.  No atomics, bit-reproducible.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "afb_internal.h"

using namespace afb;

namespace {

template <int G, bool SIGNS, typename PosT>
__global__ void __launch_bounds__(256) k_gather(long long nrows, long long ntet, int nrow_loc, int ncol_loc, int max_len,
                                                const long long* __restrict__ rowptr, const long long* __restrict__ radj_ptr,
                                                const unsigned* __restrict__ radj, const PosT* __restrict__ pos,
                                                const int32_t* __restrict__ e2r, const int32_t* __restrict__ e2c,
                                                const double* __restrict__ stageA, const double* __restrict__ stageF,
                                                double* __restrict__ val, double* __restrict__ rhs, int accumulate, double drop_val,
                                                int* __restrict__ status, long long e_lo, long long e_hi, int len_lo) {
    extern __shared__ double sacc[];
    const int gl = threadIdx.x % G;                    // lane inside the group
    const int gib = threadIdx.x / G;                   // group inside the block
    const int gpb = blockDim.x / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    double* acc = sacc + (size_t)gib * max_len;
    bool bad = false;
    for (long long r = (long long)blockIdx.x * gpb + gib; r < nrows; r += (long long)gridDim.x * gpb) {
        const long long p0 = rowptr[r];
        const int len = (int)(rowptr[r + 1] - p0);
        if (len <= len_lo) continue;   // rows of at most len_lo entries were done by k_gather_cols (group-uniform)
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        double fsum = 0.0;
        if (stageA) {
            for (int s = gl; s < len; s += G) acc[s] = 0.0;
            __syncwarp(gmask);
        }
        for (long long a = a0; a < a1; ++a) {
            const unsigned t = __ldg(radj + a);
            double sr = 1.0;
            const long long e = t / nrow_loc;
            if (e < e_lo || e >= e_hi) continue;   // the staged element matrices cover the elements [e_lo, e_hi) (group-uniform branch)
            const long long tl = (long long)t - e_lo * nrow_loc;  // index inside the staged chunk
            if (SIGNS) {
                const int i = (int)(t - e * nrow_loc);
                sr = e2r[(long long)i * ntet + e] < 0 ? -1.0 : 1.0;
            }
            if (stageF && gl == 0) {
                const double fv = __ldg(stageF + tl);
                bad |= !isfinite(fv);
                fsum += sr * fv;
            }
            if (stageA) {
                const long long base = tl * ncol_loc;
                for (int j = gl; j < ncol_loc; j += G) {
                    double v = __ldg(stageA + base + j);
                    const int p = pos[a * ncol_loc + j];
                    bad |= !isfinite(v);
                    if (SIGNS) v *= sr * (e2c[(long long)j * ntet + e] < 0 ? -1.0 : 1.0);
                    if (fabs(v) > drop_val) acc[p] += v;
                }
                __syncwarp(gmask);
            }
        }
        if (stageA) {
            if (accumulate) for (int s = gl; s < len; s += G) val[p0 + s] += acc[s];
            else for (int s = gl; s < len; s += G) val[p0 + s] = acc[s];
            __syncwarp(gmask);
        }
        if (stageF && gl == 0) {
            if (accumulate) rhs[r] += fsum; else rhs[r] = fsum;
        }
    }
    if (bad) *status = 1;  // benign race: every writer stores the same value
}

// k_gather_flat: the same row gather with the (adjacency entry, local column) pairs of a row flattened over the 32 lanes of a warp.
// k_gather gives a row to a group of 2^k lanes and walks the adjacency list entry by entry: with 20 local columns (P3) 12 of 32
// lanes idle and every entry costs one dependent round trip (adjacency word -> staged row + slots -> shared-memory add).  Here a
// lane fetches U items k = lane, lane + 32, ... ahead (values, slots), so 4 x 32 loads are in flight per warp and all lanes work;
// the adds then run entry by entry in ascending element order (one __syncwarp per entry), i.e. in exactly the order of k_gather:
// the results are bit-identical.
template <bool SIGNS, typename PosT>
__global__ void __launch_bounds__(256) k_gather_flat(long long nrows, long long ntet, int nrow_loc, int ncol_loc, int max_len,
                                                     const long long* __restrict__ rowptr, const long long* __restrict__ radj_ptr,
                                                     const unsigned* __restrict__ radj, const PosT* __restrict__ pos,
                                                     const int32_t* __restrict__ e2r, const int32_t* __restrict__ e2c,
                                                     const double* __restrict__ stageA, const double* __restrict__ stageF,
                                                     double* __restrict__ val, double* __restrict__ rhs, int accumulate, double drop_val,
                                                     int* __restrict__ status, long long e_lo, long long e_hi) {
    constexpr int U = 4;
    extern __shared__ double sacc[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* acc = sacc + (size_t)wib * max_len;
    bool bad = false;
    for (long long r = (long long)blockIdx.x * wpb + wib; r < nrows; r += (long long)gridDim.x * wpb) {
        const long long p0 = rowptr[r];
        const int len = (int)(rowptr[r + 1] - p0);
        const long long a0 = radj_ptr[r];
        const int nadj = (int)(radj_ptr[r + 1] - a0);
        if (stageF) {   // load contributions: fetched 32 at a time, added in adjacency order
            double fsum = 0.0;
            for (int b = 0; b < nadj; b += 32) {
                double fv = 0.0;
                if (b + lane < nadj) {
                    const unsigned t = __ldg(radj + a0 + b + lane);
                    const long long e = t / nrow_loc;
                    if (e >= e_lo && e < e_hi) {
                        fv = __ldg(stageF + ((long long)t - e_lo * nrow_loc));
                        bad |= !isfinite(fv);
                        if (SIGNS && e2r[(long long)(t - e * nrow_loc) * ntet + e] < 0) fv = -fv;
                    }
                }
                const int cnt = min(32, nadj - b);
                for (int i = 0; i < cnt; ++i) fsum += __shfl_sync(0xffffffffu, fv, i);
            }
            if (lane == 0) { if (accumulate) rhs[r] += fsum; else rhs[r] = fsum; }
        }
        if (!stageA) continue;
        for (int s = lane; s < len; s += 32) acc[s] = 0.0;
        __syncwarp();
        const int nitems = nadj * ncol_loc;
        int a = 0, j = lane;   // item of this lane: adjacency entry a, local column j
        while (j >= ncol_loc) { j -= ncol_loc; ++a; }
        for (int base = 0; base < nitems; base += 32 * U) {
            double v[U];
            int pp[U], ai[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = 0.0; pp[u] = -1; ai[u] = a;
                if (a < nadj) {
                    const unsigned t = __ldg(radj + a0 + a);
                    const long long e = t / nrow_loc;
                    if (e >= e_lo && e < e_hi) {
                        double x = __ldg(stageA + ((long long)t - e_lo * nrow_loc) * ncol_loc + j);
                        const int p = pos[(a0 + a) * ncol_loc + j];
                        bad |= !isfinite(x);
                        if (SIGNS) {
                            const int i = (int)(t - e * nrow_loc);
                            if ((e2r[(long long)i * ntet + e] < 0) != (e2c[(long long)j * ntet + e] < 0)) x = -x;
                        }
                        if (fabs(x) > drop_val) { v[u] = x; pp[u] = p; }
                    }
                }
                j += 32;
                while (j >= ncol_loc) { j -= ncol_loc; ++a; }
            }
            const int afirst = base / ncol_loc, alast = min(nadj - 1, (base + 32 * U - 1) / ncol_loc);
            for (int ac = afirst; ac <= alast; ++ac) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ai[u] == ac && pp[u] >= 0) acc[pp[u]] += v[u];
                __syncwarp();
            }
        }
        if (accumulate) for (int s = lane; s < len; s += 32) val[p0 + s] += acc[s];
        else for (int s = lane; s < len; s += 32) val[p0 + s] = acc[s];
        __syncwarp();
    }
    if (bad) *status = 1;  // benign race: every writer stores the same value
}

// k_gather_cols: lane-group gather for local matrices with NC = 10 or 20 columns (P2 / P3 through the staged path).  A group of
// G = NC / 5 lanes owns a row (8 or 16 rows per warp), every lane handles 5 columns of an adjacency entry and two entries are
// fetched per trip, so a warp keeps 2 x 5 x 32 value loads in flight where the warp-per-row gather has one dependent load per
// entry and 12 of 32 lanes idle.  Rows longer than `len_hi` entries are left to the warp-per-row kernel (len_lo there): the row
// images of this kernel are sized for the short rows (edge and face dofs: 95 % of the rows of a P3 problem).
// Same summation order as k_gather (ascending element per row): bit-identical results.
template <int G, int NC, bool SIGNS, typename PosT>
__global__ void __launch_bounds__(256) k_gather_cols(long long nrows, long long ntet, int img_len, int len_hi,
                                                     const long long* __restrict__ rowptr, const long long* __restrict__ radj_ptr,
                                                     const unsigned* __restrict__ radj, const PosT* __restrict__ pos,
                                                     const int32_t* __restrict__ e2r, const int32_t* __restrict__ e2c,
                                                     const double* __restrict__ stageA, const double* __restrict__ stageF,
                                                     double* __restrict__ val, double* __restrict__ rhs, int accumulate, double drop_val,
                                                     int* __restrict__ status, long long e_lo, long long e_hi) {
    constexpr int PER = NC / G;
    extern __shared__ double sacc[];
    const int gl = threadIdx.x % G, gib = threadIdx.x / G, gpb = blockDim.x / G;
    const unsigned gmask = ((1u << G) - 1u) << ((threadIdx.x & 31) / G * G);
    double* acc = sacc + (size_t)gib * img_len;
    bool bad = false;
    for (long long r = (long long)blockIdx.x * gpb + gib; r < nrows; r += (long long)gridDim.x * gpb) {
        const long long p0 = rowptr[r];
        const int len = (int)(rowptr[r + 1] - p0);
        if (len > len_hi) continue;   // group-uniform
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        double fsum = 0.0;
        if (stageA) {
            for (int s = gl; s < len; s += G) acc[s] = 0.0;
            __syncwarp(gmask);
        }
        for (long long a = a0; a < a1; a += 2) {
            const bool two = a + 1 < a1;
            const unsigned t0 = __ldg(radj + a), t1 = two ? __ldg(radj + a + 1) : t0;
            const long long ea = t0 / NC, eb = t1 / NC;
            const bool ina = ea >= e_lo && ea < e_hi, inb = two && eb >= e_lo && eb < e_hi;
            const long long tla = (long long)t0 - e_lo * NC, tlb = (long long)t1 - e_lo * NC;
            double va[PER], vb[PER];
            int pa[PER], pb[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int j = gl * PER + k;
                va[k] = (ina && stageA) ? __ldg(stageA + tla * NC + j) : 0.0;
                vb[k] = (inb && stageA) ? __ldg(stageA + tlb * NC + j) : 0.0;
                pa[k] = pos[a * NC + j];
                pb[k] = two ? pos[(a + 1) * NC + j] : 0;
            }
            double sra = 1.0, srb = 1.0;
            if (SIGNS) {
                sra = e2r[(long long)(t0 - ea * NC) * ntet + ea] < 0 ? -1.0 : 1.0;
                srb = e2r[(long long)(t1 - eb * NC) * ntet + eb] < 0 ? -1.0 : 1.0;
            }
            if (stageF && gl == 0) {
                if (ina) { const double fv = __ldg(stageF + tla); bad |= !isfinite(fv); fsum += sra * fv; }
                if (inb) { const double fv = __ldg(stageF + tlb); bad |= !isfinite(fv); fsum += srb * fv; }
            }
            if (stageA) {
                if (ina) {
#pragma unroll
                    for (int k = 0; k < PER; ++k) {
                        double v = va[k];
                        bad |= !isfinite(v);
                        if (SIGNS) v *= sra * (e2c[(long long)(gl * PER + k) * ntet + ea] < 0 ? -1.0 : 1.0);
                        if (fabs(v) > drop_val) acc[pa[k]] += v;
                    }
                }
                __syncwarp(gmask);
                if (inb) {
#pragma unroll
                    for (int k = 0; k < PER; ++k) {
                        double v = vb[k];
                        bad |= !isfinite(v);
                        if (SIGNS) v *= srb * (e2c[(long long)(gl * PER + k) * ntet + eb] < 0 ? -1.0 : 1.0);
                        if (fabs(v) > drop_val) acc[pb[k]] += v;
                    }
                }
                __syncwarp(gmask);
            }
        }
        if (stageA) {
            if (accumulate) for (int s = gl; s < len; s += G) val[p0 + s] += acc[s];
            else for (int s = gl; s < len; s += G) val[p0 + s] = acc[s];
            __syncwarp(gmask);
        }
        if (stageF && gl == 0) {
            if (accumulate) rhs[r] += fsum; else rhs[r] = fsum;
        }
    }
    if (bad) *status = 1;  // benign race: every writer stores the same value
}

template <int G, int NC, bool SIGNS, typename PosT>
cudaError_t launch_cols3(afb_ctx* c, int len_hi, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi) {
    const long long nrows = c->row_end - c->row_begin;
    const int gpb = 256 / G;
    const size_t smem = (size_t)gpb * len_hi * sizeof(double);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nrows + gpb - 1) / gpb, 148LL * 64));
    cudaError_t e = cudaFuncSetAttribute(k_gather_cols<G, NC, SIGNS, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_gather_cols<G, NC, SIGNS, PosT><<<grid, 256, smem, c->stream>>>(nrows, c->ntet, len_hi, len_hi, c->rowptr.as<long long>(), c->radj_ptr.as<long long>(),
                                                                    c->radj.as<unsigned>(), c->pos.as<PosT>(), c->e2r.as<int32_t>(), c->e2c.as<int32_t>(), sA, sF,
                                                                    val, rhs, accumulate, drop, status, e_lo, e_hi);
    return cudaGetLastError();
}
template <int G, int NC>
cudaError_t launch_cols(afb_ctx* c, int len_hi, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi) {
    if (c->has_signs) {
        if (c->pos_bytes == 1) return launch_cols3<G, NC, true, unsigned char>(c, len_hi, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
        return launch_cols3<G, NC, true, unsigned short>(c, len_hi, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
    }
    if (c->pos_bytes == 1) return launch_cols3<G, NC, false, unsigned char>(c, len_hi, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
    return launch_cols3<G, NC, false, unsigned short>(c, len_hi, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
}

template <bool SIGNS, typename PosT>
cudaError_t launch_flat2(afb_ctx* c, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi) {
    const long long nrows = c->row_end - c->row_begin;
    const int wpb = 8;
    const size_t smem = (size_t)wpb * std::max(1, c->max_row_len) * sizeof(double);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nrows + wpb - 1) / wpb, 148LL * 64));
    cudaError_t e = cudaFuncSetAttribute(k_gather_flat<SIGNS, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_gather_flat<SIGNS, PosT><<<grid, 256, smem, c->stream>>>(nrows, c->ntet, c->nrow_loc, c->ncol_loc, std::max(1, c->max_row_len), c->rowptr.as<long long>(),
                                                               c->radj_ptr.as<long long>(), c->radj.as<unsigned>(), c->pos.as<PosT>(), c->e2r.as<int32_t>(),
                                                               c->e2c.as<int32_t>(), sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
    return cudaGetLastError();
}
cudaError_t launch_flat(afb_ctx* c, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi) {
    if (c->has_signs) {
        if (c->pos_bytes == 1) return launch_flat2<true, unsigned char>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
        return launch_flat2<true, unsigned short>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
    }
    if (c->pos_bytes == 1) return launch_flat2<false, unsigned char>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
    return launch_flat2<false, unsigned short>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi);
}

template <int G, bool SIGNS, typename PosT>
cudaError_t launch_g2(afb_ctx* c, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi, int len_lo) {
    const long long nrows = c->row_end - c->row_begin;
    const int gpb = 256 / G;
    const size_t smem = (size_t)gpb * std::max(1, c->max_row_len) * sizeof(double);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nrows + gpb - 1) / gpb, 148LL * 64));
    cudaError_t e = cudaFuncSetAttribute(k_gather<G, SIGNS, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_gather<G, SIGNS, PosT><<<grid, 256, smem, c->stream>>>(nrows, c->ntet, c->nrow_loc, c->ncol_loc, std::max(1, c->max_row_len), c->rowptr.as<long long>(),
                                                             c->radj_ptr.as<long long>(), c->radj.as<unsigned>(), c->pos.as<PosT>(), c->e2r.as<int32_t>(),
                                                             c->e2c.as<int32_t>(), sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi, len_lo);
    return cudaGetLastError();
}

template <int G>
cudaError_t launch_g(afb_ctx* c, const double* sA, const double* sF, double* val, double* rhs, int accumulate, double drop, int* status, long long e_lo, long long e_hi, int len_lo = -1) {
    if (c->has_signs) {
        if (c->pos_bytes == 1) return launch_g2<G, true, unsigned char>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi, len_lo);
        return launch_g2<G, true, unsigned short>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi, len_lo);
    }
    if (c->pos_bytes == 1) return launch_g2<G, false, unsigned char>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi, len_lo);
    return launch_g2<G, false, unsigned short>(c, sA, sF, val, rhs, accumulate, drop, status, e_lo, e_hi, len_lo);
}

}  // namespace

namespace {
// A_in[e][j][i] (column-major nrl x ncl per element, the reference's m_A[j*nRows + i]) -> A_out[e][i][j]
__global__ void k_elem_transpose(long long nel, int nrl, int ncl, const double* __restrict__ in, double* __restrict__ out) {
    const long long n = nel * nrl * ncl;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long e = t / (nrl * ncl);
        const int k = (int)(t - e * nrl * ncl), i = k / ncl, j = k - i * ncl;
        out[t] = in[e * nrl * ncl + (long long)j * nrl + i];
    }
}
__global__ void k_axpy(long long n, const double* __restrict__ x, double* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
int launch_axpy(afb_ctx* ctx, long long n, const double* x, double* y) {
    if (n <= 0) return 0;
    k_axpy<<<(unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 32)), 256, 0, ctx->stream>>>(n, x, y);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : afb::cuda_fail(ctx, e, "k_axpy launch");
}
}  // namespace

namespace afb {

int assemble_tensor_path(afb_ctx* ctx, int nfA, int nfF, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa,
                         const std::vector<OpInfo>& ob, const std::vector<const double*>& Dd, double* dval, double* drhs,
                         int accumulate, double drop_val, int* status_flag, int phase);

int launch_gather(afb_ctx* ctx, const double* stageA, const double* stageF, double* val, double* rhs, int accumulate, double drop_val,
                  int* status_flag, long long e_lo, long long e_hi) {
    // rows per block = 256 / G with the lane-group size G the dispatch below selects (4, 8, 16 or 32 lanes per row)
    const int Gsel = ctx->ncol_loc <= 4 ? 4 : (ctx->ncol_loc <= 8 ? 8 : (ctx->ncol_loc <= 16 ? 16 : 32));
    const size_t need = (size_t)(256 / Gsel) * std::max(1, ctx->max_row_len) * sizeof(double);
    if (need > 200 * 1024) { set_error(ctx, "afb_assemble: matrix rows too long for the shared-memory row image"); return -3; }
    cudaError_t e;
    const int nc = ctx->ncol_loc;
    // P2 / P3 local matrices: rows of up to 128 entries (all edge / face dofs) through k_gather_cols, the long rows after them
    // through the warp-per-row kernel
    if ((nc == 10 || nc == 20) && ctx->nrow_loc == nc && !getenv("AFB_GATHER_GROUPS")) {
        if (getenv("AFB_GATHER_FLAT")) {
            e = launch_flat(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
            ctx->launches++;
            if (e != cudaSuccess) return cuda_fail(ctx, e, "k_gather_flat launch");
            ctx->k_gath = "k_gather_flat";
            return 0;
        }
        const int len_hi = nc == 20 ? 128 : 48;   // 64 KB / 48 KB of row images per 256-thread block
        e = nc == 20 ? launch_cols<4, 20>(ctx, len_hi, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi)
                     : launch_cols<2, 10>(ctx, len_hi, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
        ctx->launches++;
        if (e != cudaSuccess) return cuda_fail(ctx, e, "k_gather_cols launch");
        ctx->k_gath = ctx->max_row_len > len_hi ? "k_gather_cols + k_gather" : "k_gather_cols";
        if (ctx->max_row_len > len_hi) {
            e = nc == 20 ? launch_g<32>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi, len_hi)
                         : launch_g<16>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi, len_hi);
            ctx->launches++;
            if (e != cudaSuccess) return cuda_fail(ctx, e, "k_gather launch");
        }
        return 0;
    }
    if (nc <= 4) e = launch_g<4>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
    else if (nc <= 8) e = launch_g<8>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
    else if (nc <= 16) e = launch_g<16>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
    else e = launch_g<32>(ctx, stageA, stageF, val, rhs, accumulate, drop_val, status_flag, e_lo, e_hi);
    ctx->launches++;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_gather launch");
    ctx->k_gath = "k_gather";
    return 0;
}

}  // namespace afb

extern "C" {

static int assemble_impl(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs, const afb_form* rhs_forms, double* csr_val, double* rhs,
                         int accumulate, double drop_val, int mem_space, int phase_req);

int afb_assemble(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs, const afb_form* rhs_forms,
                 double* csr_val, double* rhs, int accumulate, double drop_val, int mem_space) {
    return assemble_impl(ctx, nforms, forms, nrhs, rhs_forms, csr_val, rhs, accumulate, drop_val, mem_space, 0);
}

int afb_priority_rows_set(afb_ctx* ctx, int64_t first_priority_row) {
    if (!ctx) return -7;
    if (!ctx->has_pattern) { set_error(ctx, "pattern was not built"); return -6; }
    cudaSetDevice(ctx->device);
    ctx->priority_row = first_priority_row;
    if (first_priority_row < 0) { ctx->rp_prio_valid = false; ctx->rg_prio_valid = false; return 0; }
    const int rc = rows_priority_build(ctx, first_priority_row);
    if (rc) return rc;
    return ctx->has_ring_plan ? rings_priority_build(ctx, first_priority_row) : 0;
}

int afb_assemble_phase(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs, const afb_form* rhs_forms, double* csr_val, double* rhs,
                       double drop_val, int phase) {
    if (phase != 1 && phase != 2) { if (ctx) set_error(ctx, "afb_assemble_phase: phase must be 1 or 2"); return -7; }
    return assemble_impl(ctx, nforms, forms, nrhs, rhs_forms, csr_val, rhs, 0, drop_val, AFB_DEVICE, phase);
}

static int assemble_impl(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs, const afb_form* rhs_forms, double* csr_val, double* rhs,
                         int accumulate, double drop_val, int mem_space, int phase_req) {
    if (!ctx) return -7;
    if (phase_req == 2 && ctx->phase_done) { ctx->phase_done = false; return ctx->phase_status; }
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    if (!ctx->has_pattern) { set_error(ctx, "pattern was not built: call afb_pattern_build first"); return -6; }
    if ((nforms > 0 && !forms) || (nrhs > 0 && !rhs_forms) || nforms < 0 || nrhs < 0) { set_error(ctx, "afb_assemble: bad form arrays"); return -7; }
    const bool userA = csr_val != nullptr, doF = rhs != nullptr;
    if (!userA && !doF) return 0;
    // essential boundary conditions (afb_dirichlet_set): the rhs of free rows needs the matrix entries of the Dirichlet
    // columns, so a rhs-only call assembles the matrix into scratch; an accumulating call assembles this call's contribution
    // into scratch, constrains it, then adds it (the reference adds constrained element matrices, assembler.inl:397-425)
    const bool dir = ctx->has_dirichlet;
    const bool doA = userA || (dir && doF && nforms > 0);
    const int user_accumulate = accumulate;
    if (dir) accumulate = 0;
    // phased assembly (afb_assemble_phase): phase 1 runs k_geom + the clusters holding priority rows and returns without
    // synchronising, phase 2 runs the other clusters and reports the status.  Anything the phased cluster gather cannot do is
    // completed in phase 1 (phase 2 then only returns the status).
    const int ph = (phase_req != 0 && !dir && ctx->rp_prio_valid && !ctx->blocks_ready) ? phase_req : 0;
    if (phase_req == 2 && ph == 0) return 0;
    if (doA && nforms == 0 && doF && nrhs == 0) { set_error(ctx, "System local evaluator is not specified"); return -6; }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const long long ntet = ctx->ntet, nrows = ctx->row_end - ctx->row_begin;
    const int nrl = ctx->nrow_loc, ncl = ctx->ncol_loc;
    const int nfA = doA ? nforms : 0, nfF = doF ? nrhs : 0;

    // ---- resolve + validate
    std::vector<OpInfo> oa(nfA + nfF), ob(nfA + nfF);
    std::vector<afb_form> fm(nfA + nfF);
    for (int k = 0; k < nfA + nfF; ++k) {
        fm[k] = k < nfA ? forms[k] : rhs_forms[k - nfA];
        if (resolve_op(fm[k].opA, fm[k].femA, fm[k].vecA, &oa[k]) || resolve_op(fm[k].opB, fm[k].femB, fm[k].vecB, &ob[k])) {
            set_error(ctx, "unsupported operator/space in form");
            return -3;
        }
        if (k >= nfA) {
            if (!(fm[k].opA == AFB_IDEN && fm[k].femA == AFB_FEM_P0 && fm[k].vecA == 1)) { set_error(ctx, "rhs form must use OpA = IDEN(P0)"); return -7; }
            fm[k].col_off = 0;
            if (fm[k].row_off < 0 || fm[k].row_off + ob[k].nfa > nrl) { set_error(ctx, "rhs form block outside the element vector"); return -7; }
        } else if (fm[k].row_off < 0 || fm[k].col_off < 0 || fm[k].row_off + ob[k].nfa > nrl || fm[k].col_off + oa[k].nfa > ncl) {
            set_error(ctx, "form block outside the element matrix");
            return -7;
        }
    }
    // ---- coefficients to the device
    std::vector<const double*> Dd(nfA + nfF, nullptr);
    {
        size_t total = 0;
        std::vector<size_t> offs(nfA + nfF, 0), sizes(nfA + nfF, 0);
        for (int k = 0; k < nfA + nfF; ++k) {
            const int dlen = form_dlen(fm[k], oa[k], ob[k]);
            const int q = afb_tet_quadrature(fm[k].quad_order, nullptr, nullptr, 0);
            if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
            const size_t n = fm[k].coef_layout == AFB_COEF_CONST ? 1 : (fm[k].coef_layout == AFB_COEF_PER_TET ? (size_t)ntet : (size_t)ntet * q);
            sizes[k] = dlen ? n * dlen : 0;
            if (sizes[k] && !fm[k].D) { set_error(ctx, "tensor data missing"); return -7; }
            if (sizes[k] && fm[k].coef_space == AFB_HOST) { offs[k] = total; total += (sizes[k] + 1) & ~(size_t)1; }
        }
        if (total) AFB_CUDA(ctx, ctx->coef.reserve(total * sizeof(double)));
        cudaEventRecord(ctx->ev[0], st);
        for (int k = 0; k < nfA + nfF; ++k) {
            if (!sizes[k]) continue;
            if (fm[k].coef_space == AFB_HOST) {
                AFB_CUDA(ctx, cudaMemcpyAsync(ctx->coef.as<double>() + offs[k], fm[k].D, sizes[k] * sizeof(double), cudaMemcpyHostToDevice, st));
                Dd[k] = ctx->coef.as<double>() + offs[k];
            } else Dd[k] = fm[k].D;
        }
    }
    // ---- output buffers (device images of csr_val / rhs when the caller's live on the host)
    double* dval = csr_val;
    double* drhs = rhs;
    double *uval = nullptr, *urhs = nullptr;  // device images of the caller's arrays when this call's contribution goes to scratch
    if (mem_space == AFB_HOST) {
        if (userA) {
            AFB_CUDA(ctx, ctx->io_val.reserve(std::max<long long>(1, ctx->nnz) * sizeof(double)));
            dval = ctx->io_val.as<double>();
            if (user_accumulate) AFB_CUDA(ctx, cudaMemcpyAsync(dval, csr_val, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (doF) {
            AFB_CUDA(ctx, ctx->io_rhs.reserve(std::max<long long>(1, nrows) * sizeof(double)));
            drhs = ctx->io_rhs.as<double>();
            if (user_accumulate) AFB_CUDA(ctx, cudaMemcpyAsync(drhs, rhs, nrows * sizeof(double), cudaMemcpyHostToDevice, st));
        }
    }
    if (dir && (user_accumulate || !userA)) {
        if (doA) {
            AFB_CUDA(ctx, ctx->tmp2.reserve(std::max<long long>(1, ctx->nnz) * sizeof(double)));
            uval = userA ? dval : nullptr;
            dval = ctx->tmp2.as<double>();
        }
        if (doF && user_accumulate) {
            AFB_CUDA(ctx, ctx->tmp3.reserve(std::max<long long>(1, nrows) * sizeof(double)));
            urhs = drhs;
            drhs = ctx->tmp3.as<double>();
        }
    }
    AFB_CUDA(ctx, ctx->flag.reserve(64));
    if (ph != 2) AFB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, 64, st));

    // ---- fused fast path (afb_tensor.cu): element-wise constant coefficients on scalar P0..P3 spaces
    //      and, for vector / mixed spaces numbered by afb_dofmap_natural, the same path block by block (afb_blocks.cu)
    // an output without forms (e.g. Assemble(A, b) of a problem without load) is not touched by the fused kernels
    double* fval = (doA && nfA > 0) ? dval : nullptr;
    double* frhs = (doF && nfF > 0) ? drhs : nullptr;
    int handled = ph ? 0 : assemble_block_path(ctx, nfA, nfF, fm, oa, ob, Dd, fval, frhs, accumulate, drop_val, ctx->flag.as<int>());
    if (handled < 0) return handled;
    if (!handled) handled = assemble_tensor_path(ctx, nfA, nfF, fm, oa, ob, Dd, fval, frhs, accumulate, drop_val, ctx->flag.as<int>(), ph);
    if (handled < 0) return handled;
    if (ph == 1 && (handled == 2 || handled == 3)) { ctx->phase_done = false; return 0; }   // phase 2 follows
    if (ph == 2 && handled != 2 && handled != 3) { set_error(ctx, "afb_assemble_phase: phase 2 does not match phase 1"); return -6; }
    if (handled && !accumulate) {
        if (doA && !fval && ctx->nnz) AFB_CUDA(ctx, cudaMemsetAsync(dval, 0, ctx->nnz * sizeof(double), st));
        if (doF && !frhs && nrows) AFB_CUDA(ctx, cudaMemsetAsync(drhs, 0, nrows * sizeof(double), st));
    }
    if (!handled) {
    // ---- generic path: stage full element matrices, then gather
    // the staged matrices of all elements may not fit (P2^3 x P1 at 12.6 M tets = 116 GB): elements are processed in chunks
    // of bounded staging size, every chunk gathered with accumulate (the row sums stay in ascending element order)
    double *sA = nullptr, *sF = nullptr;
    // Every chunk after the first walks all rows again and adds into the CSR values (read + write), so one chunk is the goal:
    // the staging buffer may take what is free on the device minus a reserve (P3 at 10.1 M tets: 32 GB of 180 GB).
    size_t stage_cap = (size_t)8 << 30;
    {
        size_t mfree = 0, mtotal = 0;
        if (cudaMemGetInfo(&mfree, &mtotal) == cudaSuccess) {
            const size_t have = ctx->stageA.cap + ctx->stageF.cap;   // already reserved by an earlier call
            const size_t avail = mfree + have;
            if (avail > ((size_t)12 << 30)) stage_cap = std::max(stage_cap, avail - ((size_t)8 << 30));
        }
    }
    if (const char* sc = getenv("AFB_STAGE_BYTES")) stage_cap = (size_t)std::max(1LL, atoll(sc));
    const size_t per_elem = ((doA ? (size_t)nrl * ncl : 0) + (doF ? (size_t)nrl : 0)) * sizeof(double);
    const long long chunk = std::max<long long>(1, std::min<long long>(ntet, (long long)(stage_cap / std::max<size_t>(1, per_elem))));
    if (doA) {
        AFB_CUDA(ctx, ctx->stageA.reserve((size_t)chunk * nrl * ncl * sizeof(double)));
        sA = ctx->stageA.as<double>();
    }
    if (doF) {
        AFB_CUDA(ctx, ctx->stageF.reserve((size_t)chunk * nrl * sizeof(double)));
        sF = ctx->stageF.as<double>();
    }
    // which launches may store and which must add: a block region already written -> add; uncovered area -> memset
    auto plan = [&](int k0, int k1, bool matrix, std::vector<int>& add, bool& need_zero) {
        long long covered = 0;
        bool partial = false;
        std::vector<int> seen;
        for (int k = k0; k < k1; ++k) {
            const int r0 = fm[k].row_off, r1 = r0 + ob[k].nfa, c0 = matrix ? fm[k].col_off : 0, c1 = matrix ? c0 + oa[k].nfa : 1;
            bool same = false;
            for (int s : seen) {
                const int sr0 = fm[s].row_off, sr1 = sr0 + ob[s].nfa, sc0 = matrix ? fm[s].col_off : 0, sc1 = matrix ? sc0 + oa[s].nfa : 1;
                if (sr0 == r0 && sr1 == r1 && sc0 == c0 && sc1 == c1) same = true;
                else if (r0 < sr1 && sr0 < r1 && c0 < sc1 && sc0 < c1) partial = true;
            }
            add[k] = same ? 1 : 0;
            if (!same) { covered += (long long)(r1 - r0) * (c1 - c0); seen.push_back(k); }
        }
        const long long area = (long long)nrl * (matrix ? ncl : 1);
        need_zero = partial || covered < area;
        if (partial) for (int k = k0; k < k1; ++k) add[k] = 1;
    };
    std::vector<int> add(nfA + nfF, 0);
    bool zeroA = false, zeroF = false;
    // forms that are square on one scalar space with the same operator on both sides and cover the whole local matrix go
    // through the register-tiled kernel k_element_sq in one launch (the FP64 contraction-bound case: P3, many points)
    std::vector<int> sq;
    if (doA && nrl == ncl && (nrl == 4 || nrl == 10 || nrl == 20) && !getenv("AFB_DISABLE_SQ_ELEMENT")) {
        for (int k = 0; k < nfA && (int)sq.size() < 4; ++k) {
            const afb_form& f = fm[k];
            const int dl = form_dlen(f, oa[k], ob[k]);
            const bool same = oa[k].vec == 1 && ob[k].vec == 1 && oa[k].fem == ob[k].fem && f.opA == f.opB && (f.opA == AFB_GRAD || f.opA == AFB_IDEN);
            const bool dims_ok = f.tensor_type < AFB_TENSOR_SYMMETRIC || dl == (f.opA == AFB_GRAD ? 9 : 1);
            if (same && dims_ok && oa[k].nfa == nrl && f.row_off == 0 && f.col_off == 0 && (sq.empty() || oa[k].fem == oa[sq[0]].fem)) sq.push_back(k);
        }
    }
    if (doA) {
        plan(0, nfA, true, add, zeroA);
        if (!sq.empty()) {  // the tiled launch stores the full matrix first; every other matrix form adds to it
            zeroA = false;
            for (int k = 0; k < nfA; ++k) add[k] = 1;
        }
    }
    if (doF) plan(nfA, nfA + nfF, false, add, zeroF);
    cudaEventRecord(ctx->ev[1], st);
    for (long long e_lo = 0; e_lo < ntet; e_lo += chunk) {
        const long long e_hi = std::min<long long>(ntet, e_lo + chunk), nel = e_hi - e_lo;
        if (doA && zeroA) AFB_CUDA(ctx, cudaMemsetAsync(sA, 0, (size_t)nel * nrl * ncl * sizeof(double), st));
        if (doF && zeroF) AFB_CUDA(ctx, cudaMemsetAsync(sF, 0, (size_t)nel * nrl * sizeof(double), st));
        // per-element coefficient data follow the element index
        std::vector<const double*> Dc(Dd);
        for (int k = 0; k < nfA + nfF; ++k) {
            if (!Dc[k] || fm[k].coef_layout == AFB_COEF_CONST) continue;
            const int dl = form_dlen(fm[k], oa[k], ob[k]);
            const int qk = fm[k].coef_layout == AFB_COEF_PER_POINT ? afb_tet_quadrature(fm[k].quad_order, nullptr, nullptr, 0) : 1;
            Dc[k] += (size_t)dl * qk * e_lo;
        }
        // ---- K1: element blocks
        ctx->k_elem = "k_element_generic";
        if (!sq.empty()) {
            int rc = launch_forms_sq(ctx, fm, oa, Dc, sq, e_lo, nel, sA, (long long)nrl * ncl);
            if (rc) return rc;
        }
        for (int k = 0; k < nfA + nfF; ++k) {
            if (std::find(sq.begin(), sq.end(), k) != sq.end()) continue;
            const bool matrix = k < nfA;
            int rc = launch_form(ctx, fm[k], oa[k], ob[k], nel, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(),
                                 ctx->v[0].as<int32_t>() + e_lo, ctx->v[1].as<int32_t>() + e_lo, ctx->v[2].as<int32_t>() + e_lo,
                                 ctx->v[3].as<int32_t>() + e_lo, nullptr, matrix ? sA : sF, matrix ? (long long)nrl * ncl : nrl, matrix ? ncl : 1,
                                 matrix ? 1 : 0, add[k], Dc[k]);
            if (rc) return rc;
        }
        if (e_hi == ntet) cudaEventRecord(ctx->ev[2], st);
        // ---- K3: gather into CSR / rhs
        int rc = launch_gather(ctx, doA ? sA : nullptr, doF ? sF : nullptr, dval, drhs, e_lo == 0 ? accumulate : 1, drop_val,
                               ctx->flag.as<int>(), e_lo, e_hi);
        if (rc) return rc;
    }
    cudaEventRecord(ctx->ev[3], st);
    }
    if (dir) {
        int rcd = dirichlet_apply(ctx, doA ? dval : nullptr, doF ? drhs : nullptr);
        if (rcd) return rcd;
        if (uval) { rcd = launch_axpy(ctx, ctx->nnz, dval, uval); if (rcd) return rcd; dval = uval; }
        if (urhs) { rcd = launch_axpy(ctx, nrows, drhs, urhs); if (rcd) return rcd; drhs = urhs; }
    }
    if (mem_space == AFB_HOST) {
        if (userA && ctx->nnz) AFB_CUDA(ctx, cudaMemcpyAsync(csr_val, dval, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (doF && nrows) AFB_CUDA(ctx, cudaMemcpyAsync(rhs, drhs, nrows * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int bad = 0;
    AFB_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t23, ctx->ev[2], ctx->ev[3]);
    ctx->times[0] = t12; ctx->times[1] = t23; ctx->times[2] = t01; ctx->times[3] = (double)handled;
    const int status = bad ? -1 : 0;
    if (bad) set_error(ctx, "not a number in local matrix or rhs");
    if (phase_req == 1) { ctx->phase_done = true; ctx->phase_status = status; }   // everything was done in phase 1
    return status;
}

// Scatter of element matrices the CALLER evaluated (the MatFuncWrap plug-in point: a host lambda per cell, func_wrap.h:96-187,
// installed with AssemblerT::SetMatRHSFunc, assembler.h:326-328): A_elem holds, for the cells [e_lo, e_lo + nel) of the context's
// mesh, the local matrices in the reference's layout m_A[j*nRows + i] (column-major nrow_loc x ncol_loc, assembler.inl:417),
// F_elem the local right-hand sides.  They are ADDED into csr_val / rhs with the reference's rule (|A| > drop_val, signs of the
// index codes, non-finite value -> -1); the kernels are the row gather of the generic path.
int afb_assemble_elemental(afb_ctx* ctx, int64_t e_lo, int64_t nel, const double* A_elem, const double* F_elem, int elem_space,
                           double* csr_val, double* rhs, double drop_val, int mem_space) {
    if (!ctx) return -7;
    if (!ctx->has_dofmap) { set_error(ctx, "Description of fem expression is empty (no dof map)"); return -6; }
    if (!ctx->has_pattern) { set_error(ctx, "pattern was not built (afb_pattern_build)"); return -6; }
    if (e_lo < 0 || nel < 0 || e_lo + nel > ctx->ntet || (!A_elem && !F_elem) || (A_elem && !csr_val) || (F_elem && !rhs)) {
        set_error(ctx, "afb_assemble_elemental: bad arguments");
        return -7;
    }
    if (nel == 0) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_error(ctx, "no CUDA device (the library has no CPU fallback)"); return -4; }
    cudaStream_t st = ctx->stream;
    const int nrl = ctx->nrow_loc, ncl = ctx->ncol_loc;
    const long long nrows = ctx->row_end - ctx->row_begin;
    const cudaMemcpyKind kin = elem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    double *sA = nullptr, *sF = nullptr;
    if (A_elem) {
        const size_t n = (size_t)nel * nrl * ncl;
        AFB_CUDA(ctx, ctx->tmp2.reserve(n * sizeof(double)));
        AFB_CUDA(ctx, ctx->stageA.reserve(n * sizeof(double)));
        AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp2.p, A_elem, n * sizeof(double), kin, st));
        sA = ctx->stageA.as<double>();
        // column-major (reference) -> row-major staging of the gather
        k_elem_transpose<<<(unsigned)std::max<long long>(1, std::min<long long>(((long long)n + 255) / 256, 148LL * 32)), 256, 0, st>>>(
            (long long)nel, nrl, ncl, ctx->tmp2.as<double>(), sA);
        ctx->launches++;
    }
    if (F_elem) {
        const size_t n = (size_t)nel * nrl;
        AFB_CUDA(ctx, ctx->stageF.reserve(n * sizeof(double)));
        AFB_CUDA(ctx, cudaMemcpyAsync(ctx->stageF.p, F_elem, n * sizeof(double), kin, st));
        sF = ctx->stageF.as<double>();
    }
    double *dval = csr_val, *drhs = rhs;
    if (mem_space == AFB_HOST) {
        if (sA) {
            AFB_CUDA(ctx, ctx->io_val.reserve(std::max<long long>(1, ctx->nnz) * sizeof(double)));
            dval = ctx->io_val.as<double>();
            AFB_CUDA(ctx, cudaMemcpyAsync(dval, csr_val, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (sF) {
            AFB_CUDA(ctx, ctx->io_rhs.reserve(std::max<long long>(1, nrows) * sizeof(double)));
            drhs = ctx->io_rhs.as<double>();
            AFB_CUDA(ctx, cudaMemcpyAsync(drhs, rhs, nrows * sizeof(double), cudaMemcpyHostToDevice, st));
        }
    }
    AFB_CUDA(ctx, ctx->flag.reserve(64));
    AFB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    const int rc = launch_gather(ctx, sA, sF, sA ? dval : nullptr, sF ? drhs : nullptr, 1, drop_val, ctx->flag.as<int>(), e_lo, e_lo + nel);
    if (rc) return rc;
    if (mem_space == AFB_HOST) {
        if (sA) AFB_CUDA(ctx, cudaMemcpyAsync(csr_val, dval, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (sF) AFB_CUDA(ctx, cudaMemcpyAsync(rhs, drhs, nrows * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int bad = 0;
    AFB_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    if (bad) set_error(ctx, "not a number in local matrix or rhs");
    return bad ? -1 : 0;
}

/* names of the element / gather kernels of the generic staged path that ran last (bench evidence), "elem|gather" */
int afb_last_kernels(afb_ctx* ctx, char* buf, int capacity) {
    if (!ctx || !buf || capacity <= 0) return -7;
    snprintf(buf, (size_t)capacity, "%s|%s", ctx->k_elem, ctx->k_gath);
    return 0;
}

int afb_last_times(afb_ctx* ctx, double* ms4) {
    if (!ctx || !ms4) return -7;
    for (int i = 0; i < 4; ++i) ms4[i] = ctx->times[i];
    return 0;
}

}  // extern "C"
