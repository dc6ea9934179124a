// Row-stationary gather of the fused tensor-representation path: one THREAD owns one CSR row (k_rows_ell), plus the
// class-sorted sliced-ELL plan it streams.
//
// Reference semantics reproduced: the scatter of AssemblerT::Assemble (inmost_interface/assembler.inl:397-481)
//     matrix[r][c] += A_e(i,j) if |A_e(i,j)| > drop_val;  rhs[r] += F_e(i);  non-finite local value -> status -1
// for element matrices A_e(i,j) = sum_c T[c][i][j] g_e[c] (afb_tensor.cu explains the factorisation of fem3Dtet).
//
// Why this shape (ncu of the lane-group kernel k_gather_tensor_sq, profiles/r01b_*.md: the LSU data pipe was 96 % busy,
// 60 % of it table reads from shared memory):
//   * visits (element, local row i) of a row are ordered by i and the rows of a warp ("slice" of 32 rows with similar
//     class signature) step through the classes together, padded to the slice maximum.  i is therefore warp-uniform
//     and compile-time inside the unrolled class loop, so the table T[.][i][j] is read as immediate-offset constant-bank
//     operands of the DFMAs (the table travels as a __grid_constant__ kernel parameter): no table loads at all.
//   * thread-per-row accumulators live in shared memory as acc[slot][lane]: every lane touches only its own bank pair,
//     so the 10 read-modify-writes of a visit are conflict-free whatever the slots are.
//   * the plan is streamed coalesced: per visit-step and lane NW 32-bit words (the NLOC slot bytes + the element id).
//   * finished rows are transposed through a small padded shared-memory tile and stored with lanes <-> slots.
// Deterministic: the summation order of a row is fixed by the plan (class, then ascending element); no atomics.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "afb_internal.h"

using namespace afb;

namespace {

constexpr int NBUCKET = 9;
__host__ __device__ inline int bucket_L(int b) {
    switch (b) { case 0: return 20; case 1: return 28; case 2: return 36; case 3: return 48; case 4: return 66;
                 case 5: return 96; case 6: return 128; case 7: return 192; default: return 256; }
}
__host__ __device__ inline int len_bucket(int len) {
    int b = 0;
    while (b < NBUCKET - 1 && len > bucket_L(b)) ++b;
    return b;
}

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

// ---------------------------------------------------------------------------------------------------------------------
// plan kernels
// ---------------------------------------------------------------------------------------------------------------------
template <int NLOC>
__device__ __forceinline__ void class_counts(const unsigned* __restrict__ radj, long long a0, long long a1, int* cnt) {
#pragma unroll
    for (int c = 0; c < NLOC; ++c) cnt[c] = 0;
    for (long long a = a0; a < a1; ++a) {
        const int i = (int)(radj[a] % (unsigned)NLOC);
#pragma unroll
        for (int c = 0; c < NLOC; ++c) cnt[c] += (i == c);
    }
}

// sort key of a row: (length bucket, degree, clipped class counts) -> rows with the same class signature become
// neighbours, ties keep the ascending row order (stable sort) which preserves the locality of the numbering
template <int NLOC>
__global__ void k_row_key(long long nrows, const long long* __restrict__ rowptr, const long long* __restrict__ radj_ptr,
                          const unsigned* __restrict__ radj, unsigned long long* key, unsigned* rowid) {
    constexpr int BITS = (50 / NLOC) > 8 ? 8 : (50 / NLOC);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x) {
        const int len = (int)(rowptr[r + 1] - rowptr[r]);
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        int cnt[NLOC];
        class_counts<NLOC>(radj, a0, a1, cnt);
        unsigned long long pack = 0;
#pragma unroll
        for (int c = 0; c < NLOC; ++c) pack = (pack << BITS) | (unsigned long long)min(cnt[c], (1 << BITS) - 1);
        const unsigned long long deg = (unsigned long long)min((long long)255, a1 - a0);
        key[r] = ((unsigned long long)len_bucket(len) << 58) | (deg << 50) | pack;
        rowid[r] = (unsigned)r;
    }
}

// one warp per slice: class counts padded to the slice maximum, number of visit-steps, longest row
template <int NLOC>
__global__ void k_slice_info(long long nrows, long long nslices, const unsigned* __restrict__ order, const long long* __restrict__ rowptr,
                             const long long* __restrict__ radj_ptr, const unsigned* __restrict__ radj, unsigned short* cntS,
                             long long* steps, int* slen, int* overflow) {
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = wglob; s < nslices; s += nw) {
        const long long ri = s * 32 + lane;
        int cnt[NLOC], len = 0;
        if (ri < nrows) {
            const long long r = order[ri];
            len = (int)(rowptr[r + 1] - rowptr[r]);
            class_counts<NLOC>(radj, radj_ptr[r], radj_ptr[r + 1], cnt);
        } else {
#pragma unroll
            for (int c = 0; c < NLOC; ++c) cnt[c] = 0;
        }
        long long tot = 0;
#pragma unroll
        for (int c = 0; c < NLOC; ++c) {
            const int m = __reduce_max_sync(0xffffffffu, cnt[c]);
            if (m > 65535 && lane == 0) *overflow = 1;
            if (lane == 0) cntS[s * NLOC + c] = (unsigned short)m;
            tot += m;
        }
        const int ml = __reduce_max_sync(0xffffffffu, len);
        if (lane == 0) { steps[s] = tot; slen[s] = ml; }
    }
}

// fills the ELL stream: word w of visit-step t of lane l at ell[(t*NW + w)*32 + l]; words 0..NWP-1 = slot bytes of the
// NLOC columns, word NWP = element id + 1 (0 = padding visit; the buffer is pre-set to 0, so padding visits address slot 0)
template <int NLOC>
__global__ void k_ell_fill(long long nrows, long long nslices, const unsigned* __restrict__ order, const long long* __restrict__ radj_ptr,
                           const unsigned* __restrict__ radj, const unsigned char* __restrict__ pos, const unsigned short* __restrict__ cntS,
                           const long long* __restrict__ sptr, unsigned* ell) {
    constexpr int NWP = (NLOC + 3) / 4, NW = NWP + 1;
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = wglob; s < nslices; s += nw) {
        const long long ri = s * 32 + lane;
        if (ri >= nrows) continue;
        const long long r = order[ri];
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        long long base = sptr[s];
        for (int c = 0; c < NLOC; ++c) {
            int k = 0;
            for (long long a = a0; a < a1; ++a) {
                const unsigned t = radj[a];
                if ((int)(t % (unsigned)NLOC) != c) continue;
                unsigned* dst = ell + (size_t)(base + k) * NW * 32 + lane;
                const unsigned char* pa = pos + (size_t)a * NLOC;
#pragma unroll
                for (int w = 0; w < NWP; ++w) {
                    unsigned word = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (4 * w + b < NLOC) word |= (unsigned)pa[4 * w + b] << (8 * b);
                    dst[w * 32] = word;
                }
                dst[NWP * 32] = t / (unsigned)NLOC + 1u;
                ++k;
            }
            base += cntS[s * NLOC + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the gather
// ---------------------------------------------------------------------------------------------------------------------
template <int NLOC, int NGA, int NGF>
struct RowTab {
    double A[NGA > 0 ? NLOC * NLOC * NGA : 1];  // [(i*NLOC + j)*NGA + c]
    double F[NGF > 0 ? NLOC * NGF : 1];         // [i*NGF + c]
};

struct RowsArgs {
    long long s0, s1;   // slice range of this launch (one length bucket)
    int L;              // slots per row image in this launch
    long long nrows;
    const unsigned* order;
    const unsigned short* cnt;
    const long long* sptr;
    const unsigned* ell;
    const long long* rowptr;
    const double* gbuf;
    double* val;
    double* rhs;
    int accumulate;
    long long dropbits;  // bit pattern of drop_val (>= 0), -1 when drop_val < 0: |v| > drop  <=>  bits(|v|) > dropbits
    int* status;
};

constexpr int RW_CH = 16;                                        // slots per transposition tile
constexpr int RW_EXTRA = 32 * (RW_CH + 1) + 32 + 16;             // doubles per warp besides the row images
inline size_t rows_smem_per_warp(int L) { return ((size_t)L * 32 + RW_EXTRA) * sizeof(double); }

__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

template <int NLOC, int NGA, int NGF>
__global__ void __launch_bounds__(128) k_rows_ell(const __grid_constant__ RowTab<NLOC, NGA, NGF> T, const RowsArgs p) {
    constexpr int NWP = (NLOC + 3) / 4, NW = NWP + 1;
    constexpr int NG = NGA + NGF, NGP = (NG + 1) & ~1;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long s = p.s0 + (long long)blockIdx.x * wpb + warp;
    if (s >= p.s1) return;  // warps are independent: no block-level barrier below
    double* wbase = sm + (size_t)warp * ((size_t)p.L * 32 + RW_EXTRA);
    double* tr = wbase + (size_t)p.L * 32;                                   // [32][RW_CH+1]
    long long* sp0 = reinterpret_cast<long long*>(tr + 32 * (RW_CH + 1));    // [32]
    int* slen = reinterpret_cast<int*>(sp0 + 32);                            // [32]
    const unsigned acc_a = (unsigned)__cvta_generic_to_shared(wbase) + lane * 8;

    const long long ri = s * 32 + lane;
    const bool row_on = ri < p.nrows;
    long long r = 0, p0 = 0;
    int len = 0;
    if (row_on) {
        r = p.order[ri];
        p0 = p.rowptr[r];
        len = (int)(p.rowptr[r + 1] - p0);
    }
    const int Lw = __reduce_max_sync(0xffffffffu, len);
    const bool doA = NGA > 0 && p.val != nullptr, doF = NGF > 0 && p.rhs != nullptr;
    if (doA)
        for (int sl = 0; sl < Lw; ++sl) sts64(acc_a + sl * 256, 0.0);

    long long step = p.sptr[s];
    const long long end = p.sptr[s + 1];
    const unsigned short* cn = p.cnt + (size_t)s * NLOC;
    const unsigned* ellp = p.ell + lane;
    const double* __restrict__ gbuf = p.gbuf;

    unsigned w[NW], w1[NW];
    double g[NGP];
#pragma unroll
    for (int k = 0; k < NW; ++k) { w[k] = 0u; w1[k] = 0u; }
#pragma unroll
    for (int c = 0; c < NGP; ++c) g[c] = 0.0;
    if (step < end) {
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = __ldg(ellp + ((size_t)step * NW + k) * 32);
        if (w[NW - 1] != 0u) {
            const double2* ge = reinterpret_cast<const double2*>(gbuf + (size_t)(w[NW - 1] - 1u) * NGP);
#pragma unroll
            for (int c = 0; c < NGP / 2; ++c) { const double2 d = __ldg(ge + c); g[2 * c] = d.x; g[2 * c + 1] = d.y; }
        }
    }
    if (step + 1 < end) {
#pragma unroll
        for (int k = 0; k < NW; ++k) w1[k] = __ldg(ellp + ((size_t)(step + 1) * NW + k) * 32);
    }
    double fsum = 0.0;
    unsigned nonfin = 0;
#pragma unroll
    for (int i = 0; i < NLOC; ++i) {
        const int n = cn[i];
        for (int k = 0; k < n; ++k) {
            // prefetch: coefficients of the next visit, plan words of the one after
            unsigned w2[NW];
            double g1[NGP];
#pragma unroll
            for (int c = 0; c < NGP; ++c) g1[c] = 0.0;
            if (w1[NW - 1] != 0u) {
                const double2* ge = reinterpret_cast<const double2*>(gbuf + (size_t)(w1[NW - 1] - 1u) * NGP);
#pragma unroll
                for (int c = 0; c < NGP / 2; ++c) { const double2 d = __ldg(ge + c); g1[2 * c] = d.x; g1[2 * c + 1] = d.y; }
            }
#pragma unroll
            for (int q = 0; q < NW; ++q) w2[q] = 0u;
            if (step + 2 < end) {
#pragma unroll
                for (int q = 0; q < NW; ++q) w2[q] = __ldg(ellp + ((size_t)(step + 2) * NW + q) * 32);
            }
            const bool on = w[NW - 1] != 0u;
            if (doF) {
                double f = 0.0;
#pragma unroll
                for (int c = 0; c < NGF; ++c) f = fma(T.F[i * NGF + c], g[NGA + c], f);
                nonfin = max(nonfin, (unsigned)__double2hiint(f) & 0x7fffffffu);
                fsum += f;  // padding visits carry g = 0
            }
            if (doA) {
                double v[NLOC], o[NLOC];
                unsigned sa[NLOC];
#pragma unroll
                for (int j = 0; j < NLOC; ++j) {
                    const unsigned slot = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    sa[j] = acc_a + slot * 256;
                }
#pragma unroll
                for (int j = 0; j < NLOC; ++j) o[j] = lds64(sa[j]);  // the slots of one visit are distinct (checked by the plan)
#pragma unroll
                for (int j = 0; j < NLOC; ++j) {
                    double x = 0.0;
#pragma unroll
                    for (int c = 0; c < NGA; ++c) x = fma(T.A[(i * NLOC + j) * NGA + c], g[c], x);
                    v[j] = x;
                }
#pragma unroll
                for (int j = 0; j < NLOC; ++j) {
                    const long long b = __double_as_longlong(v[j]) & 0x7fffffffffffffffLL;
                    nonfin = max(nonfin, (unsigned)(b >> 32));
                    if (on && b > p.dropbits) sts64(sa[j], o[j] + v[j]);
                }
            }
#pragma unroll
            for (int q = 0; q < NW; ++q) { w[q] = w1[q]; w1[q] = w2[q]; }
#pragma unroll
            for (int c = 0; c < NGP; ++c) g[c] = g1[c];
            ++step;
        }
    }

    // ---- write-out: slot-synchronous read of the row images, transposition through tr, lanes <-> slots stores
    if (doA) {
        sp0[lane] = p0;
        slen[lane] = len;
        __syncwarp();
        for (int s0 = 0; s0 < Lw; s0 += RW_CH) {
#pragma unroll
            for (int t = 0; t < RW_CH; ++t)
                if (s0 + t < Lw) tr[lane * (RW_CH + 1) + t] = lds64(acc_a + (s0 + t) * 256);
            __syncwarp();
#pragma unroll 4
            for (int it = 0; it < RW_CH; ++it) {
                const int rl = it * 2 + (lane >> 4), t = lane & 15;
                const int sl = s0 + t;
                if (sl < slen[rl]) {
                    const double x = tr[rl * (RW_CH + 1) + t];
                    double* d = p.val + sp0[rl] + sl;
                    if (p.accumulate) *d += x; else *d = x;
                }
            }
            __syncwarp();
        }
    }
    if (doF && row_on) {
        if (p.accumulate) p.rhs[r] += fsum; else p.rhs[r] = fsum;
    }
    if (nonfin >= 0x7ff00000u) *p.status = 1;  // benign race: every writer stores the same value
}

template <int NLOC, int NGA, int NGF>
int launch_rows_t(afb_ctx* ctx, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs, int accumulate,
                  double drop_val, int* status) {
    static RowTab<NLOC, NGA, NGF> T;  // host staging of the parameter (copied by value at launch)
    // TA is [c][i][j], TF is [c][i]
    for (int i = 0; i < NLOC; ++i)
        for (int j = 0; j < NLOC; ++j)
            for (int c = 0; c < NGA; ++c) T.A[(i * NLOC + j) * NGA + c] = TA[((size_t)c * NLOC + i) * NLOC + j];
    for (int i = 0; i < NLOC; ++i)
        for (int c = 0; c < NGF; ++c) T.F[i * NGF + c] = TF[(size_t)c * NLOC + i];
    RowsArgs p;
    p.nrows = ctx->row_end - ctx->row_begin;
    p.order = ctx->rp_order.as<unsigned>(); p.cnt = ctx->rp_cnt.as<unsigned short>(); p.sptr = ctx->rp_sptr.as<long long>();
    p.ell = ctx->rp_ell.as<unsigned>(); p.rowptr = ctx->rowptr.as<long long>(); p.gbuf = gbuf;
    p.val = val; p.rhs = rhs; p.accumulate = accumulate; p.status = status;
    if (drop_val < 0) p.dropbits = -1;
    else std::memcpy(&p.dropbits, &drop_val, sizeof(double));
    auto kern = k_rows_ell<NLOC, NGA, NGF>;
    for (const auto& b : ctx->rp_buckets) {
        if (b.s1 <= b.s0) continue;
        int wpb = 4;
        while (wpb > 1 && rows_smem_per_warp(b.L) * wpb > 200 * 1024) wpb >>= 1;
        const size_t smem = rows_smem_per_warp(b.L) * wpb;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024));
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_rows_ell)");
        p.s0 = b.s0; p.s1 = b.s1; p.L = b.L;
        const long long grid = (b.s1 - b.s0 + wpb - 1) / wpb;
        kern<<<(unsigned)grid, wpb * 32, smem, ctx->stream>>>(T, p);
        ctx->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(ctx, e, "k_rows_ell launch");
    }
    return 1;
}

template <int NLOC>
int launch_rows_n(afb_ctx* ctx, int nga, int ngf, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs,
                  int accumulate, double drop_val, int* status) {
#define RW(A, F) if (nga == A && ngf == F) return launch_rows_t<NLOC, A, F>(ctx, TA, TF, gbuf, val, rhs, accumulate, drop_val, status);
    RW(6, 1) RW(6, 0) RW(7, 1) RW(7, 0) RW(1, 1) RW(1, 0) RW(0, 1)
    if constexpr (NLOC <= 10) { RW(9, 1) RW(9, 0) RW(10, 1) RW(10, 0) }
#undef RW
    return 0;
}

}  // namespace

namespace afb {

// Builds the class-sorted sliced-ELL plan from the adjacency lists and the slot table (afb_pattern.cu).  Returns 0; the plan
// is simply absent (has_rows_plan = false -> lane-group gather) when the dof map is not one of the supported shapes.
int build_rows_plan(afb_ctx* ctx) {
    ctx->has_rows_plan = false;
    ctx->rp_buckets.clear();
    const int nl = ctx->nrow_loc;
    if (getenv("AFB_DISABLE_ROWS_PLAN")) return 0;
    if (ctx->nrow_loc != ctx->ncol_loc || !(nl == 4 || nl == 10 || nl == 20)) return 0;
    if (ctx->pos_bytes != 1 || ctx->has_signs || ctx->pos_has_dup) return 0;
    const long long nrows = ctx->row_end - ctx->row_begin;
    if (nrows <= 0 || ctx->n_adj <= 0) return 0;
    const long long nslices = (nrows + 31) / 32;
    cudaStream_t st = ctx->stream;
    DevBuf key, key2, rid, cubtmp, steps, slen;
    auto cleanup = [&]() { key.release(); key2.release(); rid.release(); cubtmp.release(); steps.release(); slen.release(); };
#define R_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
    R_CUDA(key.reserve(nrows * 8)); R_CUDA(key2.reserve(nrows * 8)); R_CUDA(rid.reserve(nrows * 4));
    R_CUDA(ctx->rp_order.reserve(nrows * 4));
    R_CUDA(ctx->rp_cnt.reserve((size_t)nslices * nl * sizeof(unsigned short)));
    R_CUDA(ctx->rp_sptr.reserve((nslices + 1) * sizeof(long long)));
    R_CUDA(steps.reserve((nslices + 1) * sizeof(long long)));
    R_CUDA(slen.reserve(nslices * sizeof(int)));
    R_CUDA(ctx->flag.reserve(64));
    R_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    const long long* rowptr = ctx->rowptr.as<long long>();
    const long long* radj_ptr = ctx->radj_ptr.as<long long>();
    const unsigned* radj = ctx->radj.as<unsigned>();
#define BY_NLOC(KERN, ...)                                   \
    do {                                                     \
        if (nl == 4) KERN<4> __VA_ARGS__;                    \
        else if (nl == 10) KERN<10> __VA_ARGS__;             \
        else KERN<20> __VA_ARGS__;                           \
    } while (0)
    BY_NLOC(k_row_key, <<<grid_for(nrows), 256, 0, st>>>(nrows, rowptr, radj_ptr, radj, key.as<unsigned long long>(), rid.as<unsigned>()));
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key.as<unsigned long long>(), key2.as<unsigned long long>(), rid.as<unsigned>(),
                                    ctx->rp_order.as<unsigned>(), nrows, 0, 62, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, steps.as<long long>(), ctx->rp_sptr.as<long long>(), nslices + 1, st);
    R_CUDA(cubtmp.reserve(std::max(tb, tb2)));
    R_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, key.as<unsigned long long>(), key2.as<unsigned long long>(), rid.as<unsigned>(),
                                           ctx->rp_order.as<unsigned>(), nrows, 0, 62, st));
    R_CUDA(cudaMemsetAsync(steps.p, 0, (nslices + 1) * sizeof(long long), st));
    BY_NLOC(k_slice_info, <<<grid_for(nslices * 32), 256, 0, st>>>(nrows, nslices, ctx->rp_order.as<unsigned>(), rowptr, radj_ptr, radj,
                                                                 ctx->rp_cnt.as<unsigned short>(), steps.as<long long>(), slen.as<int>(),
                                                                 ctx->flag.as<int>()));
    R_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp.p, tb2, steps.as<long long>(), ctx->rp_sptr.as<long long>(), nslices + 1, st));
    long long total = 0;
    int overflow = 0;
    std::vector<int> hlen(nslices);
    R_CUDA(cudaMemcpyAsync(&total, ctx->rp_sptr.as<long long>() + nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(&overflow, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(hlen.data(), slen.p, nslices * sizeof(int), cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaStreamSynchronize(st));
    ctx->launches += 4;
    if (overflow) { cleanup(); return 0; }
    const int nw = (nl + 3) / 4 + 1;
    const size_t ell_bytes = (size_t)std::max<long long>(1, total) * nw * 32 * sizeof(unsigned);
    R_CUDA(ctx->rp_ell.reserve(ell_bytes));
    R_CUDA(cudaMemsetAsync(ctx->rp_ell.p, 0, ell_bytes, st));
    BY_NLOC(k_ell_fill, <<<grid_for(nslices * 32), 256, 0, st>>>(nrows, nslices, ctx->rp_order.as<unsigned>(), radj_ptr, radj,
                                                               ctx->pos.as<unsigned char>(), ctx->rp_cnt.as<unsigned short>(),
                                                               ctx->rp_sptr.as<long long>(), ctx->rp_ell.as<unsigned>()));
    R_CUDA(cudaGetLastError());
    R_CUDA(cudaStreamSynchronize(st));
    ctx->launches++;
#undef BY_NLOC
#undef R_CUDA
    cleanup();
    // launch ranges: consecutive slices of one length bucket (the sort key makes the bucket non-decreasing)
    long long s0 = 0;
    while (s0 < nslices) {
        const int b = len_bucket(hlen[s0]);
        long long s1 = s0 + 1;
        while (s1 < nslices && len_bucket(hlen[s1]) == b) ++s1;
        ctx->rp_buckets.push_back({s0, s1, bucket_L(b)});
        s0 = s1;
    }
    ctx->rp_nloc = nl;
    ctx->rp_steps = total;
    ctx->has_rows_plan = true;
    return 0;
}

// 1 = launched, 0 = combination not covered (caller uses the lane-group gather), < 0 error
int launch_rows(afb_ctx* ctx, int nga, int ngf, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs,
                int accumulate, double drop_val, int* status) {
    if (!ctx->has_rows_plan || ctx->rp_nloc != ctx->nrow_loc) return 0;
    if (getenv("AFB_DISABLE_ROWS_KERNEL")) return 0;
    switch (ctx->rp_nloc) {
        case 4: return launch_rows_n<4>(ctx, nga, ngf, TA, TF, gbuf, val, rhs, accumulate, drop_val, status);
        case 10: return launch_rows_n<10>(ctx, nga, ngf, TA, TF, gbuf, val, rhs, accumulate, drop_val, status);
        case 20: return launch_rows_n<20>(ctx, nga, ngf, TA, TF, gbuf, val, rhs, accumulate, drop_val, status);
    }
    return 0;
}

}  // namespace afb
