// Cluster-tiled, row-stationary gather of the fused tensor-representation path (k_rows_cl) and the plan it streams.
//
// Reference semantics reproduced: the scatter of AssemblerT::Assemble (inmost_interface/assembler.inl:397-481)
//     matrix[r][c] += A_e(i,j) if |A_e(i,j)| > drop_val;  rhs[r] += F_e(i);  non-finite local value -> status -1
// for element matrices A_e(i,j) = sum_c T[c][i][j] g_e[c] (afb_tensor.cu explains the factorisation of fem3Dtet).
// The reference's parallel driver lets every MPI rank recompute the cells around its owned rows
// (assembler.inl:162-183); the same idea is applied here per CTA.
//
// Shape of the kernel, each point answering an ncu finding (profiles/r01*_*.md):
//   * CTA = cluster.  Elements are renumbered along a Morton curve of their centroids; a row belongs to the cluster
//     (chunk of CH consecutive elements) of its first adjacent element.  The CTA first copies the per-element
//     coefficients g_e of every element its rows touch into shared memory (coalesced 16-byte loads), then its warps work
//     through the cluster's row slices.  [lane-group kernel: g_e fetched 10x per element through L1 at 4 LSU wavefronts
//     per visit; thread-per-row kernel with global g_e: L2 hit rate 20 %, 1.2 kB/tet of DRAM traffic.]
//   * thread = row.  Visits (element, local row i) of a row are ordered by i and the 32 rows of a slice (same cluster,
//     same length bucket, similar class signature) step through the classes together, padded to the slice maximum: i is
//     warp-uniform and compile-time inside the unrolled class loop, so T[.][i][j] is an immediate-offset constant-bank /
//     uniform-register operand of the DFMAs (the table travels as a __grid_constant__ kernel parameter): no table
//     loads.  [lane-group kernel: 60 % of the shared-memory wavefronts were table reads.]
//   * accumulators acc[slot][lane] in shared memory: a lane only ever touches its own bank pair, so the NLOC
//     read-modify-writes of a visit are conflict-free whatever the slots are.
//   * the plan is streamed coalesced: per visit-step and lane NW 32-bit words = NLOC slot bytes + 16-bit local element.
//   * finished rows: 16-slot tiles are transposed in place (XOR swizzle) and stored with lanes <-> slots.
// Deterministic: the summation order of a row is fixed by the plan (class, then ascending element); no atomics on data
// (one shared-memory counter hands out slices to warps; it does not influence any sum).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

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

#define GRID_STRIDE(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ---------------------------------------------------------------------------------------------------------------------
// plan kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct BBox { double lo[3], inv[3]; };

__global__ void k_morton(long long ntet, BBox bb, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                         const int32_t* __restrict__ v0, const int32_t* __restrict__ v1, const int32_t* __restrict__ v2,
                         const int32_t* __restrict__ v3, unsigned* code, unsigned* eid) {
    GRID_STRIDE(e, ntet) {
        const int n[4] = {v0[e], v1[e], v2[e], v3[e]};
        double c[3] = {0, 0, 0};
        for (int k = 0; k < 4; ++k) { c[0] += x[n[k]]; c[1] += y[n[k]]; c[2] += z[n[k]]; }
        unsigned q[3];
        for (int d = 0; d < 3; ++d) {
            double t = (0.25 * c[d] - bb.lo[d]) * bb.inv[d] * 1024.0;
            t = fmin(fmax(t, 0.0), 1023.0);
            q[d] = (unsigned)t;
        }
        code[e] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
        eid[e] = (unsigned)e;
    }
}

__global__ void k_invert(long long n, const unsigned* __restrict__ new2old, unsigned* old2new) {
    GRID_STRIDE(i, n) old2new[new2old[i]] = (unsigned)i;
}

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

// sort key of a row: (cluster | length bucket | degree | clipped class counts).  cluster = chunk of the first adjacent
// element in Morton order; rows without elements go to cluster 0.
template <int NLOC>
__global__ void k_row_key(long long nrows, int chunk, const long long* __restrict__ rowptr, const long long* __restrict__ radj_ptr,
                          const unsigned* __restrict__ radj, const unsigned* __restrict__ old2new, unsigned long long* key,
                          unsigned* rowid, unsigned* rcl) {
    constexpr int BITS = (30 / NLOC) > 6 ? 6 : ((30 / NLOC) < 1 ? 1 : (30 / NLOC));
    GRID_STRIDE(r, nrows) {
        const int len = (int)(rowptr[r + 1] - rowptr[r]);
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        int cnt[NLOC];
        class_counts<NLOC>(radj, a0, a1, cnt);
        unsigned mn = 0xffffffffu;
        for (long long a = a0; a < a1; ++a) mn = min(mn, old2new[radj[a] / (unsigned)NLOC]);
        const unsigned cl = a1 > a0 ? mn / (unsigned)chunk : 0u;
        unsigned long long pack = 0;
#pragma unroll
        for (int c = 0; c < NLOC; ++c)
            if (c * BITS < 30) pack = (pack << BITS) | (unsigned long long)min(cnt[c], (1 << BITS) - 1);
        const unsigned long long deg = (unsigned long long)min((long long)63, a1 - a0);
        key[r] = ((unsigned long long)cl << 40) | ((unsigned long long)len_bucket(len) << 36) | (deg << 30) | (pack & 0x3fffffffULL);
        rowid[r] = (unsigned)r;
        rcl[r] = cl;
    }
}

// position p (in sorted order) starts a group when its (cluster, bucket) differs from p-1; gs[p] = p at starts, else 0
__global__ void k_group_start(long long n, const unsigned long long* __restrict__ key, long long* gs) {
    GRID_STRIDE(p, n) gs[p] = (p == 0 || (key[p] >> 36) != (key[p - 1] >> 36)) ? p : 0;
}
// slice starts: every 32nd row of a group
__global__ void k_slice_start(long long n, const long long* __restrict__ gstart, int* flag) {
    GRID_STRIDE(p, n) flag[p] = ((p - gstart[p]) & 31) == 0 ? 1 : 0;
}
struct MaxLL { __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; } };

// slice tables: rows of slice s (0xFFFFFFFF = empty lane), cluster of slice s
__global__ void k_slice_rows(long long n, const int* __restrict__ sof /* inclusive scan of flags */, const unsigned* __restrict__ order,
                             const unsigned long long* __restrict__ key, const long long* __restrict__ gstart, unsigned* srow, unsigned* scl) {
    GRID_STRIDE(p, n) {
        const long long s = sof[p] - 1;
        const int lane = (int)((p - gstart[p]) & 31);
        srow[s * 32 + lane] = order[p];
        if (lane == 0) scl[s] = (unsigned)(key[p] >> 40);
    }
}

// ptr[c] = first index with arr[idx] >= c  (arr non-decreasing), c = 0..nc
__global__ void k_lower_bound_u32(long long n, const unsigned* __restrict__ arr, long long nc, int* ptr) {
    GRID_STRIDE(c, nc + 1) {
        long long lo = 0, hi = n;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (arr[mid] < (unsigned)c) lo = mid + 1; else hi = mid; }
        ptr[c] = (int)lo;
    }
}
__global__ void k_lower_bound_hi32(long long n, const unsigned long long* __restrict__ arr, long long nc, int* ptr) {
    GRID_STRIDE(c, nc + 1) {
        long long lo = 0, hi = n;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((arr[mid] >> 32) < (unsigned long long)c) lo = mid + 1; else hi = mid; }
        ptr[c] = (int)lo;
    }
}

// (cluster of the row, Morton id of the element) of every visit
template <int NLOC>
__global__ void k_visit_pairs(long long nrows, const long long* __restrict__ radj_ptr, const unsigned* __restrict__ radj,
                              const unsigned* __restrict__ old2new, const unsigned* __restrict__ rcl, unsigned long long* pair) {
    GRID_STRIDE(r, nrows) {
        const unsigned long long hi = (unsigned long long)rcl[r] << 32;
        for (long long a = radj_ptr[r]; a < radj_ptr[r + 1]; ++a) pair[a] = hi | old2new[radj[a] / (unsigned)NLOC];
    }
}
__global__ void k_low32(long long n, const unsigned long long* __restrict__ src, unsigned* dst) {
    GRID_STRIDE(i, n) dst[i] = (unsigned)src[i];
}
__global__ void k_max_diff(long long n, const int* __restrict__ ptr, int* out) {
    int m = 0;
    GRID_STRIDE(c, n) m = max(m, ptr[c + 1] - ptr[c]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);  // setup-time integer max
}

// one warp per slice: class counts padded to the slice maximum, number of visit-steps, longest row
template <int NLOC>
__global__ void k_slice_info(long long nslices, const unsigned* __restrict__ srow, const long long* __restrict__ rowptr,
                             const long long* __restrict__ radj_ptr, const unsigned* __restrict__ radj, unsigned short* cntS,
                             long long* steps, long long* sp0, unsigned short* slen, unsigned short* smax,
                             int* flags /* [0] overflow, [2] max row length */) {
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = wglob; s < nslices; s += nw) {
        const unsigned r = srow[s * 32 + lane];
        int cnt[NLOC], len = 0;
        long long p0 = 0;
        if (r != 0xffffffffu) {
            p0 = rowptr[r];
            len = (int)(rowptr[r + 1] - p0);
            class_counts<NLOC>(radj, radj_ptr[r], radj_ptr[r + 1], cnt);
        } else {
#pragma unroll
            for (int c = 0; c < NLOC; ++c) cnt[c] = 0;
        }
        long long tot = 0;
#pragma unroll
        for (int c = 0; c < NLOC; ++c) {
            const int m = __reduce_max_sync(0xffffffffu, cnt[c]);
            if (m > 65535 && lane == 0) flags[0] = 1;
            if (lane == 0) cntS[s * NLOC + c] = (unsigned short)m;
            tot += m;
        }
        const int ml = __reduce_max_sync(0xffffffffu, len);
        sp0[s * 32 + lane] = p0;
        slen[s * 32 + lane] = (unsigned short)len;
        if (lane == 0) { steps[s] = tot; smax[s] = (unsigned short)ml; atomicMax(flags + 2, ml); }
    }
}

// fills the ELL stream: word w of visit-step t of lane l at ell[(t*NW + w)*32 + l]; bytes 0..NC-1 = slots of the NC local
// columns (NLOC = local rows = visit classes), the last two bytes = local element index + 1 inside the cluster (0 = padding visit; buffer pre-set to 0)
template <int NLOC, int NC>
__global__ void k_ell_fill(long long nslices, const unsigned* __restrict__ srow, const unsigned* __restrict__ scl,
                           const long long* __restrict__ radj_ptr, const unsigned* __restrict__ radj, const unsigned char* __restrict__ pos,
                           const unsigned short* __restrict__ cntS, const long long* __restrict__ sptr, const unsigned* __restrict__ old2new,
                           const int* __restrict__ eptr, const unsigned* __restrict__ elist, unsigned* ell) {
    constexpr int NW = (NC + 2 + 3) / 4;
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = wglob; s < nslices; s += nw) {
        const unsigned r = srow[s * 32 + lane];
        if (r == 0xffffffffu) continue;
        const unsigned cl = scl[s];
        const int e0 = eptr[cl], e1 = eptr[cl + 1];
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        long long base = sptr[s];
        for (int c = 0; c < NLOC; ++c) {
            int k = 0;
            for (long long a = a0; a < a1; ++a) {
                const unsigned t = radj[a];
                if ((int)(t % (unsigned)NLOC) != c) continue;
                const unsigned en = old2new[t / (unsigned)NLOC];
                int lo = e0, hi = e1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (elist[mid] < en) lo = mid + 1; else hi = mid; }
                const unsigned eloc = (unsigned)(lo - e0) + 1u;
                unsigned* dst = ell + (size_t)(base + k) * NW * 32 + lane;
                const unsigned char* pa = pos + (size_t)a * NC;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    unsigned word = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (4 * w + b < NC) word |= (unsigned)pa[4 * w + b] << (8 * b);
                    if (w == NW - 1) word |= eloc << 16;
                    dst[w * 32] = word;
                }
                ++k;
            }
            base += cntS[s * NLOC + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the gather
// ---------------------------------------------------------------------------------------------------------------------
template <int NLOC, int NC, int NGA, int NGF>
struct alignas(16) RowTab {
    static constexpr int NGAP = NGA + (NGA & 1);      // even pitch: every (i,j) group is 16-byte aligned
    double A[NGA > 0 ? NLOC * NC * NGAP : 2];         // [(i*NC + j)*NGAP + c], i = local row (test), j = local column (trial)
    double F[NGF > 0 ? NLOC * NGF + 2 : 2];           // [i*NGF + c]
};

struct RowsArgs {
    int L16b, L16s;     // slots per row image in a big / small warp region (multiples of 16)
    int small_len;      // a slice is "small" when its longest row has <= small_len entries
    int nbig;           // warps 0..nbig-1 own big regions and serve the long slices first
    int gcap;           // elements of coefficient staging per CTA
    int zero;           // always 0; k & zero keeps the table loads inside the visit loop (see k_rows_cl)
    const int* clist;   // NULL: CTA b works on cluster b; else on cluster clist[b] (priority / remaining clusters of a phased assembly)
    const int* cs;      // [ncl+1] slice range of a cluster
    const int* eptr;    // [ncl+1] element-list range of a cluster
    const unsigned* elist;   // Morton ids of the elements a cluster touches
    const unsigned* srow;    // [nslices*32] rows of a slice (0xFFFFFFFF = empty lane)
    const long long* sp0;    // [nslices*32] first CSR entry of the row
    const int* tix;          // NULL, or [nslices*32]: 0 = the row's entries are contiguous from sp0, else 1 + first entry of its offset table
    const unsigned short* rtab;   // offsets (relative to sp0) of the entries of non-contiguous rows (blocks of segmented numberings, afb_blocks.cu)
    const int* rdst;         // NULL (rhs[row]), or [nslices*32]: index into rhs that receives the row's load entry
    const unsigned short* slen;   // [nslices*32] row length
    const unsigned short* smax;   // [nslices] longest row of the slice
    const unsigned short* cnt;    // [nslices*NLOC]
    const long long* sptr;
    const unsigned* ell;
    const double* gbuf;      // Morton order
    double* val;
    double* rhs;
    int accumulate;
    double drop;         // drop_val: contributions with |v| <= drop are not added (assembler.inl:416)
    int* status;
    unsigned long long* diag;   // AFB_DIAG_TIMELINE builds only
};

constexpr int RW_RING = 4;  // visit-steps of plan words in flight per warp (cp.async ring)
struct alignas(16) RowDst { double* ptr; int len; int pad; };  // first CSR value and length of a row
__host__ __device__ inline size_t rows_warp_bytes(int L16, int nw) { return (size_t)L16 * 256 + 32 * 16 + (size_t)RW_RING * nw * 128; }

__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// slot of an element record inside a coefficient plane: the low three bits (= the 16-byte bank group) are hashed with
// the next three, so that the regular element strides of structured meshes (6 tets per hex, ...) still spread a warp's
// 32 records over all eight bank groups
__device__ __forceinline__ unsigned rec_slot(unsigned el1) { return el1 ^ (((el1 >> 3) ^ (el1 >> 6) ^ (el1 >> 9)) & 7u); }

template <int NLOC>
struct SliceMeta {
    unsigned r;
    long long p0;
    int len;
    int tix, rdst;
    long long st, en;
    int cn[NLOC];
};

// TAB: rows may be non-contiguous blocks written through offset tables, and the load entry may go to a mapped row (blocks of
// segmented numberings, afb_blocks.cu).  A template switch: the extra row metadata and the table branch in the write-out cost
// the plain kernel 30 % when they are merely present (138 instead of 131 registers, the 8 row stores of a batch serialised).
template <int NLOC, int NC, int NGA, int NGF, bool TAB>
__global__ void __launch_bounds__(384, 1) k_rows_cl(const __grid_constant__ RowTab<NLOC, NC, NGA, NGF> T, const RowsArgs p) {
    constexpr int NW = (NC + 2 + 3) / 4;
    constexpr int NG = NGA + NGF, NGP = (NG + 1) & ~1, PARTS = NGP / 2;
    constexpr int NGAP = RowTab<NLOC, NC, NGA, NGF>::NGAP;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ int s_q[2];  // slices handed out: [0] long ones, [1] short ones
    __shared__ int s_sb;    // first long slice of the cluster
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef AFB_DIAG_TIMELINE   // diagnostic: [0] sum over warps of (end of the warp's last slice - CTA start), [1] sum of (CTA end - CTA start),
                           // [2] sum of (staging barrier - CTA start), all per warp; p.diag = 3 counters
    const long long tl0 = clock64();
#endif
    const int c = p.clist ? p.clist[blockIdx.x] : (int)blockIdx.x;
    const int e0 = p.eptr[c], ne = p.eptr[c + 1] - e0;
    const int sl0 = p.cs[c], sl1 = p.cs[c + 1];
    const unsigned g_a = (unsigned)__cvta_generic_to_shared(smraw);
    // ---- stage the coefficients of the cluster's elements as PARTS planes of 16-byte pieces (SoA): record 0 stays zero
    //      (padding visits read it), element el of the cluster's list lands in slot rec_slot(el+1) of every plane
    const unsigned plane = (((unsigned)p.gcap + 8u) & ~7u) * 16u;  // rec_slot permutes inside blocks of 8 slots
    {
        const char* gsrc = reinterpret_cast<const char*>(p.gbuf);
        if (threadIdx.x < PARTS) {
            double2 zz; zz.x = 0.0; zz.y = 0.0;
            *reinterpret_cast<double2*>(smraw + (size_t)threadIdx.x * plane) = zz;
        }
        // one thread per element: its id, then PARTS asynchronous 16-byte copies (fire and forget; a batch of ids is loaded
        // first so that their latencies overlap).  Measured alternatives: lane-per-piece cp.async and LDG.128+STS.128 are
        // 5-8 % slower end to end -- the phase is latency-bound (one CTA per SM), not LSU-bound.
        constexpr int U = 4;
#ifdef AFB_DIAG_SKIP_STAGING   // timing diagnostic only (wrong results)
        for (int base = 0; base < (p.zero ? ne : 0); base += U * (int)blockDim.x) {
#else
        for (int base = 0; base < ne; base += U * (int)blockDim.x) {
#endif
            unsigned id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int el = base + u * (int)blockDim.x + (int)threadIdx.x;
                id[u] = el < ne ? __ldg(p.elist + e0 + el) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned el1 = (unsigned)(base + u * (int)blockDim.x + (int)threadIdx.x) + 1u;
                if ((int)el1 <= ne) {
                    const unsigned dst = g_a + rec_slot(el1) * 16;
#pragma unroll
                    for (int part = 0; part < PARTS; ++part) cp_async16(dst + part * plane, gsrc + ((size_t)id[u] * PARTS + part) * 16);
                }
            }
        }
        cp_async_commit();
    }
    if (warp == 0) {
        int nsmall = 0;
        for (int b = sl0; b < sl1; b += 32) {
            const int s = b + lane;
            const bool sm = s < sl1 && (int)__ldg(p.smax + s) <= p.small_len;
            nsmall += __popc(__ballot_sync(0xffffffffu, sm));
        }
        if (lane == 0) { s_sb = sl0 + nsmall; s_q[0] = 0; s_q[1] = 0; }
    }
    cp_async_wait<0>();
    __syncthreads();
#ifdef AFB_DIAG_TIMELINE
    const long long tl1 = clock64();
#endif
    const int sb = s_sb;
    const bool isbig = warp < p.nbig;
    const int L16 = isbig ? p.L16b : p.L16s;
    const size_t goff = (size_t)plane * PARTS;
    unsigned char* wraw = smraw + goff +
                          (isbig ? (size_t)warp * rows_warp_bytes(p.L16b, NW)
                                 : (size_t)p.nbig * rows_warp_bytes(p.L16b, NW) + (size_t)(warp - p.nbig) * rows_warp_bytes(p.L16s, NW));
    const unsigned tile_a = (unsigned)__cvta_generic_to_shared(wraw);
    const unsigned acc_a = tile_a + lane * 8;
    RowDst* sdst = reinterpret_cast<RowDst*>(wraw + (size_t)L16 * 256);  // [32] destination of every row of the slice
    const unsigned ring_a = tile_a + L16 * 256 + 32 * 16 + lane * 4;
    const double drop = p.drop;
    bool bad = false;

    auto next_slice = [&]() -> int {
        int s = -1;
        if (lane == 0) {
            if (isbig) {
                const int t = atomicAdd(&s_q[0], 1);
                s = sl1 - 1 - t;
                if (s < sb) s = -1;
            }
            if (s < 0) {
                const int t = atomicAdd(&s_q[1], 1);
                s = sb - 1 - t;
                if (s < sl0) s = -1;
            }
        }
        return __reduce_max_sync(0xffffffffu, s);  // lane 0 holds the value (>= -1), the others -1; warp-uniform for the compiler
    };
    auto load_meta = [&](int s, SliceMeta<NLOC>& m) {
        m.r = __ldg(p.srow + (size_t)s * 32 + lane);
        m.p0 = __ldg(p.sp0 + (size_t)s * 32 + lane);
        m.len = (int)__ldg(p.slen + (size_t)s * 32 + lane);
        if (TAB) {
            m.tix = p.tix ? __ldg(p.tix + (size_t)s * 32 + lane) : 0;
            m.rdst = p.rdst ? __ldg(p.rdst + (size_t)s * 32 + lane) : (int)m.r;
        }
        m.st = __ldg(p.sptr + s);
        m.en = __ldg(p.sptr + s + 1);
#pragma unroll
        for (int i = 0; i < NLOC; ++i) m.cn[i] = (int)__ldg(p.cnt + (size_t)s * NLOC + i);
    };

    // plan words: cp.async ring, RW_RING-1 visit-steps ahead (one group per step, possibly empty)
    auto ring_prologue = [&](long long st0, long long en0) {
        const int nst0 = (int)(en0 - st0);
        const unsigned* pf0 = p.ell + (size_t)st0 * NW * 32 + lane;
#pragma unroll
        for (int d = 0; d < RW_RING - 1; ++d) {
            if (d < nst0) {
#pragma unroll
                for (int q = 0; q < NW; ++q) cp_async4(ring_a + (d * NW + q) * 128, pf0 + q * 32);
            }
            cp_async_commit();
            pf0 += NW * 32;
        }
    };

    SliceMeta<NLOC> cur, nxt;
    int s = next_slice();
    if (s >= 0) { load_meta(s, cur); ring_prologue(cur.st, cur.en); }
    while (s >= 0) {
        const int sn = next_slice();
        if (sn >= 0) load_meta(sn, nxt);  // in flight while this slice is processed

        const int Lw = __reduce_max_sync(0xffffffffu, cur.len);
        if (NGA > 0) {
            const int Lz = (Lw + 15) & ~15;  // whole tiles: the write-out reads (and NaN-checks) 16 slots at a time
#pragma unroll 8
            for (int sl = 0; sl < Lz; ++sl) sts64(acc_a + sl * 256, 0.0);
        }

        const int nst = (int)(cur.en - cur.st);
        // the first RW_RING-1 visit-steps are already in flight (ring_prologue: issued before the write-out of the previous
        // slice; measured gain 0.6 %)
        const unsigned* pf = p.ell + ((size_t)cur.st + (RW_RING - 1)) * NW * 32 + lane;  // next visit-step to prefetch
        int t = 0;
        double fsum = 0.0;
#pragma unroll
        for (int i = 0; i < NLOC; ++i) {
#ifdef AFB_DIAG_SKIP_VISITS   // timing diagnostic only (wrong results)
            const int n = p.zero ? __reduce_max_sync(0xffffffffu, cur.cn[i]) : 0;
#else
            const int n = __reduce_max_sync(0xffffffffu, cur.cn[i]);  // warp-uniform trip count in a uniform register
#endif
            for (int k = 0; k < n; ++k) {
                // z2 is 0 at run time but loop-variant for the compiler: the table entries are then fetched by uniform
                // constant loads (LDCU) next to their DFMA instead of being hoisted out of the loop, where 600 values
                // overflow the uniform register file and end up in vector registers + R2UR moves
                const int z2 = (k & p.zero) * 2;
                if (t + (RW_RING - 1) < nst) {
                    const unsigned ra = ring_a + ((t + RW_RING - 1) & (RW_RING - 1)) * (NW * 128);
#pragma unroll
                    for (int q = 0; q < NW; ++q) cp_async4(ra + q * 128, pf + q * 32);
                }
                cp_async_commit();
                pf += NW * 32;
                cp_async_wait<RW_RING - 1>();
                unsigned w[NW];
                {
                    const unsigned ra = ring_a + (t & (RW_RING - 1)) * (NW * 128);
#pragma unroll
                    for (int q = 0; q < NW; ++q) w[q] = lds32(ra + q * 128);
                }
                const unsigned el1 = w[NW - 1] >> 16;  // 0 = padding visit -> the zero record
                double g[NGP];
                {
                    const unsigned rec = g_a + rec_slot(el1) * 16;
#pragma unroll
                    for (int q = 0; q < PARTS; ++q) {
                        const double2 d = lds128(rec + q * plane);
                        g[2 * q] = d.x; g[2 * q + 1] = d.y;
                    }
                }
                if (NGF > 0) {  // compile-time condition: a run-time one here makes ptxas fall back to per-thread LDC table loads
                    double f = 0.0;
#pragma unroll
                    for (int q = 0; q < NGF; ++q) f = fma(T.F[i * NGF + q + z2], g[NGA + q], f);
                    fsum += f;  // padding visits add 0; NaN/Inf propagate into fsum and are caught below
                }
                if (NGA > 0) {
                    // columns in groups of GJ: load the old sums, GJ*NGA DFMAs with uniform-register table operands, store.
                    // The volatile shared-memory accesses fence the groups, which keeps the number of live table values small
                    // enough for ptxas to use uniform constant loads (LDCU) instead of per-thread LDC (ADU pipe, 8x slower).
                    constexpr int NGA1 = NGA > 0 ? NGA : 1;
                    constexpr int GJ = NGA1 >= 30 ? 1 : (30 / NGA1 > NC ? NC : 30 / NGA1);
#pragma unroll
                    for (int j0 = 0; j0 < NC; j0 += GJ) {
                        double v[GJ], o[GJ];
                        unsigned sa[GJ];
#pragma unroll
                        for (int jj = 0; jj < GJ; ++jj) {
                            const int j = j0 + jj;
                            if (j < NC) {
                                // slot byte -> byte offset slot*256 in one PRMT
                                sa[jj] = acc_a + __byte_perm(w[j >> 2], 0u, 0x4404u | ((unsigned)(j & 3) << 4));
                                o[jj] = lds64(sa[jj]);  // the slots of one visit are distinct (checked by the plan)
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < GJ; ++jj) {
                            const int j = j0 + jj;
                            if (j < NC) {
                                double x = 0.0;
#pragma unroll
                                for (int q = 0; q < NGA; ++q) x = fma(T.A[(i * NC + j) * NGAP + q + z2], g[q], x);
                                v[jj] = x;
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < GJ; ++jj) {
                            const int j = j0 + jj;
                            if (j < NC) {
                                // |A_e(i,j)| > drop_val is the reference's rule (assembler.inl:416); written as !(<=) so that a NaN
                                // IS added and poisons the row sum, which the write-out reports as status -1
                                if (!(fabs(v[jj]) <= drop)) sts64(sa[jj], o[jj] + v[jj]);
                            }
                        }
                    }
                }
                ++t;
            }
        }
        cp_async_wait<0>();
        if (sn >= 0) ring_prologue(nxt.st, nxt.en);   // every lane has read its ring words of this slice: the ring is free

        // ---- write-out: 16-slot tiles transposed in place (XOR swizzle), then one row per step, lanes <-> slots
#ifdef AFB_DIAG_SKIP_WRITEOUT   // timing diagnostic only (wrong results): what the write-out costs (profiles/r01e_*.md)
        if (NGA > 0 && p.zero) {
#else
        if (NGA > 0) {
#endif
            RowDst rd;
            rd.ptr = p.val + cur.p0;
            rd.len = cur.len;
            rd.pad = TAB ? cur.tix : 0;
            sdst[lane] = rd;
            double chk = 0.0;  // becomes NaN iff some finished entry is NaN or +-Inf (x*0 is NaN for those)
            for (int s0 = 0; s0 < Lw; s0 += 16) {
                double x[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q] = lds64(acc_a + (s0 + q) * 256);
#pragma unroll
                for (int q = 0; q < 16; ++q) chk = fma(x[q], 0.0, chk);
                __syncwarp();
                const unsigned ta = tile_a + s0 * 256;
#pragma unroll
                for (int q = 0; q < 16; ++q) sts64(ta + (lane * 16 + (q ^ (lane & 15))) * 8, x[q]);
            }
            bad |= chk != chk;
            __syncwarp();
            const int q = lane & 15, hi = lane >> 4;
            const double* tl = reinterpret_cast<const double*>(wraw);
            for (int s0 = 0; s0 < Lw; s0 += 32) {
                const int sl = s0 + lane;
                const bool tile_on = s0 + hi * 16 < Lw;  // the upper half-warp may face a tile past the end of the row images
                const double* tb = tl + (tile_on ? s0 + hi * 16 : 0) * 32;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 8) {
                    RowDst d[8];
                    double y[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int rl = r0 + u;
                        d[u] = sdst[rl];
                        y[u] = tb[rl * 16 + (q ^ (rl & 15))];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (sl < d[u].len) {
                            // d[u] is warp-uniform (one row per step): the table branch does not diverge
                            const int off = (TAB && d[u].pad) ? (int)__ldg(p.rtab + (d[u].pad - 1 + sl)) : sl;
                            if (p.accumulate) d[u].ptr[off] += y[u]; else d[u].ptr[off] = y[u];
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (NGF > 0) {
            bad |= ((unsigned)__double2hiint(fsum) & 0x7ff00000u) == 0x7ff00000u;
            if (cur.r != 0xffffffffu) {
                const long long rr = TAB ? (long long)cur.rdst : (long long)cur.r;
                if (p.accumulate) p.rhs[rr] += fsum; else p.rhs[rr] = fsum;
            }
        }
        s = sn;
        cur = nxt;
    }
#ifdef AFB_DIAG_TIMELINE
    {
        const long long tl2 = clock64();
        __syncthreads();
        const long long tl3 = clock64();
        if (lane == 0) {
            atomicAdd(p.diag + 0, (unsigned long long)(tl2 - tl0));
            atomicAdd(p.diag + 1, (unsigned long long)(tl3 - tl0));
            atomicAdd(p.diag + 2, (unsigned long long)(tl1 - tl0));
        }
    }
#endif
    if (bad) *p.status = 1;  // benign race: every writer stores the same value
}

// shared-memory layout of a launch: how many warps, how many of them with long-row regions
struct RowsShape { int nwarps, nbig, L16b, L16s, small_len; size_t smem; bool ok; };
RowsShape rows_shape(const afb_ctx* ctx, int ngp) {
    RowsShape r{};
    const int nw = (ctx->rp_ncol + 2 + 3) / 4;
    const size_t budget = 225 * 1024;
    const size_t gbytes = (((size_t)ctx->rp_gcap + 8) & ~(size_t)7) * 16 * (ngp / 2);  // planes of 16-byte pieces, + the zero record
    const int L16 = (ctx->rp_maxlen + 15) & ~15;
    int maxw = 12;
    if (const char* wv = getenv("AFB_ROWS_WARPS")) maxw = std::max(1, std::min(12, atoi(wv)));
    r.L16b = L16; r.ok = false;
    // candidates: uniform regions, or short regions (rows <= 48 / <= 28 entries) + a few long ones
    const int small_lens[3] = {1 << 30, 48, 28};
    double best = 1e300;
    for (int k = 0; k < 3; ++k) {
        const int sl = small_lens[k];
        if (k > 0 && ctx->rp_maxlen <= sl) continue;
        const int L16s = k == 0 ? L16 : ((sl + 15) & ~15);
        const size_t big = rows_warp_bytes(L16, nw), small = rows_warp_bytes(L16s, nw);
        // visit-steps in slices that need a long-row region; the long-region warps also serve short slices afterwards
        const double longs = k == 0 ? 0.0 : (double)ctx->rp_long_steps[k - 1], total = (double)std::max<long long>(1, ctx->rp_steps);
        for (int nbig = (k == 0 ? 0 : 1); nbig <= (k == 0 ? 0 : 6); ++nbig) {
            if (gbytes + big * nbig > budget) break;
            const int nsm = (int)std::min<size_t>(maxw - nbig, (budget - gbytes - big * nbig) / small);
            if (nsm < 0) continue;
            const int tot = nbig + nsm;
            if (tot < 1) continue;
            // time model: the cluster is done when the long slices are done and when all slices are done
            const double t = std::max(nbig ? longs / nbig : 0.0, total / tot);
            if (t < best * 0.999) {
                best = t;
                r.nwarps = tot; r.nbig = nbig; r.L16s = L16s; r.small_len = k == 0 ? (1 << 30) : sl;
                r.smem = gbytes + big * nbig + small * nsm;
                r.ok = tot >= 2;
            }
        }
    }
    return r;
}

template <int NLOC, int NC, int NGA, int NGF>
int launch_rows_t(afb_ctx* ctx, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs, int accumulate,
                  double drop_val, int* status, const long long* p0_override, int phase, const int* tix, const unsigned short* rtab, const int* rdst) {
    constexpr int NGP = (NGA + NGF + 1) & ~1;
    // host staging of the kernel parameter (copied by value at launch); per call, so that contexts on different threads do not share it
    std::unique_ptr<RowTab<NLOC, NC, NGA, NGF>> Tp(new RowTab<NLOC, NC, NGA, NGF>());
    RowTab<NLOC, NC, NGA, NGF>& T = *Tp;
    // TA is [c][i][j], TF is [c][i]
    for (int i = 0; i < NLOC; ++i)
        for (int j = 0; j < NC; ++j)
            for (int c = 0; c < NGA; ++c) T.A[(i * NC + j) * RowTab<NLOC, NC, NGA, NGF>::NGAP + c] = TA[((size_t)c * NLOC + i) * NC + j];
    for (int i = 0; i < NLOC; ++i)
        for (int c = 0; c < NGF; ++c) T.F[i * NGF + c] = TF[(size_t)c * NLOC + i];
    const RowsShape sh = rows_shape(ctx, NGP);
    if (!sh.ok || (NGA > 0 && !val) || (NGF > 0 && !rhs)) return 0;  // does not fit: the caller uses the lane-group gather
    RowsArgs p;
    p.L16b = sh.L16b; p.L16s = sh.L16s; p.small_len = sh.small_len; p.nbig = sh.nbig;
    p.gcap = ctx->rp_gcap;
    p.zero = 0;
    p.cs = ctx->rp_cs.as<int>(); p.eptr = ctx->rp_eptr.as<int>(); p.elist = ctx->rp_elist.as<unsigned>();
    p.srow = ctx->rp_order.as<unsigned>(); p.sp0 = p0_override ? p0_override : ctx->rp_p0.as<long long>(); p.slen = ctx->rp_len.as<unsigned short>();
    p.smax = ctx->rp_smax.as<unsigned short>();
    p.tix = rtab ? tix : nullptr; p.rtab = rtab; p.rdst = rdst;
    p.cnt = ctx->rp_cnt.as<unsigned short>(); p.sptr = ctx->rp_sptr.as<long long>();
    p.ell = ctx->rp_ell.as<unsigned>(); p.gbuf = gbuf;
    p.val = val; p.rhs = rhs; p.accumulate = accumulate; p.status = status;
    p.drop = drop_val;
    auto kern = (p.rtab || p.rdst) ? k_rows_cl<NLOC, NC, NGA, NGF, true> : k_rows_cl<NLOC, NC, NGA, NGF, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_rows_cl)");
    // phased assembly (afb_assemble_phase): 1 = the clusters holding priority rows, 2 = the others, 0 = all
    long long nblocks = ctx->rp_ncl;
    p.clist = nullptr;
    if (phase != 0 && ctx->rp_prio_valid) {
        nblocks = phase == 1 ? ctx->rp_nprio : ctx->rp_ncl - ctx->rp_nprio;
        p.clist = ctx->rp_clist.as<int>() + (phase == 1 ? 0 : ctx->rp_nprio);
    } else if (phase == 2) nblocks = 0;   // everything ran in phase 1
    if (nblocks <= 0) return 1;
#ifdef AFB_DIAG_TIMELINE
    static unsigned long long* ddiag = nullptr;
    if (!ddiag) cudaMalloc(&ddiag, 24);
    cudaMemsetAsync(ddiag, 0, 24, ctx->stream);
    p.diag = ddiag;
#else
    p.diag = nullptr;
#endif
    kern<<<(unsigned)nblocks, sh.nwarps * 32, sh.smem, ctx->stream>>>(T, p);
#ifdef AFB_DIAG_TIMELINE
    {
        unsigned long long h[3];
        cudaMemcpyAsync(h, ddiag, 24, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[afb diag] warps busy until %.1f%% of their CTA's lifetime on average; staging barrier at %.1f%%; mean CTA lifetime %.0f cycles\n",
                100.0 * h[0] / h[1], 100.0 * h[2] / h[1], (double)h[1] / ((double)nblocks * sh.nwarps));
    }
#endif
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_rows_cl launch");
    return 1;
}

template <int NLOC, int NC>
int launch_rows_n(afb_ctx* ctx, int nga, int ngf, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs,
                  int accumulate, double drop_val, int* status, const long long* p0_override, int phase, const int* tix, const unsigned short* rtab,
                  const int* rdst) {
#define RW(A, F) if (nga == A && ngf == F) return launch_rows_t<NLOC, NC, A, F>(ctx, TA, TF, gbuf, val, rhs, accumulate, drop_val, status, p0_override, phase, tix, rtab, rdst);
    RW(6, 1) RW(6, 0) RW(7, 1) RW(7, 0) RW(1, 1) RW(1, 0) RW(0, 1) RW(3, 0) RW(3, 1)
    if constexpr (NLOC <= 10) { RW(9, 1) RW(9, 0) RW(10, 1) RW(10, 0) }
#undef RW
    return 0;
}

}  // namespace

namespace afb {

bool rows_supports(const afb_ctx* ctx, int nga, int ngf) {
    if (!ctx->has_rows_plan || ctx->rp_nloc != ctx->nrow_loc || ctx->rp_ncol != ctx->ncol_loc) return false;
    if (getenv("AFB_DISABLE_ROWS_KERNEL")) return false;
    const bool base = (ngf == 0 || ngf == 1) && (nga == 6 || nga == 7 || nga == 1 || nga == 3 || (nga == 0 && ngf == 1));
    const bool wide = ctx->rp_nloc <= 10 && (ngf == 0 || ngf == 1) && (nga == 9 || nga == 10);
    if (!(base || wide)) return false;
    return rows_shape(ctx, (nga + ngf + 1) & ~1).ok;
}

// Builds the cluster plan from the adjacency lists and the slot table (afb_pattern.cu).  Returns 0; the plan is simply
// absent (has_rows_plan = false -> lane-group gather) when the dof map is not one of the supported shapes.
int build_rows_plan(afb_ctx* ctx) {
    ctx->has_rows_plan = false;
    const int nl = ctx->nrow_loc, nc = ctx->ncol_loc;
    if (getenv("AFB_DISABLE_ROWS_PLAN")) return 0;
    // shapes: square P1/P2/P3 blocks and the rectangular P2 x P1 / P1 x P2 blocks of mixed spaces
    if (!((nl == nc && (nl == 4 || nl == 10 || nl == 20)) || (nl == 10 && nc == 4) || (nl == 4 && nc == 10))) return 0;
    if (ctx->pos_bytes != 1 || ctx->has_signs || ctx->pos_has_dup) return 0;
    const long long nrows = ctx->row_end - ctx->row_begin, ntet = ctx->ntet, nadj = ctx->n_adj;
    if (nrows <= 0 || nadj <= 0 || ctx->nnode <= 0) return 0;
    cudaStream_t st = ctx->stream;
    DevBuf key, key2, rid, cubtmp, steps, code, code2, eid, gstart, sflag, sof, rcl, pair, pair2, scl, red, nsel, sorted_rows;
    auto cleanup = [&]() {
        for (DevBuf* b : {&key, &key2, &rid, &cubtmp, &steps, &code, &code2, &eid, &gstart, &sflag, &sof, &rcl, &pair, &pair2, &scl, &red, &nsel, &sorted_rows})
            b->release();
    };
#define R_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
#define BY_NLOC(KERN, ...)                                   \
    do {                                                     \
        if (nl == 4) KERN<4> __VA_ARGS__;                    \
        else if (nl == 10) KERN<10> __VA_ARGS__;             \
        else KERN<20> __VA_ARGS__;                           \
    } while (0)
    const long long* rowptr = ctx->rowptr.as<long long>();
    const long long* radj_ptr = ctx->radj_ptr.as<long long>();
    const unsigned* radj = ctx->radj.as<unsigned>();

    // ---- temp storage for every cub call below
    size_t tb = 0, tmax = 0;
    R_CUDA(red.reserve(8 * sizeof(double)));
    cub::DeviceReduce::Min(nullptr, tb, ctx->x.as<double>(), red.as<double>(), ctx->nnode, st); tmax = std::max(tmax, tb);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, ntet, 0, 30, st); tmax = std::max(tmax, tb);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, nrows, 0, 64, st); tmax = std::max(tmax, tb);
    cub::DeviceRadixSort::SortKeys(nullptr, tb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, nadj, 0, 64, st); tmax = std::max(tmax, tb);
    cub::DeviceSelect::Unique(nullptr, tb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long*)nullptr, nadj, st); tmax = std::max(tmax, tb);
    cub::DeviceScan::InclusiveScan(nullptr, tb, (long long*)nullptr, (long long*)nullptr, MaxLL(), nrows, st); tmax = std::max(tmax, tb);
    cub::DeviceScan::InclusiveSum(nullptr, tb, (int*)nullptr, (int*)nullptr, nrows, st); tmax = std::max(tmax, tb);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (long long*)nullptr, (long long*)nullptr, nrows + 1, st); tmax = std::max(tmax, tb);
    R_CUDA(cubtmp.reserve(tmax));
    tb = tmax;

    // ---- Morton order of the elements
    double hb[6];
    {
        const double* xs[3] = {ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>()};
        for (int d = 0; d < 3; ++d) {
            size_t t1 = tmax;
            R_CUDA(cub::DeviceReduce::Min(cubtmp.p, t1, xs[d], red.as<double>() + d, ctx->nnode, st));
            t1 = tmax;
            R_CUDA(cub::DeviceReduce::Max(cubtmp.p, t1, xs[d], red.as<double>() + 3 + d, ctx->nnode, st));
        }
        R_CUDA(cudaMemcpyAsync(hb, red.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
        R_CUDA(cudaStreamSynchronize(st));
    }
    BBox bb;
    for (int d = 0; d < 3; ++d) {
        bb.lo[d] = hb[d];
        const double ext = hb[3 + d] - hb[d];
        bb.inv[d] = ext > 0 ? 1.0 / ext : 0.0;
    }
    R_CUDA(code.reserve(ntet * 4)); R_CUDA(code2.reserve(ntet * 4)); R_CUDA(eid.reserve(ntet * 4));
    R_CUDA(ctx->rp_new2old.reserve(ntet * 4)); R_CUDA(ctx->rp_old2new.reserve(ntet * 4));
    k_morton<<<grid_for(ntet), 256, 0, st>>>(ntet, bb, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(), ctx->v[0].as<int32_t>(),
                                            ctx->v[1].as<int32_t>(), ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>(), code.as<unsigned>(), eid.as<unsigned>());
    tb = tmax;
    R_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, code.as<unsigned>(), code2.as<unsigned>(), eid.as<unsigned>(), ctx->rp_new2old.as<unsigned>(), ntet, 0, 30, st));
    k_invert<<<grid_for(ntet), 256, 0, st>>>(ntet, ctx->rp_new2old.as<unsigned>(), ctx->rp_old2new.as<unsigned>());
    ctx->launches += 3;
    R_CUDA(cudaStreamSynchronize(st));
    code.release(); code2.release(); eid.release();
    const unsigned* old2new = ctx->rp_old2new.as<unsigned>();

    R_CUDA(key.reserve(nrows * 8)); R_CUDA(key2.reserve(nrows * 8)); R_CUDA(rid.reserve(nrows * 4)); R_CUDA(sorted_rows.reserve(nrows * 4));
    R_CUDA(rcl.reserve(nrows * 4)); R_CUDA(gstart.reserve(nrows * 8)); R_CUDA(sflag.reserve(nrows * 4)); R_CUDA(sof.reserve(nrows * 4));
    R_CUDA(pair.reserve(nadj * 8)); R_CUDA(pair2.reserve(nadj * 8)); R_CUDA(nsel.reserve(8));
    R_CUDA(ctx->flag.reserve(64));

    const int chunks[3] = {512, 256, 128};
    int env_chunk = 0;
    if (const char* cv = getenv("AFB_ROWS_CHUNK")) env_chunk = atoi(cv);
    const int gcap_limit = 1536;  // elements whose coefficients are staged per CTA (x 64..80 bytes)
    bool ok = false;
    long long ncl = 0, nslices = 0, nu = 0;
    int gcap = 0;
    for (int attempt = 0; attempt < 3 && !ok; ++attempt) {
        const int chunk = env_chunk > 0 ? env_chunk : chunks[attempt];
        ncl = (ntet + chunk - 1) / chunk;
        if (ncl >= (1LL << 24)) continue;
        // rows: key, sort, groups, slices
        BY_NLOC(k_row_key, <<<grid_for(nrows), 256, 0, st>>>(nrows, chunk, rowptr, radj_ptr, radj, old2new, key.as<unsigned long long>(), rid.as<unsigned>(), rcl.as<unsigned>()));
        tb = tmax;
        R_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, key.as<unsigned long long>(), key2.as<unsigned long long>(), rid.as<unsigned>(), sorted_rows.as<unsigned>(), nrows, 0, 64, st));
        k_group_start<<<grid_for(nrows), 256, 0, st>>>(nrows, key2.as<unsigned long long>(), gstart.as<long long>());
        tb = tmax;
        R_CUDA(cub::DeviceScan::InclusiveScan(cubtmp.p, tb, gstart.as<long long>(), gstart.as<long long>(), MaxLL(), nrows, st));
        k_slice_start<<<grid_for(nrows), 256, 0, st>>>(nrows, gstart.as<long long>(), sflag.as<int>());
        tb = tmax;
        R_CUDA(cub::DeviceScan::InclusiveSum(cubtmp.p, tb, sflag.as<int>(), sof.as<int>(), nrows, st));
        int ns = 0;
        R_CUDA(cudaMemcpyAsync(&ns, sof.as<int>() + (nrows - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        R_CUDA(cudaStreamSynchronize(st));
        nslices = ns;
        R_CUDA(ctx->rp_order.reserve((size_t)nslices * 32 * 4));
        R_CUDA(scl.reserve(nslices * 4));
        R_CUDA(cudaMemsetAsync(ctx->rp_order.p, 0xFF, (size_t)nslices * 32 * 4, st));
        k_slice_rows<<<grid_for(nrows), 256, 0, st>>>(nrows, sof.as<int>(), sorted_rows.as<unsigned>(), key2.as<unsigned long long>(), gstart.as<long long>(),
                                                    ctx->rp_order.as<unsigned>(), scl.as<unsigned>());
        R_CUDA(ctx->rp_cs.reserve((ncl + 1) * 4));
        k_lower_bound_u32<<<grid_for(ncl + 1), 256, 0, st>>>(nslices, scl.as<unsigned>(), ncl, ctx->rp_cs.as<int>());
        // elements a cluster touches
        BY_NLOC(k_visit_pairs, <<<grid_for(nrows), 256, 0, st>>>(nrows, radj_ptr, radj, old2new, rcl.as<unsigned>(), pair.as<unsigned long long>()));
        tb = tmax;
        R_CUDA(cub::DeviceRadixSort::SortKeys(cubtmp.p, tb, pair.as<unsigned long long>(), pair2.as<unsigned long long>(), nadj, 0, 64, st));
        tb = tmax;
        R_CUDA(cub::DeviceSelect::Unique(cubtmp.p, tb, pair2.as<unsigned long long>(), pair.as<unsigned long long>(), nsel.as<long long>(), nadj, st));
        R_CUDA(cudaMemcpyAsync(&nu, nsel.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
        R_CUDA(cudaStreamSynchronize(st));
        if (nu >= 2147483000LL) continue;
        R_CUDA(ctx->rp_eptr.reserve((ncl + 1) * 4));
        R_CUDA(ctx->rp_elist.reserve(nu * 4));
        k_lower_bound_hi32<<<grid_for(ncl + 1), 256, 0, st>>>(nu, pair.as<unsigned long long>(), ncl, ctx->rp_eptr.as<int>());
        k_low32<<<grid_for(nu), 256, 0, st>>>(nu, pair.as<unsigned long long>(), ctx->rp_elist.as<unsigned>());
        R_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
        k_max_diff<<<grid_for(ncl), 256, 0, st>>>(ncl, ctx->rp_eptr.as<int>(), ctx->flag.as<int>() + 1);
        R_CUDA(cudaMemcpyAsync(&gcap, ctx->flag.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        R_CUDA(cudaStreamSynchronize(st));
        ctx->launches += 9;
        if (gcap <= gcap_limit || env_chunk > 0) ok = gcap < 65535;
        if (env_chunk > 0) break;
    }
    if (!ok) { cleanup(); return 0; }
    pair2.release(); key.release(); rid.release(); gstart.release(); sflag.release(); sof.release(); sorted_rows.release(); key2.release();

    // ---- slices: class counts, visit-step offsets, ELL stream
    R_CUDA(ctx->rp_cnt.reserve((size_t)nslices * nl * sizeof(unsigned short)));
    R_CUDA(ctx->rp_sptr.reserve((nslices + 1) * sizeof(long long)));
    R_CUDA(ctx->rp_p0.reserve((size_t)nslices * 32 * sizeof(long long)));
    R_CUDA(ctx->rp_len.reserve((size_t)nslices * 32 * sizeof(unsigned short)));
    R_CUDA(ctx->rp_smax.reserve((size_t)nslices * sizeof(unsigned short)));
    R_CUDA(steps.reserve((nslices + 1) * sizeof(long long)));
    R_CUDA(cudaMemsetAsync(steps.p, 0, (nslices + 1) * sizeof(long long), st));
    R_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    BY_NLOC(k_slice_info, <<<grid_for(nslices * 32), 256, 0, st>>>(nslices, ctx->rp_order.as<unsigned>(), rowptr, radj_ptr, radj,
                                                                 ctx->rp_cnt.as<unsigned short>(), steps.as<long long>(), ctx->rp_p0.as<long long>(),
                                                                 ctx->rp_len.as<unsigned short>(), ctx->rp_smax.as<unsigned short>(), ctx->flag.as<int>()));
    tb = tmax;
    R_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, steps.as<long long>(), ctx->rp_sptr.as<long long>(), nslices + 1, st));
    long long total = 0;
    int hflags[4] = {0, 0, 0, 0};
    R_CUDA(cudaMemcpyAsync(&total, ctx->rp_sptr.as<long long>() + nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(hflags, ctx->flag.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaStreamSynchronize(st));
    if (hflags[0]) { cleanup(); return 0; }
    const int nw = (nc + 2 + 3) / 4;
    const size_t ell_bytes = (size_t)std::max<long long>(1, total) * nw * 32 * sizeof(unsigned);
    R_CUDA(ctx->rp_ell.reserve(ell_bytes));
    R_CUDA(cudaMemsetAsync(ctx->rp_ell.p, 0, ell_bytes, st));
#define ELL_FILL(NRr, NCc) k_ell_fill<NRr, NCc><<<grid_for(nslices * 32), 256, 0, st>>>(nslices, ctx->rp_order.as<unsigned>(), scl.as<unsigned>(), radj_ptr, radj, \
                                                               ctx->pos.as<unsigned char>(), ctx->rp_cnt.as<unsigned short>(),                            \
                                                               ctx->rp_sptr.as<long long>(), old2new, ctx->rp_eptr.as<int>(),                           \
                                                               ctx->rp_elist.as<unsigned>(), ctx->rp_ell.as<unsigned>())
    if (nl == 4 && nc == 4) ELL_FILL(4, 4);
    else if (nl == 10 && nc == 10) ELL_FILL(10, 10);
    else if (nl == 20 && nc == 20) ELL_FILL(20, 20);
    else if (nl == 10 && nc == 4) ELL_FILL(10, 4);
    else ELL_FILL(4, 10);
#undef ELL_FILL
    R_CUDA(cudaGetLastError());
    R_CUDA(cudaStreamSynchronize(st));
    ctx->launches += 2;
#undef BY_NLOC
#undef R_CUDA
    cleanup();
    ctx->rp_nloc = nl;
    ctx->rp_ncol = nc;
    ctx->rp_steps = total;
    ctx->rp_ncl = ncl;
    ctx->rp_gcap = gcap;
    ctx->rp_maxlen = std::max(1, hflags[2]);
    ctx->rp_nslices = nslices;
    {   // visit-steps in slices whose longest row exceeds 48 / 28 entries (launch shape model, rows_shape)
        std::vector<long long> hs(nslices + 1);
        std::vector<unsigned short> hm(nslices);
        if (cudaMemcpy(hs.data(), ctx->rp_sptr.p, (nslices + 1) * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(hm.data(), ctx->rp_smax.p, nslices * sizeof(unsigned short), cudaMemcpyDeviceToHost) != cudaSuccess)
            return cuda_fail(ctx, cudaGetLastError(), "plan statistics");
        ctx->rp_long_steps[0] = ctx->rp_long_steps[1] = 0;
        for (long long s2 = 0; s2 < nslices; ++s2) {
            const long long st2 = hs[s2 + 1] - hs[s2];
            if (hm[s2] > 48) ctx->rp_long_steps[0] += st2;
            if (hm[s2] > 28) ctx->rp_long_steps[1] += st2;
        }
    }
    ctx->has_rows_plan = true;
    if (getenv("AFB_VERBOSE")) {
        const RowsShape sh = rows_shape(ctx, 8);
        fprintf(stderr, "[afb] launch shape for 8 coefficient doubles: %d warps (%d long-row regions of %d slots, short ones %d slots for rows <= %d), %zu B shared\n",
                sh.nwarps, sh.nbig, sh.L16b, sh.L16s, sh.small_len, sh.smem);
    }
    if (getenv("AFB_VERBOSE"))
        fprintf(stderr, "[afb] cluster plan: nloc %d, %lld clusters, %lld slices (%.1f%% lanes filled), %lld visit-steps (%.1f%% real visits), "
                        "%lld staged elements (x%.2f), gcap %d, max row length %d\n",
                nl, ncl, nslices, 100.0 * nrows / (32.0 * nslices), total, 100.0 * nadj / (32.0 * std::max<long long>(1, total)), nu,
                (double)nu / ntet, gcap, ctx->rp_maxlen);
    return 0;
}

// Splits the clusters into those that hold a row >= first_priority_row (they are listed first) and the others: a phased
// assembly runs the first group, lets the caller start the exchange of those rows, and runs the rest meanwhile.
int rows_priority_build(afb_ctx* ctx, long long first_priority_row) {
    ctx->rp_prio_valid = false;
    if (!ctx->has_rows_plan) return 0;
    const long long ncl = ctx->rp_ncl, nsl = ctx->rp_nslices;
    std::vector<int> cs(ncl + 1);
    std::vector<unsigned> srow((size_t)nsl * 32);
    if (cudaMemcpy(cs.data(), ctx->rp_cs.p, (ncl + 1) * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(srow.data(), ctx->rp_order.p, (size_t)nsl * 32 * sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess)
        return cuda_fail(ctx, cudaGetLastError(), "rows_priority_build");
    std::vector<int> first, rest;
    for (long long c = 0; c < ncl; ++c) {
        bool prio = false;
        for (long long t = (long long)cs[c] * 32; t < (long long)cs[c + 1] * 32 && !prio; ++t)
            prio = srow[t] != 0xffffffffu && (long long)srow[t] >= first_priority_row;
        (prio ? first : rest).push_back((int)c);
    }
    ctx->rp_nprio = (long long)first.size();
    first.insert(first.end(), rest.begin(), rest.end());
    if (ctx->rp_clist.reserve(std::max<long long>(1, ncl) * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(ctx->rp_clist.p, first.data(), ncl * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
        return cuda_fail(ctx, cudaGetLastError(), "rows_priority_build");
    ctx->rp_prio_valid = true;
    return 0;
}

// 1 = launched, 0 = combination not covered (caller uses the lane-group gather), < 0 error.  gbuf must be in Morton order.
// p0_override: first CSR entry of every (slice, lane) row when the rows of this plan are sub-blocks of longer rows (afb_blocks.cu).
int launch_rows(afb_ctx* ctx, int nga, int ngf, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs,
                int accumulate, double drop_val, int* status, const long long* p0_override, int phase, const int* tix, const unsigned short* rtab,
                const int* rdst) {
    if (!rows_supports(ctx, nga, ngf)) return 0;
    const int nl = ctx->rp_nloc, nc = ctx->rp_ncol;
#define RWD(NRr, NCc) if (nl == NRr && nc == NCc) return launch_rows_n<NRr, NCc>(ctx, nga, ngf, TA, TF, gbuf, val, rhs, accumulate, drop_val, status, p0_override, phase, tix, rtab, rdst);
    RWD(4, 4) RWD(10, 10) RWD(20, 20) RWD(10, 4) RWD(4, 10)
#undef RWD
    return 0;
}

}  // namespace afb
