// Sparsity pattern + gather plan, built on the device from the elem->dof tables.
//
// Replaces (reference): AssemblerT::AssembleTemplate (inmost_interface/assembler.inl:589-695: per owned row the
// sorted union of the column ids of all adjacent cells) and set_elements_on_matrix_diagonal (:114-136).  Where the
// reference merges one cell at a time into growable INMOST::Sparse::Row objects under row locks, here
//   1. the (element,local row) pairs are stably radix-sorted by row (CUB) -> row adjacency lists `radj`, ascending
//      element index inside a row: this fixes the summation order of the assembly once and for all;
//   2. one warp per row gathers the candidate columns of the adjacent elements (+ the forced diagonal) into shared
//      memory, bitonic-sorts and uniques them: pass 1 counts (-> rowptr by prefix sum), pass 2 writes colind;
//   3. every (element,i) pair of the row lists looks up the slots of the element's columns inside the row (binary
//      search) -> `pos` table in adjacency order (8-bit when all rows have <= 256 entries, else 16-bit).
// The plan (radj, pos) is what makes the value scatter atomic-free and deterministic (afb_gather.cu).
#include <cub/cub.cuh>

#include <algorithm>

#include "afb_internal.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

__global__ void k_adj_keys(long long ntet, int nrow_loc, unsigned nrows, const int32_t* e2r, unsigned* key, unsigned* val) {
    const long long n = ntet * nrow_loc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long e = t / nrow_loc;
        const int i = (int)(t - e * nrow_loc);
        const int c = e2r[(long long)i * ntet + e];
        key[t] = c == 0 ? nrows : (unsigned)(abs(c) - 1);
        val[t] = (unsigned)t;
    }
}

// ptr[r] = first position with key >= r, r = 0..nrows
__global__ void k_lower_bounds(long long n, const unsigned* key, long long nrows, long long* ptr, int* max_deg) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r <= nrows; r += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (key[mid] < (unsigned)r) lo = mid + 1; else hi = mid;
        }
        ptr[r] = lo;
    }
}
__global__ void k_max_deg(long long nrows, const long long* ptr, int* max_deg) {
    int m = 0;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x)
        m = max(m, (int)(ptr[r + 1] - ptr[r]));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_deg, m);  // setup-time integer max: order independent
}

// One warp per row: candidates -> shared, bitonic sort, unique.  FILL = false: counts; true: writes colind.
template <bool FILL>
__global__ void k_row_columns(long long nrows, long long row_begin, long long ntet, int nrow_loc, int ncol_loc,
                              const long long* radj_ptr, const unsigned* radj, const int32_t* e2c, int cap,
                              long long* rowcnt /* [nrows+1], counts at r+1 */, const long long* rowptr, int32_t* colind,
                              const int32_t* diag_col /* NULL: row_begin + r; else per row, -1 = no forced diagonal */) {
    extern __shared__ int sbuf[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int* buf = sbuf + (size_t)wib * cap;
    for (long long r = (long long)blockIdx.x * wpb + wib; r < nrows; r += (long long)gridDim.x * wpb) {
        const long long a0 = radj_ptr[r], a1 = radj_ptr[r + 1];
        const int dcol = diag_col ? diag_col[r] : (int)(row_begin + r);
        const int ncand = (int)(a1 - a0) * ncol_loc + (dcol >= 0 ? 1 : 0);
        int n2 = 32;
        while (n2 < ncand) n2 <<= 1;
        for (int t = lane; t < n2; t += 32) {
            int c = 0x7fffffff;
            if (t < (int)(a1 - a0) * ncol_loc) {
                const unsigned ei = radj[a0 + t / ncol_loc];
                const long long e = ei / nrow_loc;
                const int j = t % ncol_loc;
                c = abs(e2c[(long long)j * ntet + e]) - 1;
            } else if (t < ncand) c = dcol;  // forced diagonal
            buf[t] = c;
        }
        __syncwarp();
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < n2; t += 32) {
                    const int p = t ^ j;
                    if (p > t) {
                        const int a = buf[t], b = buf[p];
                        const bool up = (t & k) == 0;
                        if ((a > b) == up) { buf[t] = b; buf[p] = a; }
                    }
                }
                __syncwarp();
            }
        // unique
        int cnt = 0;
        const long long base = FILL ? rowptr[r] : 0;
        for (int t0 = 0; t0 < ncand; t0 += 32) {
            const int t = t0 + lane;
            const bool head = t < ncand && (t == 0 || buf[t] != buf[t - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, head);
            if (FILL && head) colind[base + cnt + __popc(m & ((1u << lane) - 1))] = buf[t];
            cnt += __popc(m);
        }
        if (!FILL && lane == 0) rowcnt[r + 1] = cnt;
        __syncwarp();
    }
}

// slot table in adjacency order: for the a-th (element, local row) pair of the row lists, the slot of each of the
// element's columns inside that row.  8-bit when every row has <= 256 entries, else 16-bit.
template <typename PosT>
__global__ void k_pos(long long n_adj, long long ntet, int nrow_loc, int ncol_loc, const unsigned* radj, const int32_t* e2r,
                      const int32_t* e2c, const long long* rowptr, const int32_t* colind, PosT* pos, int* missing) {
    for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < n_adj; a += (long long)gridDim.x * blockDim.x) {
        const unsigned t = radj[a];
        const long long e = t / nrow_loc;
        const int i = (int)(t - e * nrow_loc);
        const long long r = abs(e2r[(long long)i * ntet + e]) - 1;
        const long long b = rowptr[r], en = rowptr[r + 1];
        for (int j = 0; j < ncol_loc; ++j) {
            const int c = abs(e2c[(long long)j * ntet + e]) - 1;
            long long lo = b, hi = en;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (colind[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo >= en || colind[lo] != c) *missing = 1;  // the pattern does not contain this structural entry
            pos[a * ncol_loc + j] = (PosT)(lo - b);
        }
        // two local columns on one global column (collapsed dofs): the batched row update of afb_rows.cu must not be used
        for (int j = 1; j < ncol_loc; ++j)
            for (int j2 = 0; j2 < j; ++j2)
                if (pos[a * ncol_loc + j] == pos[a * ncol_loc + j2]) missing[1] = 1;
    }
}

__global__ void k_max_len(long long nrows, const long long* rowptr, int* out) {
    int m = 0;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x)
        m = max(m, (int)(rowptr[r + 1] - rowptr[r]));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

}  // namespace

extern "C" {

static int pattern_impl(afb_ctx* ctx, int64_t* nnz_out, const int64_t* user_rowptr, const int32_t* user_colind, int64_t user_nnz, int mem_space);

int afb_pattern_build(afb_ctx* ctx, int64_t* nnz_out) { return pattern_impl(ctx, nnz_out, nullptr, nullptr, 0, AFB_HOST); }

int afb_pattern_set(afb_ctx* ctx, const int64_t* rowptr, const int32_t* colind, int64_t nnz, int mem_space) {
    if (!rowptr || (nnz > 0 && !colind) || nnz < 0) { afb::set_error(ctx, "afb_pattern_set: bad arguments"); return -7; }
    return pattern_impl(ctx, nullptr, rowptr, colind, nnz, mem_space);
}

static int pattern_impl(afb_ctx* ctx, int64_t* nnz_out, const int64_t* user_rowptr, const int32_t* user_colind, int64_t user_nnz, int mem_space) {
    if (!ctx) return -7;
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified (afb_dofmap_set / afb_dofmap_natural)"); return -6; }
    cudaSetDevice(ctx->device);
    blocks_clear_dst(ctx);
    ctx->dir_rows_valid = false;
    ctx->bf_plan_valid = false;
    const long long ntet = ctx->ntet, nrows = ctx->row_end - ctx->row_begin;
    const int nrl = ctx->nrow_loc, ncl = ctx->ncol_loc;
    const long long nitem = ntet * nrl;
    cudaStream_t st = ctx->stream;
    afb::DevBuf key, key2, val, cubtmp;
    auto cleanup = [&]() { key.release(); key2.release(); val.release(); cubtmp.release(); };
#define P_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
    // 1. adjacency
    P_CUDA(key.reserve(nitem * 4)); P_CUDA(key2.reserve(nitem * 4)); P_CUDA(val.reserve(nitem * 4));
    P_CUDA(ctx->radj.reserve(nitem * 4));
    P_CUDA(ctx->radj_ptr.reserve((nrows + 1) * sizeof(long long)));
    P_CUDA(ctx->rowptr.reserve((nrows + 1) * sizeof(long long)));
    P_CUDA(ctx->flag.reserve(64));
    P_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    k_adj_keys<<<grid_for(nitem), 256, 0, st>>>(ntet, nrl, (unsigned)nrows, ctx->e2r.as<int32_t>(), key.as<unsigned>(), val.as<unsigned>());
    int bits = 1;
    while ((1LL << bits) <= nrows) ++bits;
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key.as<unsigned>(), key2.as<unsigned>(), val.as<unsigned>(), ctx->radj.as<unsigned>(), nitem, 0, bits, st);
    cub::DeviceScan::InclusiveSum(nullptr, tb2, ctx->rowptr.as<long long>(), ctx->rowptr.as<long long>(), nrows + 1, st);
    P_CUDA(cubtmp.reserve(std::max(tb, tb2)));
    P_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, key.as<unsigned>(), key2.as<unsigned>(), val.as<unsigned>(), ctx->radj.as<unsigned>(), nitem, 0, bits, st));
    k_lower_bounds<<<grid_for(nrows + 1), 256, 0, st>>>(nitem, key2.as<unsigned>(), nrows, ctx->radj_ptr.as<long long>(), nullptr);
    k_max_deg<<<grid_for(nrows), 256, 0, st>>>(nrows, ctx->radj_ptr.as<long long>(), ctx->flag.as<int>());
    ctx->launches += 4;
    int max_deg = 0;
    long long n_adj = 0;
    P_CUDA(cudaMemcpyAsync(&max_deg, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaMemcpyAsync(&n_adj, ctx->radj_ptr.as<long long>() + nrows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
    ctx->n_adj = n_adj;
    key.release(); val.release();
    const int32_t* diag = ctx->has_diag ? ctx->diag_col.as<int32_t>() : nullptr;
    long long nnz = 0;
    if (user_rowptr) {
        // caller-supplied (superset) pattern, e.g. the union with columns contributed by other ranks
        const cudaMemcpyKind kin = mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        nnz = user_nnz;
        P_CUDA(ctx->colind.reserve(std::max<long long>(nnz, 1) * sizeof(int32_t)));
        P_CUDA(cudaMemcpyAsync(ctx->rowptr.p, user_rowptr, (nrows + 1) * sizeof(long long), kin, st));
        if (nnz) P_CUDA(cudaMemcpyAsync(ctx->colind.p, user_colind, nnz * sizeof(int32_t), kin, st));
    } else {
    // 2. rows: count, scan, fill
    int cap = 32;
    while (cap < max_deg * ncl + 1) cap <<= 1;
    if ((size_t)cap * 4 > 200 * 1024) { cleanup(); set_error(ctx, "afb_pattern_build: a row has too many candidate columns"); return -3; }
    int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (96 * 1024) / ((size_t)cap * 4)));
    const size_t smem = (size_t)wpb * cap * 4;
    P_CUDA(cudaFuncSetAttribute(k_row_columns<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P_CUDA(cudaFuncSetAttribute(k_row_columns<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned gridr = (unsigned)std::max<long long>(1, std::min<long long>((nrows + wpb - 1) / wpb, 148LL * 64));
    P_CUDA(cudaMemsetAsync(ctx->rowptr.p, 0, sizeof(long long), st));
    k_row_columns<false><<<gridr, wpb * 32, smem, st>>>(nrows, ctx->row_begin, ntet, nrl, ncl, ctx->radj_ptr.as<long long>(), ctx->radj.as<unsigned>(),
                                                      ctx->e2c.as<int32_t>(), cap, ctx->rowptr.as<long long>(), nullptr, nullptr, diag);
    P_CUDA(cudaGetLastError());
    P_CUDA(cub::DeviceScan::InclusiveSum(cubtmp.p, tb2, ctx->rowptr.as<long long>(), ctx->rowptr.as<long long>(), nrows + 1, st));
    P_CUDA(cudaMemcpyAsync(&nnz, ctx->rowptr.as<long long>() + nrows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
    P_CUDA(ctx->colind.reserve(std::max<long long>(nnz, 1) * sizeof(int32_t)));
    k_row_columns<true><<<gridr, wpb * 32, smem, st>>>(nrows, ctx->row_begin, ntet, nrl, ncl, ctx->radj_ptr.as<long long>(), ctx->radj.as<unsigned>(),
                                                     ctx->e2c.as<int32_t>(), cap, nullptr, ctx->rowptr.as<long long>(), ctx->colind.as<int32_t>(), diag);
    P_CUDA(cudaGetLastError());
    }
    P_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    k_max_len<<<grid_for(nrows), 256, 0, st>>>(nrows, ctx->rowptr.as<long long>(), ctx->flag.as<int>());
    int max_len = 0;
    P_CUDA(cudaMemcpyAsync(&max_len, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
    if (max_len > 65535) { cleanup(); set_error(ctx, "afb_pattern_build: row longer than 65535 entries"); return -3; }
    // 3. slot table
    ctx->pos_bytes = max_len <= 256 ? 1 : 2;
    P_CUDA(ctx->pos.reserve((size_t)std::max<long long>(1, n_adj) * ncl * ctx->pos_bytes));
    if (ctx->pos_bytes == 1)
        k_pos<unsigned char><<<grid_for(n_adj), 256, 0, st>>>(n_adj, ntet, nrl, ncl, ctx->radj.as<unsigned>(), ctx->e2r.as<int32_t>(), ctx->e2c.as<int32_t>(),
                                                             ctx->rowptr.as<long long>(), ctx->colind.as<int32_t>(), ctx->pos.as<unsigned char>(), ctx->flag.as<int>() + 1);
    else
        k_pos<unsigned short><<<grid_for(n_adj), 256, 0, st>>>(n_adj, ntet, nrl, ncl, ctx->radj.as<unsigned>(), ctx->e2r.as<int32_t>(), ctx->e2c.as<int32_t>(),
                                                              ctx->rowptr.as<long long>(), ctx->colind.as<int32_t>(), ctx->pos.as<unsigned short>(), ctx->flag.as<int>() + 1);
    P_CUDA(cudaGetLastError());
    int missing2[2] = {0, 0};
    P_CUDA(cudaMemcpyAsync(missing2, ctx->flag.as<int>() + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
    const int missing = missing2[0];
    ctx->pos_has_dup = missing2[1] != 0;
    ctx->launches += 5;
    if (missing) { cleanup(); set_error(ctx, "afb_pattern_set: the pattern does not contain every structural entry of the dof map"); return -7; }
#undef P_CUDA
    cleanup();
    ctx->nnz = nnz;
    ctx->max_row_len = max_len;
    ctx->has_pattern = true;
    if (nnz_out) *nnz_out = nnz;
    const int rcp = build_rows_plan(ctx);
    if (rcp) return rcp;
    ctx->rp_prio_valid = false;
    if (ctx->priority_row >= 0 && !ctx->is_sub) { const int rcq = rows_priority_build(ctx, ctx->priority_row); if (rcq) return rcq; }
    // the ring-traversal plan of square P2 problems (afb_rings.cu) is built by the first assembly that can use it
    ctx->has_ring_plan = false; ctx->ring_plan_tried = false; ctx->rg_prio_valid = false;
    return blocks_build(ctx);  // pair plans of vector / mixed spaces; block destinations are searched in whatever pattern is installed
}

int afb_pattern_get(afb_ctx* ctx, int64_t* rowptr, int32_t* colind, int mem_space) {
    if (!ctx) return -7;
    if (!ctx->has_pattern) { set_error(ctx, "pattern was not built (afb_pattern_build)"); return -6; }
    cudaSetDevice(ctx->device);
    const long long nrows = ctx->row_end - ctx->row_begin;
    const cudaMemcpyKind k = mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (rowptr) AFB_CUDA(ctx, cudaMemcpyAsync(rowptr, ctx->rowptr.p, (nrows + 1) * sizeof(long long), k, ctx->stream));
    if (colind && ctx->nnz) AFB_CUDA(ctx, cudaMemcpyAsync(colind, ctx->colind.p, ctx->nnz * sizeof(int32_t), k, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
