// Essential boundary conditions on the device.
//
// Reference semantics: every example applies applyDir(A, F, k, bc) (fem/operations/dc_on_dof.h:27-45) to the element
// matrix inside the local assembler for every Dirichlet dof k of the cell (examples/tutorials/ex1.cpp:96-105,
// examples/Fem/Ani/diffusion.cpp:236-245):
//     F(i) -= A(i,k) * bc  for all i;   F(k) = bc;   row k and column k of A zeroed;   A(k,k) = 1.
// Summed over the cells (the boundary value depends only on the global dof) this is, for a global row r:
//     r Dirichlet:  A(r,:) = 0, A(r,r) = deg(r), b(r) = deg(r) * bc_r      (deg = number of cells containing the dof)
//     r free:       b(r) -= sum_{c Dirichlet} A(r,c) * bc_c,  A(r,c) = 0   for the Dirichlet columns c
// which is linear in the assembled (unconstrained) row, so it is applied as a post-pass over the rows that are
// Dirichlet or touch a Dirichlet column (a list built once per pattern); one warp per affected row, deterministic
// (lane-ordered partial sums + fixed shuffle tree).
#include <cub/cub.cuh>

#include <algorithm>

#include "afb_internal.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

// rows [0,nrows): flag = 1 when the row is Dirichlet or has a Dirichlet column
__global__ void k_dir_rows(long long nrows, long long row_begin, const long long* __restrict__ rowptr, const int32_t* __restrict__ colind,
                           const unsigned char* __restrict__ isdir, const int32_t* __restrict__ diag_col, const int32_t* __restrict__ row_gid,
                           unsigned char* flag) {
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = wglob; r < nrows; r += nw) {
        // global id of the row's own dof: explicit table (extended row space of a partitioned mesh), else the forced-diagonal column
        const long long gcol = row_gid ? row_gid[r] : (diag_col ? diag_col[r] : row_begin + r);
        bool any = gcol >= 0 && isdir[gcol];
        for (long long k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) any |= isdir[colind[k]] != 0;
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) flag[r] = any ? 1 : 0;
    }
}

// global dof id of every local row of an explicit dof map whose test and trial sides share the local numbering: the row code and
// the column code of local dof i of an element name the same dof (assembler.inl:139-184 fills indexesR / indexesC from one
// template).  Rows without elements keep the default.  All writers of a row store the same value.
__global__ void k_row_gid_default(long long nrows, long long row_begin, const int32_t* __restrict__ diag_col, int32_t* gid) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x)
        gid[r] = (diag_col && diag_col[r] >= 0) ? diag_col[r] : (int32_t)(row_begin + r);
}
__global__ void k_row_gid_fill(long long n /* nloc*ntet */, const int32_t* __restrict__ e2r, const int32_t* __restrict__ e2c, int32_t* gid) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int rc = e2r[t], cc = e2c[t];
        if (rc != 0 && cc != 0) gid[(rc > 0 ? rc : -rc) - 1] = (cc > 0 ? cc : -cc) - 1;
    }
}

__global__ void k_iota(long long n, int* v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = (int)i;
}

// one warp per affected row
__global__ void k_apply_dir(long long nlist, const int* __restrict__ list, long long row_begin, const long long* __restrict__ rowptr,
                            const int32_t* __restrict__ colind, const long long* __restrict__ radj_ptr, const unsigned char* __restrict__ isdir,
                            const double* __restrict__ bc, const int32_t* __restrict__ diag_col, const int32_t* __restrict__ row_gid,
                            double* val, double* rhs) {
    const int lane = threadIdx.x & 31;
    const long long wglob = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = wglob; t < nlist; t += nw) {
        const long long r = list[t];
        // global dof of the row: explicit table (multi-GPU extended rows), else the forced-diagonal column, else row_begin + r
        const long long gid = row_gid ? row_gid[r] : (diag_col && diag_col[r] >= 0 ? diag_col[r] : row_begin + r);
        const bool rdir = isdir[gid] != 0;
        const long long p0 = rowptr[r], p1 = rowptr[r + 1];
        if (rdir) {
            const double deg = (double)(radj_ptr[r + 1] - radj_ptr[r]);
            if (val)
                for (long long k = p0 + lane; k < p1; k += 32) val[k] = colind[k] == gid ? deg : 0.0;
            if (rhs && lane == 0) rhs[r] = deg * bc[gid];
        } else {
            double s = 0.0;
            for (long long k = p0 + lane; k < p1; k += 32) {
                const int c = colind[k];
                if (isdir[c]) {
                    if (val) { s += val[k] * bc[c]; val[k] = 0.0; }
                }
            }
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (rhs && val && lane == 0) rhs[r] -= s;
        }
    }
}

}  // namespace

namespace afb {

// (re)builds the list of affected rows for the current pattern; 0 ok
static int dirichlet_rows(afb_ctx* ctx) {
    const long long nrows = ctx->row_end - ctx->row_begin;
    cudaStream_t st = ctx->stream;
    DevBuf flag, idx, tmp, nsel;
    auto cleanup = [&]() { flag.release(); idx.release(); tmp.release(); nsel.release(); };
#define D_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
    D_CUDA(flag.reserve(std::max<long long>(1, nrows)));
    D_CUDA(idx.reserve(std::max<long long>(1, nrows) * 4));
    D_CUDA(ctx->dir_rows.reserve(std::max<long long>(1, nrows) * 4));
    D_CUDA(nsel.reserve(8));
    // explicit dof maps with forced-diagonal tables (the extended row space of parallel.InterfacePlan has rows whose own dof lives
    // on another rank: diag_col = -1): the global id of every row comes from the element codes
    ctx->row_gid_valid = false;
    if (ctx->has_diag && ctx->nrow_loc == ctx->ncol_loc) {
        D_CUDA(ctx->row_gid.reserve(std::max<long long>(1, nrows) * 4));
        k_row_gid_default<<<grid_for(nrows), 256, 0, st>>>(nrows, ctx->row_begin, ctx->diag_col.as<int32_t>(), ctx->row_gid.as<int32_t>());
        const long long nn = (long long)ctx->nrow_loc * ctx->ntet;
        k_row_gid_fill<<<grid_for(nn), 256, 0, st>>>(nn, ctx->e2r.as<int32_t>(), ctx->e2c.as<int32_t>(), ctx->row_gid.as<int32_t>());
        ctx->launches += 2;
        ctx->row_gid_valid = true;
    }
    k_dir_rows<<<grid_for(nrows * 32), 256, 0, st>>>(nrows, ctx->row_begin, ctx->rowptr.as<long long>(), ctx->colind.as<int32_t>(),
                                                    ctx->dir_flag.as<unsigned char>(), ctx->has_diag ? ctx->diag_col.as<int32_t>() : nullptr,
                                                    ctx->row_gid_valid ? ctx->row_gid.as<int32_t>() : nullptr, flag.as<unsigned char>());
    k_iota<<<grid_for(nrows), 256, 0, st>>>(nrows, idx.as<int>());
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, idx.as<int>(), flag.as<unsigned char>(), ctx->dir_rows.as<int>(), nsel.as<long long>(), nrows, st);
    D_CUDA(tmp.reserve(tb));
    D_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, idx.as<int>(), flag.as<unsigned char>(), ctx->dir_rows.as<int>(), nsel.as<long long>(), nrows, st));
    long long n = 0;
    D_CUDA(cudaMemcpyAsync(&n, nsel.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    D_CUDA(cudaStreamSynchronize(st));
#undef D_CUDA
    cleanup();
    ctx->launches += 3;
    ctx->n_dir_rows = n;
    ctx->dir_rows_valid = true;
    return 0;
}

int dirichlet_prepare(afb_ctx* ctx) {
    if (!ctx->has_dirichlet || ctx->dir_rows_valid) return 0;
    return dirichlet_rows(ctx);
}

// applied by afb_assemble to this call's contribution (val / rhs may be NULL)
int dirichlet_apply(afb_ctx* ctx, double* val, double* rhs) {
    if (!ctx->has_dirichlet) return 0;
    if (!ctx->dir_rows_valid) {
        const int rc = dirichlet_rows(ctx);
        if (rc) return rc;
    }
    if (ctx->n_dir_rows == 0) return 0;
    k_apply_dir<<<grid_for(ctx->n_dir_rows * 32), 256, 0, ctx->stream>>>(
        ctx->n_dir_rows, ctx->dir_rows.as<int>(), ctx->row_begin, ctx->rowptr.as<long long>(), ctx->colind.as<int32_t>(),
        ctx->radj_ptr.as<long long>(), ctx->dir_flag.as<unsigned char>(), ctx->dir_val.as<double>(),
        ctx->has_diag ? ctx->diag_col.as<int32_t>() : nullptr, ctx->row_gid_valid ? ctx->row_gid.as<int32_t>() : nullptr, val, rhs);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_apply_dir launch");
    return 0;
}

}  // namespace afb

extern "C" int afb_dirichlet_set(afb_ctx* ctx, const unsigned char* is_dirichlet, const double* value, int mem_space) {
    if (!ctx) return -7;
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified"); return -6; }
    cudaSetDevice(ctx->device);
    ctx->dir_rows_valid = false;
    if (!is_dirichlet) { ctx->has_dirichlet = false; return 0; }
    if (!value) { set_error(ctx, "afb_dirichlet_set: values missing"); return -7; }
    const long long n = ctx->ncols_global;
    const cudaMemcpyKind k = mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    AFB_CUDA(ctx, ctx->dir_flag.reserve(std::max<long long>(1, n)));
    AFB_CUDA(ctx, ctx->dir_val.reserve(std::max<long long>(1, n) * sizeof(double)));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->dir_flag.p, is_dirichlet, n, k, ctx->stream));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->dir_val.p, value, n * sizeof(double), k, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_dirichlet = true;
    return 0;
}
