// Context, mesh and dof-map entry points of the C ABI.
//
// Replaces (reference): the per-cell inputs the Assembler pulls out of INMOST -- node coordinates and
// cell->node connectivity (inmost_interface/ordering.inl:37-52, elemental_assembler.cpp:105-112), the
// positive-orientation rule (ordering.inl:8-26), the benchmark mesh generator
// (utils/mesh_utils.cpp:20-48,110-145) and the NATURAL dof enumeration
// (inmost_interface/global_enumerator.cpp:562-605,702-777) -- as flat SoA device arrays.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "afb_internal.h"

namespace afb {

static thread_local std::string g_err;

void set_error(afb_ctx* ctx, const std::string& msg) {
    g_err = msg;
    if (ctx) ctx->err = msg;
}
int cuda_fail(afb_ctx* ctx, cudaError_t e, const char* what) {
    set_error(ctx, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
    return -4;
}
cudaError_t DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p) return cudaSuccess;
    if (p && !borrowed) cudaFree(p);
    p = nullptr; cap = 0; borrowed = false;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes; else p = nullptr;
    return e;
}
void DevBuf::release() {
    if (p && !borrowed) cudaFree(p);
    p = nullptr; cap = 0; borrowed = false;
}

int get_tables(afb_ctx* ctx, int fem, int order, const double** W, const double** phi, const double** grd, bool faces) {
    const int key = faces ? order + 1000 : order;
    for (auto& t : ctx->table_cache)
        if (t.fem == fem && t.order == key) { *W = t.W; *phi = t.phi; *grd = t.grd; return 0; }
    const double *p, *w;
    const int q = faces ? tri_rule(order, &p, &w) : tet_rule(order, &p, &w);
    if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
    OpInfo o;
    if (resolve_op(AFB_IDEN, fem, 1, &o)) { set_error(ctx, "unsupported finite element space"); return -3; }
    const int nf = o.nf_base;
    const int nvar = faces ? 4 : 1;   // faces: the triangle rule lifted to each of the four faces (int_face.inl:175-180)
    std::vector<double> h((size_t)q + (size_t)nvar * q * nf * 4);
    std::copy(w, w + q, h.begin());
    std::vector<double> xyl((size_t)4 * q);
    for (int fc = 0; fc < nvar; ++fc) {
        const double* pts = p;
        if (faces) {
            for (int n = 0; n < q; ++n) {
                xyl[4 * n + fc] = p[3 * n + 0];
                xyl[4 * n + (fc + 1) % 4] = p[3 * n + 1];
                xyl[4 * n + (fc + 2) % 4] = p[3 * n + 2];
                xyl[4 * n + (fc + 3) % 4] = 0.0;
            }
            pts = xyl.data();
        }
        basis_values(fem, q, pts, h.data() + q + (size_t)fc * q * nf);
        basis_ref_grads(fem, q, pts, h.data() + q + (size_t)nvar * q * nf + (size_t)fc * q * nf * 3);
    }
    double* d = nullptr;
    AFB_CUDA(ctx, cudaMalloc(&d, h.size() * sizeof(double)));
    AFB_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TableEntry t{fem, key, d, d + q, d + q + (size_t)nvar * q * nf};
    ctx->table_cache.push_back(t);
    *W = t.W; *phi = t.phi; *grd = t.grd;
    return 0;
}

static cudaMemcpyKind kind_in(int mem_space) { return mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }
static cudaMemcpyKind kind_out(int mem_space) { return mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

}  // namespace afb

using namespace afb;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
namespace {

// nodes of the six tets of a hex in the order CreateNWTetElements creates them; node order inside a
// tet = first appearance in its face list (our stand-in for INMOST's cell->node order).
__constant__ int c_hex_tets[24] = {0, 1, 5, 3, 0, 3, 5, 7, 0, 7, 5, 4, 0, 3, 7, 2, 0, 7, 4, 2, 4, 6, 2, 7};

__global__ void k_cube_nodes(int nx, int ny, int nz, double size, int bx, int by, int bz, int lx, int ly, int lz,
                             double* x, double* y, double* z) {
    const long long n = (long long)(lx + 1) * (ly + 1) * (lz + 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % (lz + 1));
        const int j = (int)((t / (lz + 1)) % (ly + 1));
        const int i = (int)(t / ((long long)(lz + 1) * (ly + 1)));
        // xyz = i * (A - O) / sizes + O with O = 0 (mesh_utils.cpp:118-121)
        x[t] = (bx + i) * (size - 0.0) / nx + 0.0;
        y[t] = (by + j) * (size - 0.0) / ny + 0.0;
        z[t] = (bz + k) * (size - 0.0) / nz + 0.0;
    }
}

__global__ void k_cube_tets(int lx, int ly, int lz, int32_t* v0, int32_t* v1, int32_t* v2, int32_t* v3) {
    const long long nh = (long long)lx * ly * lz;
    for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < nh; h += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(h % lz);
        const int j = (int)((h / lz) % ly);
        const int i = (int)(h / ((long long)lz * ly));
        int hv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            hv[c] = ((i + (c & 1)) * (ly + 1) + (j + ((c >> 1) & 1))) * (lz + 1) + (k + ((c >> 2) & 1));
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const long long e = 6 * h + t;
            v0[e] = hv[c_hex_tets[4 * t + 0]];
            v1[e] = hv[c_hex_tets[4 * t + 1]];
            v2[e] = hv[c_hex_tets[4 * t + 2]];
            v3[e] = hv[c_hex_tets[4 * t + 3]];
        }
    }
}

__global__ void k_orient(long long ntet, const double* x, const double* y, const double* z,
                         const int32_t* v0, const int32_t* v1, int32_t* v2, int32_t* v3) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ntet; e += (long long)gridDim.x * blockDim.x) {
        const int n[4] = {v0[e], v1[e], v2[e], v3[e]};
        double m[3][3];
        for (int r = 0; r < 3; ++r) {
            m[r][0] = x[n[r]] - x[n[3]];
            m[r][1] = y[n[r]] - y[n[3]];
            m[r][2] = z[n[r]] - z[n[3]];
        }
        double det = 0;
        det += m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
        det += m[0][1] * (m[1][2] * m[2][0] - m[1][0] * m[2][2]);
        det += m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        if (det < 0) { v2[e] = n[3]; v3[e] = n[2]; }
    }
}

// entity keys of every (tet, local entity) instance; sorted node tuples
__global__ void k_edge_keys(long long ntet, const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                            unsigned long long* key, unsigned int* inst) {
    const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ntet; e += (long long)gridDim.x * blockDim.x) {
        const unsigned n[4] = {(unsigned)v0[e], (unsigned)v1[e], (unsigned)v2[e], (unsigned)v3[e]};
        for (int l = 0; l < 6; ++l) {
            const unsigned a = min(n[ea[l]], n[eb[l]]), b = max(n[ea[l]], n[eb[l]]);
            key[6 * e + l] = ((unsigned long long)a << 32) | b;
            inst[6 * e + l] = (unsigned)(6 * e + l);
        }
    }
}

// faces: three 32-bit key arrays (sorted triple), local face l = nodes (l, l+1, l+2) % 4
__global__ void k_face_keys(long long ntet, const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                            unsigned* ka, unsigned* kb, unsigned* kc, unsigned* inst) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ntet; e += (long long)gridDim.x * blockDim.x) {
        const unsigned n[4] = {(unsigned)v0[e], (unsigned)v1[e], (unsigned)v2[e], (unsigned)v3[e]};
        for (int l = 0; l < 4; ++l) {
            unsigned a = n[l], b = n[(l + 1) & 3], c = n[(l + 2) & 3], t;
            if (a > b) { t = a; a = b; b = t; }
            if (b > c) { t = b; b = c; c = t; }
            if (a > b) { t = a; a = b; b = t; }
            ka[4 * e + l] = a; kb[4 * e + l] = b; kc[4 * e + l] = c;
            inst[4 * e + l] = (unsigned)(4 * e + l);
        }
    }
}

__global__ void k_gather_u32(long long n, const unsigned* src, const unsigned* perm, unsigned* dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[perm[i]];
}

__global__ void k_heads64(long long n, const unsigned long long* key, unsigned* head) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        head[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}
__global__ void k_heads3(long long n, const unsigned* a, const unsigned* b, const unsigned* c, const unsigned* perm, unsigned* head) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i == 0) { head[i] = 1; continue; }
        const unsigned p = perm[i], q = perm[i - 1];
        head[i] = (a[p] != a[q] || b[p] != b[q] || c[p] != c[q]) ? 1u : 0u;
    }
}
// ids[inst[i]] = scan[i] - 1   (scan = inclusive scan of heads)
__global__ void k_scatter_ids(long long n, const unsigned* scan, const unsigned* inst, int32_t* ids) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) ids[inst[i]] = (int32_t)(scan[i] - 1);
}

struct NatVar { int fem, vec; };
struct NatDesc {
    int nvars;
    NatVar var[8];
    long long off[8][3][4];  // first global index of group (var, comp, etype)
    int nloc;
};

// NATURAL enumeration: writes code = id+1 for local dof slot s of element e at e2[s*ntet + e]
__global__ void k_natural(long long ntet, NatDesc nd, const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                          const int32_t* tet_edge, const int32_t* tet_face, int32_t* e2) {
    const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ntet; e += (long long)gridDim.x * blockDim.x) {
        const int n[4] = {v0[e], v1[e], v2[e], v3[e]};
        int s = 0;
        for (int v = 0; v < nd.nvars; ++v) {
            const int fem = nd.var[v].fem;
            const int nde = fem == AFB_FEM_P2 ? 1 : (fem == AFB_FEM_P3 ? 2 : 0);
            const int ndn = (fem == AFB_FEM_P0) ? 0 : 1;
            const int ndf = fem == AFB_FEM_P3 ? 1 : 0;
            const int ndc = fem == AFB_FEM_P0 ? 1 : 0;
            for (int c = 0; c < nd.var[v].vec; ++c) {
                if (ndn) for (int l = 0; l < 4; ++l) e2[(long long)(s++) * ntet + e] = (int32_t)(nd.off[v][c][0] + n[l] + 1);
                if (nde) for (int l = 0; l < 6; ++l) {
                    const long long base = nd.off[v][c][1] + (long long)tet_edge[6 * e + l] * nde;
                    if (nde == 1) e2[(long long)(s++) * ntet + e] = (int32_t)(base + 1);
                    else {
                        // S2 pair: dof 0 sits next to the endpoint with the smaller global node id
                        const int flip = n[ea[l]] > n[eb[l]] ? 1 : 0;
                        e2[(long long)(s++) * ntet + e] = (int32_t)(base + flip + 1);
                        e2[(long long)(s++) * ntet + e] = (int32_t)(base + 1 - flip + 1);
                    }
                }
                if (ndf) for (int l = 0; l < 4; ++l) e2[(long long)(s++) * ntet + e] = (int32_t)(nd.off[v][c][2] + tet_face[4 * e + l] + 1);
                if (ndc) e2[(long long)(s++) * ntet + e] = (int32_t)(nd.off[v][c][3] + e + 1);
            }
        }
    }
}

// int64 AoS codes [i + nloc*e] -> int32 SoA codes [i*ntet + e], rows rebased to row_begin
// lo/hi: legal range of |code| - 1 (rows: [row_begin, row_end), 0 = ghost row; columns: [0, ncols_global), never 0): the reference
// aborts on an index outside its interval (assembler.inl:399-412); any_neg[1] reports it here
__global__ void k_codes_in(long long ntet, int nloc, const long long* src, long long rebase, long long lo, long long hi, int zero_ok,
                           int32_t* dst, int* any_neg) {
    const long long n = ntet * nloc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long e = t / nloc;
        const int i = (int)(t - e * nloc);
        const long long c = src[t];
        long long o = 0;
        const long long id = (c > 0 ? c : -c) - 1;
        if (c == 0 ? !zero_ok : (id < lo || id >= hi)) { any_neg[1] = 1; dst[(long long)i * ntet + e] = 0; continue; }
        if (c > 0) o = c - rebase;
        else if (c < 0) { o = c + rebase; *any_neg = 1; }
        dst[(long long)i * ntet + e] = (int32_t)o;
    }
}
__global__ void k_codes_out(long long ntet, int nloc, const int32_t* src, long long rebase, long long* dst) {
    const long long n = ntet * nloc;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long e = t / nloc;
        const int i = (int)(t - e * nloc);
        const int c = src[(long long)i * ntet + e];
        dst[t] = c > 0 ? c + rebase : (c < 0 ? c - rebase : 0);
    }
}

__global__ void k_quad_points(long long f, int q, const double* XYL, const double* X0, const double* X1, const double* X2,
                              const double* X3, double* XYG) {
    const long long n = f * q;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / q;
        const int p = (int)(t - r * q);
        for (int k = 0; k < 3; ++k) {
            const double p0 = X0[k + 3 * r];
            double s = p0;
            s += XYL[4 * p + 1] * (X1[k + 3 * r] - p0);
            s += XYL[4 * p + 2] * (X2[k + 3 * r] - p0);
            s += XYL[4 * p + 3] * (X3[k + 3 * r] - p0);
            XYG[k + 3 * t] = s;
        }
    }
}

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

// dst[slot[k]] += src[k]; the slots of one call are distinct, so no atomics are needed
__global__ void k_halo_add(long long n, const long long* __restrict__ slot, const double* __restrict__ src, double* __restrict__ dst) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) dst[slot[k]] += src[k];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int afb_ctx_create(int device, void* cuda_stream, afb_ctx** out) {
    if (!out) return -7;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        set_error(nullptr, "no CUDA device: the anifem_b200 library has no CPU fallback");
        return -4;
    }
    if (device < 0 || device >= ndev) { set_error(nullptr, "bad device index"); return -7; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    afb_ctx* c = new afb_ctx();
    c->device = device;
    if (cuda_stream) c->stream = static_cast<cudaStream_t>(cuda_stream);
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
        c->own_stream = true;
    }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev[i]);
    *out = c;
    return 0;
}

void afb_ctx_destroy(afb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    afb::comm_release(c);
    afb::blocks_clear(c);
    afb::DevBuf* bufs[] = {&c->x, &c->y, &c->z, &c->v[0], &c->v[1], &c->v[2], &c->v[3], &c->e2r, &c->e2c, &c->rowptr, &c->colind,
                           &c->radj_ptr, &c->radj, &c->pos, &c->stageA, &c->stageF, &c->tables, &c->coef, &c->io_val, &c->io_rhs,
                           &c->flag, &c->tmp1, &c->tmp2, &c->tmp3, &c->xy, &c->diag_col, &c->rp_order, &c->rp_cnt, &c->rp_sptr, &c->rp_ell, &c->rp_new2old, &c->rp_old2new, &c->rp_cs, &c->rp_eptr, &c->rp_elist, &c->rp_p0, &c->rp_len, &c->rp_smax, &c->dir_flag, &c->dir_val, &c->dir_rows, &c->rp_clist,
                           &c->bf_tet, &c->bf_face, &c->bf_item, &c->bf_urow, &c->bf_uoff, &c->bf_aidx, &c->row_gid,
                           &c->rg_cinfo, &c->rg_elist, &c->rg_hdr, &c->rg_steps, &c->rg_desc, &c->rg_xpos, &c->rg_vptr, &c->rg_vlist, &c->rg_vdpos, &c->rg_vrow, &c->rg_zlist, &c->rg_scratch, &c->rg_clist};
    for (auto* b : bufs) b->release();
    for (auto& t : c->table_cache) cudaFree(t.W);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* afb_last_error(const afb_ctx* ctx) { return ctx ? ctx->err.c_str() : afb::g_err.c_str(); }

int afb_sync(afb_ctx* ctx) {
    if (!ctx) return -7;
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int64_t afb_launch_count(afb_ctx* ctx, int reset) {
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    if (reset) ctx->launches = 0;
    return n;
}

void* afb_stream_get(afb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int afb_mesh_set(afb_ctx* ctx, int64_t nnode, const double* x, const double* y, const double* z,
                 int64_t ntet, const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3, int mem_space) {
    if (!ctx) return -7;
    if (nnode < 0 || ntet < 0 || (nnode && (!x || !y || !z)) || (ntet && (!v0 || !v1 || !v2 || !v3))) { set_error(ctx, "afb_mesh_set: bad arguments"); return -7; }
    cudaSetDevice(ctx->device);
    const double* xs[3] = {x, y, z};
    afb::DevBuf* xd[3] = {&ctx->x, &ctx->y, &ctx->z};
    for (int k = 0; k < 3; ++k) {
        AFB_CUDA(ctx, xd[k]->reserve(nnode * sizeof(double)));
        if (nnode) AFB_CUDA(ctx, cudaMemcpyAsync(xd[k]->p, xs[k], nnode * sizeof(double), kind_in(mem_space), ctx->stream));
    }
    const int32_t* vs[4] = {v0, v1, v2, v3};
    for (int k = 0; k < 4; ++k) {
        AFB_CUDA(ctx, ctx->v[k].reserve(ntet * sizeof(int32_t)));
        if (ntet) AFB_CUDA(ctx, cudaMemcpyAsync(ctx->v[k].p, vs[k], ntet * sizeof(int32_t), kind_in(mem_space), ctx->stream));
    }
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->nnode = nnode; ctx->ntet = ntet;
    ctx->has_dofmap = false; ctx->has_pattern = false;
    ctx->fields.clear(); ctx->fields_custom = false; afb::blocks_clear(ctx); ctx->bf_plan_valid = false;
    return 0;
}

int afb_mesh_cube(afb_ctx* ctx, int nx, int ny, int nz, double size, int bx, int by, int bz, int lx, int ly, int lz) {
    if (!ctx) return -7;
    if (nx <= 0 || ny <= 0 || nz <= 0 || lx <= 0 || ly <= 0 || lz <= 0 || bx < 0 || by < 0 || bz < 0 ||
        bx + lx > nx || by + ly > ny || bz + lz > nz) { set_error(ctx, "afb_mesh_cube: bad block"); return -7; }
    cudaSetDevice(ctx->device);
    const int64_t nnode = (int64_t)(lx + 1) * (ly + 1) * (lz + 1), ntet = 6LL * lx * ly * lz;
    if (nnode > 2147483000LL || ntet > 2147483000LL) { set_error(ctx, "afb_mesh_cube: mesh too large for int32 ids"); return -7; }
    AFB_CUDA(ctx, ctx->x.reserve(nnode * sizeof(double)));
    AFB_CUDA(ctx, ctx->y.reserve(nnode * sizeof(double)));
    AFB_CUDA(ctx, ctx->z.reserve(nnode * sizeof(double)));
    for (int k = 0; k < 4; ++k) AFB_CUDA(ctx, ctx->v[k].reserve(ntet * sizeof(int32_t)));
    k_cube_nodes<<<grid_for(nnode), 256, 0, ctx->stream>>>(nx, ny, nz, size, bx, by, bz, lx, ly, lz, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>());
    k_cube_tets<<<grid_for(ntet / 6), 256, 0, ctx->stream>>>(lx, ly, lz, ctx->v[0].as<int32_t>(), ctx->v[1].as<int32_t>(), ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>());
    ctx->launches += 2;
    AFB_CUDA(ctx, cudaGetLastError());
    ctx->nnode = nnode; ctx->ntet = ntet;
    ctx->has_dofmap = false; ctx->has_pattern = false;
    ctx->fields.clear(); ctx->fields_custom = false; afb::blocks_clear(ctx); ctx->bf_plan_valid = false;
    return afb_mesh_orient(ctx);
}

int afb_mesh_orient(afb_ctx* ctx) {
    if (!ctx) return -7;
    if (ctx->ntet == 0) return 0;
    cudaSetDevice(ctx->device);
    k_orient<<<grid_for(ctx->ntet), 256, 0, ctx->stream>>>(ctx->ntet, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(),
                                                         ctx->v[0].as<int32_t>(), ctx->v[1].as<int32_t>(), ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>());
    ctx->launches++;
    AFB_CUDA(ctx, cudaGetLastError());
    return 0;
}

int afb_mesh_get(afb_ctx* ctx, int64_t* nnode, int64_t* ntet, double* xyz, int32_t* vv, int mem_space) {
    if (!ctx) return -7;
    if (nnode) *nnode = ctx->nnode;
    if (ntet) *ntet = ctx->ntet;
    cudaSetDevice(ctx->device);
    if (xyz) {
        afb::DevBuf* xd[3] = {&ctx->x, &ctx->y, &ctx->z};
        for (int k = 0; k < 3; ++k) AFB_CUDA(ctx, cudaMemcpyAsync(xyz + k * ctx->nnode, xd[k]->p, ctx->nnode * sizeof(double), kind_out(mem_space), ctx->stream));
    }
    if (vv) for (int k = 0; k < 4; ++k) AFB_CUDA(ctx, cudaMemcpyAsync(vv + k * ctx->ntet, ctx->v[k].p, ctx->ntet * sizeof(int32_t), kind_out(mem_space), ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int afb_dofmap_set(afb_ctx* ctx, int nrow_loc, int ncol_loc, const int64_t* elem2row, const int64_t* elem2col,
                   int64_t row_begin, int64_t row_end, int64_t ncols_global, int mem_space) {
    if (!ctx) return -7;
    if (nrow_loc <= 0 || ncol_loc <= 0 || !elem2row || !elem2col || row_end < row_begin || ncols_global <= 0) { set_error(ctx, "afb_dofmap_set: bad arguments"); return -7; }
    if (ncols_global > 2147483000LL || row_end - row_begin > 2147483000LL) { set_error(ctx, "afb_dofmap_set: index range exceeds int32 per context"); return -7; }
    if ((double)ctx->ntet * nrow_loc >= 4294967295.0) { set_error(ctx, "afb_dofmap_set: ntet*nrow_loc exceeds 2^32 per context"); return -7; }
    cudaSetDevice(ctx->device);
    const int64_t ntet = ctx->ntet;
    AFB_CUDA(ctx, ctx->e2r.reserve((size_t)ntet * nrow_loc * sizeof(int32_t)));
    AFB_CUDA(ctx, ctx->e2c.reserve((size_t)ntet * ncol_loc * sizeof(int32_t)));
    AFB_CUDA(ctx, ctx->flag.reserve(64));
    AFB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, 64, ctx->stream));
    const long long* dr = reinterpret_cast<const long long*>(elem2row);
    const long long* dc = reinterpret_cast<const long long*>(elem2col);
    if (mem_space == AFB_HOST) {
        AFB_CUDA(ctx, ctx->tmp1.reserve((size_t)ntet * std::max(nrow_loc, ncol_loc) * sizeof(long long)));
    }
    for (int side = 0; side < 2; ++side) {
        const int nloc = side ? ncol_loc : nrow_loc;
        const long long* src = side ? dc : dr;
        if (mem_space == AFB_HOST) {
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp1.p, src, (size_t)ntet * nloc * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
            src = ctx->tmp1.as<long long>();
        }
        if (ntet)
            k_codes_in<<<grid_for(ntet * nloc), 256, 0, ctx->stream>>>(ntet, nloc, src, side ? 0 : row_begin, side ? 0 : row_begin, side ? ncols_global : row_end,
                                                                       side ? 0 : 1, side ? ctx->e2c.as<int32_t>() : ctx->e2r.as<int32_t>(), ctx->flag.as<int>());
        ctx->launches++;
        AFB_CUDA(ctx, cudaGetLastError());
        AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    int negbad[2] = {0, 0};
    AFB_CUDA(ctx, cudaMemcpy(negbad, ctx->flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    if (negbad[1]) {
        ctx->has_dofmap = false; ctx->has_pattern = false;
        set_error(ctx, "afb_dofmap_set: index code outside its range (rows: 0 or [row_begin,row_end); columns: [1,ncols_global], never 0)");
        return -7;
    }
    const int neg = negbad[0];
    ctx->has_signs = neg != 0;
    ctx->nrow_loc = nrow_loc; ctx->ncol_loc = ncol_loc;
    ctx->row_begin = row_begin; ctx->row_end = row_end; ctx->ncols_global = ncols_global;
    ctx->has_dofmap = true; ctx->has_pattern = false; ctx->has_diag = false;
    ctx->fields.clear(); ctx->fields_custom = false; afb::blocks_clear(ctx); ctx->bf_plan_valid = false;
    ctx->has_dirichlet = false; ctx->dir_rows_valid = false;
    return 0;
}

int afb_dofmap_set_diag(afb_ctx* ctx, const int64_t* diag_col, int mem_space) {
    if (!ctx) return -7;
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified"); return -6; }
    cudaSetDevice(ctx->device);
    const long long nrows = ctx->row_end - ctx->row_begin;
    if (!diag_col) { ctx->has_diag = false; ctx->has_pattern = false; return 0; }
    std::vector<long long> h(nrows);
    if (mem_space == AFB_DEVICE) AFB_CUDA(ctx, cudaMemcpy(h.data(), diag_col, nrows * sizeof(long long), cudaMemcpyDeviceToHost));
    else std::memcpy(h.data(), diag_col, nrows * sizeof(long long));
    std::vector<int32_t> d(nrows);
    for (long long r = 0; r < nrows; ++r) {
        if (h[r] >= ctx->ncols_global) { set_error(ctx, "afb_dofmap_set_diag: column out of range"); return -7; }
        d[r] = h[r] < 0 ? -1 : (int32_t)h[r];
    }
    AFB_CUDA(ctx, ctx->diag_col.reserve(std::max<long long>(1, nrows) * sizeof(int32_t)));
    AFB_CUDA(ctx, cudaMemcpy(ctx->diag_col.p, d.data(), nrows * sizeof(int32_t), cudaMemcpyHostToDevice));
    ctx->has_diag = true; ctx->has_pattern = false;
    return 0;
}

int afb_dofmap_get(afb_ctx* ctx, int* nrow_loc, int* ncol_loc, int64_t* row_begin, int64_t* row_end, int64_t* ncols_global,
                   int64_t* elem2row, int64_t* elem2col, int mem_space) {
    if (!ctx) return -7;
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified"); return -6; }
    if (nrow_loc) *nrow_loc = ctx->nrow_loc;
    if (ncol_loc) *ncol_loc = ctx->ncol_loc;
    if (row_begin) *row_begin = ctx->row_begin;
    if (row_end) *row_end = ctx->row_end;
    if (ncols_global) *ncols_global = ctx->ncols_global;
    cudaSetDevice(ctx->device);
    for (int side = 0; side < 2; ++side) {
        int64_t* dst = side ? elem2col : elem2row;
        if (!dst || ctx->ntet == 0) continue;
        const int nloc = side ? ctx->ncol_loc : ctx->nrow_loc;
        long long* d = reinterpret_cast<long long*>(dst);
        if (mem_space == AFB_HOST) {
            AFB_CUDA(ctx, ctx->tmp1.reserve((size_t)ctx->ntet * nloc * sizeof(long long)));
            d = ctx->tmp1.as<long long>();
        }
        k_codes_out<<<grid_for(ctx->ntet * nloc), 256, 0, ctx->stream>>>(ctx->ntet, nloc, side ? ctx->e2c.as<int32_t>() : ctx->e2r.as<int32_t>(), side ? 0 : ctx->row_begin, d);
        ctx->launches++;
        AFB_CUDA(ctx, cudaGetLastError());
        if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(dst, d, (size_t)ctx->ntet * nloc * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int afb_dofmap_natural(afb_ctx* ctx, int nvars, const int* fem, const int* vec) {
    if (!ctx) return -7;
    if (nvars <= 0 || nvars > 8 || !fem || !vec) { set_error(ctx, "afb_dofmap_natural: 1..8 variables"); return -7; }
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    cudaSetDevice(ctx->device);
    const long long ntet = ctx->ntet;
    bool need_e = false, need_f = false;
    for (int v = 0; v < nvars; ++v) {
        if (fem[v] < AFB_FEM_P0 || fem[v] > AFB_FEM_P3 || (vec[v] != 1 && vec[v] != 3)) { set_error(ctx, "afb_dofmap_natural: unsupported variable"); return -3; }
        need_e |= fem[v] >= AFB_FEM_P2;
        need_f |= fem[v] == AFB_FEM_P3;
    }
    const int32_t *v0 = ctx->v[0].as<int32_t>(), *v1 = ctx->v[1].as<int32_t>(), *v2 = ctx->v[2].as<int32_t>(), *v3 = ctx->v[3].as<int32_t>();
    long long nedge = 0, nface = 0;
    afb::DevBuf tet_edge, tet_face, keys, keys2, inst, inst2, head, scan, cubtmp;
    auto cleanup = [&]() { tet_edge.release(); tet_face.release(); keys.release(); keys2.release(); inst.release(); inst2.release(); head.release(); scan.release(); cubtmp.release(); };
#define NAT_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
    if (need_e) {
        const long long n = 6 * ntet;
        NAT_CUDA(tet_edge.reserve(n * sizeof(int32_t)));
        NAT_CUDA(keys.reserve(n * sizeof(unsigned long long))); NAT_CUDA(keys2.reserve(n * sizeof(unsigned long long)));
        NAT_CUDA(inst.reserve(n * sizeof(unsigned))); NAT_CUDA(inst2.reserve(n * sizeof(unsigned)));
        NAT_CUDA(head.reserve(n * sizeof(unsigned))); NAT_CUDA(scan.reserve(n * sizeof(unsigned)));
        k_edge_keys<<<grid_for(ntet), 256, 0, ctx->stream>>>(ntet, v0, v1, v2, v3, keys.as<unsigned long long>(), inst.as<unsigned>());
        size_t tb = 0, tb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), inst.as<unsigned>(), inst2.as<unsigned>(), n, 0, 64, ctx->stream);
        cub::DeviceScan::InclusiveSum(nullptr, tb2, head.as<unsigned>(), scan.as<unsigned>(), n, ctx->stream);
        NAT_CUDA(cubtmp.reserve(std::max(tb, tb2)));
        NAT_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), inst.as<unsigned>(), inst2.as<unsigned>(), n, 0, 64, ctx->stream));
        k_heads64<<<grid_for(n), 256, 0, ctx->stream>>>(n, keys2.as<unsigned long long>(), head.as<unsigned>());
        NAT_CUDA(cub::DeviceScan::InclusiveSum(cubtmp.p, tb2, head.as<unsigned>(), scan.as<unsigned>(), n, ctx->stream));
        k_scatter_ids<<<grid_for(n), 256, 0, ctx->stream>>>(n, scan.as<unsigned>(), inst2.as<unsigned>(), tet_edge.as<int32_t>());
        unsigned last = 0;
        NAT_CUDA(cudaMemcpyAsync(&last, scan.as<unsigned>() + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        NAT_CUDA(cudaStreamSynchronize(ctx->stream));
        nedge = last;
        ctx->launches += 5;
    }
    if (need_f) {
        const long long n = 4 * ntet;
        afb::DevBuf ka, kb, kc, kt, kt2;
        auto cl2 = [&]() { ka.release(); kb.release(); kc.release(); kt.release(); kt2.release(); };
#define NAT2_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cl2(); cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
        NAT2_CUDA(tet_face.reserve(n * sizeof(int32_t)));
        NAT2_CUDA(ka.reserve(n * 4)); NAT2_CUDA(kb.reserve(n * 4)); NAT2_CUDA(kc.reserve(n * 4)); NAT2_CUDA(kt.reserve(n * 4)); NAT2_CUDA(kt2.reserve(n * 4));
        NAT2_CUDA(inst.reserve(n * 4)); NAT2_CUDA(inst2.reserve(n * 4)); NAT2_CUDA(head.reserve(n * 4)); NAT2_CUDA(scan.reserve(n * 4));
        k_face_keys<<<grid_for(ntet), 256, 0, ctx->stream>>>(ntet, v0, v1, v2, v3, ka.as<unsigned>(), kb.as<unsigned>(), kc.as<unsigned>(), inst.as<unsigned>());
        size_t tb = 0, tb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, kt.as<unsigned>(), kt2.as<unsigned>(), inst.as<unsigned>(), inst2.as<unsigned>(), n, 0, 32, ctx->stream);
        cub::DeviceScan::InclusiveSum(nullptr, tb2, head.as<unsigned>(), scan.as<unsigned>(), n, ctx->stream);
        NAT2_CUDA(cubtmp.reserve(std::max(tb, tb2)));
        // LSD passes: c, then b, then a (stable) -> lexicographic order of the triples
        unsigned* cur = inst.as<unsigned>();
        unsigned* nxt = inst2.as<unsigned>();
        const unsigned* kk[3] = {kc.as<unsigned>(), kb.as<unsigned>(), ka.as<unsigned>()};
        for (int pass = 0; pass < 3; ++pass) {
            k_gather_u32<<<grid_for(n), 256, 0, ctx->stream>>>(n, kk[pass], cur, kt.as<unsigned>());
            NAT2_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, kt.as<unsigned>(), kt2.as<unsigned>(), cur, nxt, n, 0, 32, ctx->stream));
            std::swap(cur, nxt);
        }
        k_heads3<<<grid_for(n), 256, 0, ctx->stream>>>(n, ka.as<unsigned>(), kb.as<unsigned>(), kc.as<unsigned>(), cur, head.as<unsigned>());
        NAT2_CUDA(cub::DeviceScan::InclusiveSum(cubtmp.p, tb2, head.as<unsigned>(), scan.as<unsigned>(), n, ctx->stream));
        k_scatter_ids<<<grid_for(n), 256, 0, ctx->stream>>>(n, scan.as<unsigned>(), cur, tet_face.as<int32_t>());
        unsigned last = 0;
        NAT2_CUDA(cudaMemcpyAsync(&last, scan.as<unsigned>() + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        NAT2_CUDA(cudaStreamSynchronize(ctx->stream));
        nface = last;
        ctx->launches += 10;
        cl2();
#undef NAT2_CUDA
    }
    // group offsets in NATURAL order (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID)
    NatDesc nd;
    std::memset(&nd, 0, sizeof(nd));
    nd.nvars = nvars;
    long long off = 0;
    int nloc = 0;
    const long long nent[4] = {ctx->nnode, nedge, nface, ntet};
    std::vector<Field> fields;
    for (int v = 0; v < nvars; ++v) {
        nd.var[v].fem = fem[v]; nd.var[v].vec = vec[v];
        const int ndof[4] = {fem[v] == AFB_FEM_P0 ? 0 : 1, fem[v] == AFB_FEM_P2 ? 1 : (fem[v] == AFB_FEM_P3 ? 2 : 0), fem[v] == AFB_FEM_P3 ? 1 : 0, fem[v] == AFB_FEM_P0 ? 1 : 0};
        const int nl1 = 4 * ndof[0] + 6 * ndof[1] + 4 * ndof[2] + ndof[3];
        for (int c = 0; c < vec[v]; ++c) {
            Field f{fem[v], nl1, nloc + c * nl1, off, 0};
            for (int d = 0; d < 4; ++d) {
                nd.off[v][c][d] = off;
                off += nent[d] * ndof[d];
            }
            f.count = off - f.goff;
            f.rows.n = f.cols.n = 1;
            f.rows.start[0] = f.cols.start[0] = f.goff;
            f.rows.count[0] = f.cols.count[0] = f.count;
            fields.push_back(f);
        }
        nloc += vec[v] * nl1;
    }
    nd.nloc = nloc;
    if (off > 2147483000LL) { cleanup(); set_error(ctx, "afb_dofmap_natural: more than 2^31 dofs per context"); return -7; }
    if ((double)ntet * nloc >= 4294967295.0) { cleanup(); set_error(ctx, "afb_dofmap_natural: ntet*nloc exceeds 2^32 per context"); return -7; }
    NAT_CUDA(ctx->e2r.reserve((size_t)ntet * nloc * sizeof(int32_t)));
    NAT_CUDA(ctx->e2c.reserve((size_t)ntet * nloc * sizeof(int32_t)));
    k_natural<<<grid_for(ntet), 256, 0, ctx->stream>>>(ntet, nd, v0, v1, v2, v3, tet_edge.as<int32_t>(), tet_face.as<int32_t>(), ctx->e2r.as<int32_t>());
    ctx->launches++;
    NAT_CUDA(cudaGetLastError());
    NAT_CUDA(cudaMemcpyAsync(ctx->e2c.p, ctx->e2r.p, (size_t)ntet * nloc * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    NAT_CUDA(cudaStreamSynchronize(ctx->stream));
#undef NAT_CUDA
    cleanup();
    ctx->nrow_loc = ctx->ncol_loc = nloc;
    ctx->row_begin = 0; ctx->row_end = off; ctx->ncols_global = off;
    ctx->has_signs = false; ctx->has_diag = false;
    ctx->has_dofmap = true; ctx->has_pattern = false;
    afb::blocks_clear(ctx);
    ctx->fields = fields;
    ctx->fields_custom = false;
    ctx->has_dirichlet = false; ctx->dir_rows_valid = false;
    return 0;
}

// Scalar fields of a caller-supplied dof map.  Under the per-rank NATURAL numbering of a partitioned mesh
// (inmost_interface/global_enumerator.cpp:594-604,702-777) a field (variable, component) is contiguous inside every rank's
// interval only, so its rows (local ids) and columns (global ids) are lists of intervals.
int afb_fields_set(afb_ctx* ctx, int nfields, const int* fem, const int* loff, int nseg_row, const int64_t* row_seg, int nseg_col,
                   const int64_t* col_seg) {
    if (!ctx) return -7;
    if (!ctx->has_dofmap) { set_error(ctx, "dof map was not specified"); return -6; }
    afb::blocks_clear(ctx);
    ctx->fields.clear();
    ctx->fields_custom = false;
    ctx->has_pattern = false;
    if (nfields == 0) return 0;
    if (nfields < 0 || nfields > 8 || !fem || !loff || !row_seg || !col_seg || nseg_row < 1 || nseg_col < 1 || nseg_row > AFB_MAX_SEG ||
        nseg_col > AFB_MAX_SEG) { set_error(ctx, "afb_fields_set: 1..8 fields with 1..8 row / column intervals each"); return -7; }
    std::vector<Field> fields;
    for (int f = 0; f < nfields; ++f) {
        if (fem[f] != AFB_FEM_P1 && fem[f] != AFB_FEM_P2 && fem[f] != AFB_FEM_P3) { set_error(ctx, "afb_fields_set: fields live on P1, P2 or P3"); return -3; }
        Field fl;
        std::memset(&fl, 0, sizeof(fl));
        fl.fem = fem[f];
        fl.nloc = fem[f] == AFB_FEM_P1 ? 4 : (fem[f] == AFB_FEM_P2 ? 10 : 20);
        fl.loff = loff[f];
        if (fl.loff < 0 || fl.loff + fl.nloc > ctx->nrow_loc || fl.loff + fl.nloc > ctx->ncol_loc) { set_error(ctx, "afb_fields_set: local offsets outside the dof map"); return -7; }
        for (int side = 0; side < 2; ++side) {
            SegMap& m = side ? fl.cols : fl.rows;
            const int ns = side ? nseg_col : nseg_row;
            const int64_t* src = (side ? col_seg : row_seg) + (size_t)f * ns * 2;
            const long long limit = side ? ctx->ncols_global : ctx->row_end - ctx->row_begin;
            for (int k = 0; k < ns; ++k) {
                if (src[2 * k] < 0 || src[2 * k + 1] < 0 || src[2 * k] + src[2 * k + 1] > limit) { set_error(ctx, "afb_fields_set: interval outside the row / column space"); return -7; }
                if (src[2 * k + 1] == 0) continue;
                m.start[m.n] = src[2 * k]; m.count[m.n] = src[2 * k + 1]; ++m.n;
            }
        }
        if (fl.rows.total() > 2147483000LL || fl.cols.total() > 2147483000LL) { set_error(ctx, "afb_fields_set: field too large"); return -7; }
        fl.goff = fl.rows.n ? fl.rows.start[0] : 0;
        fl.count = fl.rows.total();
        fields.push_back(fl);
    }
    ctx->fields = fields;
    ctx->fields_custom = true;
    return 0;
}

int afb_halo_add(afb_ctx* ctx, int64_t n, const int64_t* slot, const double* contrib, double* dst) {
    if (!ctx) return -7;
    if (n <= 0) return 0;
    if (!slot || !contrib || !dst) { set_error(ctx, "afb_halo_add: null buffer"); return -7; }
    cudaSetDevice(ctx->device);
    k_halo_add<<<grid_for(n), 256, 0, ctx->stream>>>(n, reinterpret_cast<const long long*>(slot), contrib, dst);
    ctx->launches++;
    AFB_CUDA(ctx, cudaGetLastError());
    return 0;
}

int afb_quad_points(afb_ctx* ctx, int order, int64_t f, const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                    double* XYG, int mem_space) {
    if (!ctx) return -7;
    const double *p, *w;
    const int q = tet_rule(order, &p, &w);
    if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
    if (f <= 0) return q;
    cudaSetDevice(ctx->device);
    const double* X[4] = {XY0, XY1, XY2, XY3};
    double* out = XYG;
    AFB_CUDA(ctx, ctx->tmp2.reserve(4 * q * sizeof(double)));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp2.p, p, 4 * q * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (mem_space == AFB_HOST) {
        AFB_CUDA(ctx, ctx->xy.reserve((size_t)12 * f * sizeof(double)));
        AFB_CUDA(ctx, ctx->tmp3.reserve((size_t)3 * q * f * sizeof(double)));
        for (int k = 0; k < 4; ++k) {
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->xy.as<double>() + (size_t)3 * f * k, X[k], (size_t)3 * f * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            X[k] = ctx->xy.as<double>() + (size_t)3 * f * k;
        }
        out = ctx->tmp3.as<double>();
    }
    k_quad_points<<<grid_for(f * q), 256, 0, ctx->stream>>>(f, q, ctx->tmp2.as<double>(), X[0], X[1], X[2], X[3], out);
    ctx->launches++;
    AFB_CUDA(ctx, cudaGetLastError());
    if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(XYG, out, (size_t)3 * q * f * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return q;
}

int afb_fem3dtet_batched(afb_ctx* ctx, const afb_form* form, int64_t f, const double* XY0, const double* XY1, const double* XY2,
                         const double* XY3, double* A, int mem_space) {
    if (!ctx || !form) return -7;
    if (f <= 0) return 0;  // int_tet.inl:7
    if (!XY0 || !XY1 || !XY2 || !XY3 || !A) { set_error(ctx, "afb_fem3dtet_batched: null buffer"); return -7; }
    cudaSetDevice(ctx->device);
    afb::OpInfo oa, ob;
    if (afb::resolve_op(form->opA, form->femA, form->vecA, &oa) || afb::resolve_op(form->opB, form->femB, form->vecB, &ob)) {
        set_error(ctx, "unsupported operator/space");
        return -3;
    }
    const int dlen = afb::form_dlen(*form, oa, ob);
    const double* X[4] = {XY0, XY1, XY2, XY3};
    // coordinates: one buffer of 4 blocks 3 x f
    AFB_CUDA(ctx, ctx->xy.reserve((size_t)12 * f * sizeof(double)));
    for (int k = 0; k < 4; ++k)
        AFB_CUDA(ctx, cudaMemcpyAsync(ctx->xy.as<double>() + (size_t)3 * f * k, X[k], (size_t)3 * f * sizeof(double), kind_in(mem_space), ctx->stream));
    // coefficient
    const double* Dd = form->D;
    if (dlen > 0) {
        if (!form->D) { set_error(ctx, "tensor data missing"); return -7; }
        const int q = afb_tet_quadrature(form->quad_order, nullptr, nullptr, 0);
        if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
        const size_t n = form->coef_layout == AFB_COEF_CONST ? 1 : (form->coef_layout == AFB_COEF_PER_TET ? (size_t)f : (size_t)f * q);
        if (form->coef_space == AFB_HOST) {
            AFB_CUDA(ctx, ctx->coef.reserve(n * dlen * sizeof(double)));
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->coef.p, form->D, n * dlen * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            Dd = ctx->coef.as<double>();
        }
    }
    double* out = A;
    const size_t asz = (size_t)oa.nfa * ob.nfa * f;
    if (mem_space == AFB_HOST) {
        AFB_CUDA(ctx, ctx->stageA.reserve(asz * sizeof(double)));
        out = ctx->stageA.as<double>();
    }
    afb_form fm = *form;
    fm.row_off = 0; fm.col_off = 0;
    // alpha is used as given (0 contributes nothing), exactly like afb_assemble
    // reference layout A[ib + nfB*(ia + nfA*r)]
    // square scalar forms with the same operator on both sides: the register-tiled kernel (afb_element.cu, k_element_sq)
    {
        const bool same = oa.vec == 1 && ob.vec == 1 && oa.fem == ob.fem && fm.opA == fm.opB && (fm.opA == AFB_GRAD || fm.opA == AFB_IDEN);
        const bool dims_ok = fm.tensor_type < AFB_TENSOR_SYMMETRIC || dlen == (fm.opA == AFB_GRAD ? 9 : 1);
        const bool valid = fm.tensor_type >= AFB_TENSOR_NULL && fm.tensor_type <= AFB_TENSOR_GENERAL && fm.coef_layout >= AFB_COEF_CONST &&
                           fm.coef_layout <= AFB_COEF_PER_POINT;
        if (same && dims_ok && valid && (oa.nfa == 4 || oa.nfa == 10 || oa.nfa == 20) && !getenv("AFB_DISABLE_SQ_ELEMENT")) {
            std::vector<afb_form> fv(1, fm);
            std::vector<afb::OpInfo> ov(1, oa);
            std::vector<const double*> dv(1, Dd);
            int rcs = afb::launch_forms_sq(ctx, fv, ov, dv, std::vector<int>(1, 0), 0, f, out, (long long)oa.nfa * ob.nfa, ctx->xy.as<double>(), 1);
            if (rcs) return rcs;
            if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(A, out, asz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return 0;
        }
    }
    int rc = afb::launch_form(ctx, fm, oa, ob, f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->xy.as<double>(),
                              out, (long long)oa.nfa * ob.nfa, 1, ob.nfa, 0, Dd);
    if (rc) return rc;
    if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(A, out, asz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Ani::fem3Dface (fem/operations/int_face.h:15-159, int_face.inl:160-199): element matrices of the surface integral over face
// face_num[r] of tet r.  Same operators, tensor kinds and layouts as afb_fem3dtet_batched; the rule is the reference's triangle
// rule of that order (PER_POINT coefficients follow its points), the measure the face area.
int afb_fem3dface_batched(afb_ctx* ctx, const afb_form* form, int64_t f, const int32_t* face_num, const double* XY0, const double* XY1,
                          const double* XY2, const double* XY3, double* A, int mem_space) {
    if (!ctx || !form) return -7;
    if (f <= 0) return 0;
    if (!XY0 || !XY1 || !XY2 || !XY3 || !A || !face_num) { set_error(ctx, "afb_fem3dface_batched: null buffer"); return -7; }
    cudaSetDevice(ctx->device);
    afb::OpInfo oa, ob;
    if (afb::resolve_op(form->opA, form->femA, form->vecA, &oa) || afb::resolve_op(form->opB, form->femB, form->vecB, &ob)) {
        set_error(ctx, "unsupported operator/space");
        return -3;
    }
    const int dlen = afb::form_dlen(*form, oa, ob);
    const double* X[4] = {XY0, XY1, XY2, XY3};
    AFB_CUDA(ctx, ctx->xy.reserve((size_t)12 * f * sizeof(double)));
    for (int k = 0; k < 4; ++k)
        AFB_CUDA(ctx, cudaMemcpyAsync(ctx->xy.as<double>() + (size_t)3 * f * k, X[k], (size_t)3 * f * sizeof(double), kind_in(mem_space), ctx->stream));
    const int32_t* dface = face_num;
    if (mem_space == AFB_HOST) {
        for (int64_t r = 0; r < f; ++r)
            if (face_num[r] < 0 || face_num[r] > 3) { set_error(ctx, "Wrong face index"); return -7; }   // int_face.inl:28
        AFB_CUDA(ctx, ctx->tmp1.reserve((size_t)f * sizeof(int32_t)));
        AFB_CUDA(ctx, cudaMemcpyAsync(ctx->tmp1.p, face_num, (size_t)f * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        dface = ctx->tmp1.as<int32_t>();
    }
    const double* Dd = form->D;
    if (dlen > 0) {
        if (!form->D) { set_error(ctx, "tensor data missing"); return -7; }
        const int q = afb_tri_quadrature(form->quad_order, nullptr, nullptr, 0);
        if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
        const size_t n = form->coef_layout == AFB_COEF_CONST ? 1 : (form->coef_layout == AFB_COEF_PER_TET ? (size_t)f : (size_t)f * q);
        if (form->coef_space == AFB_HOST) {
            AFB_CUDA(ctx, ctx->coef.reserve(n * dlen * sizeof(double)));
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->coef.p, form->D, n * dlen * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            Dd = ctx->coef.as<double>();
        }
    }
    double* out = A;
    const size_t asz = (size_t)oa.nfa * ob.nfa * f;
    if (mem_space == AFB_HOST) {
        AFB_CUDA(ctx, ctx->stageA.reserve(asz * sizeof(double)));
        out = ctx->stageA.as<double>();
    }
    afb_form fm = *form;
    fm.row_off = 0; fm.col_off = 0;
    // alpha is used as given (0 contributes nothing), exactly like afb_assemble
    int rc = afb::launch_form(ctx, fm, oa, ob, f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->xy.as<double>(),
                              out, (long long)oa.nfa * ob.nfa, 1, ob.nfa, 0, Dd, dface, nullptr);
    if (rc) return rc;
    if (mem_space == AFB_HOST) AFB_CUDA(ctx, cudaMemcpyAsync(A, out, asz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
