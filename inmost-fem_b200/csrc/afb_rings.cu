// Ring-traversal assembly of square P2 problems in the tensor representation (k_rings): the headline kernel.
//
// Reference semantics reproduced: the scatter of AssemblerT::Assemble (inmost_interface/assembler.inl:397-481)
//     matrix[r][c] += A_e(i,j) if |A_e(i,j)| > drop_val;  rhs[r] += F_e(i);  non-finite local value -> status -1
// for the local matrices A_e(i,j) = sum_c T[c][i][j] g_e[c] of fem3Dtet (fem/operations/core.inl:277-367; factorisation in
// afb_tensor.cu).  The algorithm and its plan are described in afb_ring_plan.h: one thread walks the tets around one mesh edge in
// ring order and keeps every partial sum of the entries that edge "owns" in registers, so there are no read-modify-write
// accumulators; every CSR value is produced once, in a fixed order, and written once.
//
// Shape of the kernel (CTA = cluster of edges_per_cluster edges that are close along the Morton curve of their first tet):
//   1. stage the coefficient records (64 bytes: G01,G23 | G02,G13 | G03,G12 | mass, load) of the tets the cluster's rings touch
//      into shared memory (cp.async, 16-byte pieces, one plane per piece, hashed slots: see k_rows_cl);
//   2. ring phase: warp = slice of 32 edges with similar ring length, thread = edge.  Per step (one ring tet): 3 plan words
//      (coalesced, prefetched two steps ahead), 3-4 LDS.128 through the frame permutation, 108-126 DFMAs against ONE set of
//      table rows held in the constant bank (uniform-register operands), 5 stores of finished entries into the cluster image;
//   3. copy-out: the image holds the complete rows of the cluster's edges (stored row by row, coalesced) and the vertex-row
//      entries the cluster produced (compact, one 32-bit CSR offset each).
// The vertex diagonals and vertex loads are partial sums per (edge, end point) in a scratch array, added per vertex in a fixed
// order by k_ring_vertices.  Deterministic, no atomics on data.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>

#include "afb_internal.h"
#include "afb_ring_plan.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

// table rows the ring thread needs, in the canonical frame (a,b,r,s | ab,ar,as,br,bs,rs) = local dofs 0..9
//   k 0..9: (ab, j);  10 (a,b)  11 (a,ab)  12 (b,a)  13 (b,ab)  14 (r,ab)  15 (s,ab)  16 (a,a)  17 (b,b)
constexpr int RT_N = 18;
struct alignas(16) RingTab {
    double A[RT_N][8];   // [k][q]: q 0..5 = G01,G23,G02,G13,G03,G12 (canonical), q 6 = mass coefficient, q 7 unused
    double F[6];         // load weights of ab, a, b
};
const int RT_I[RT_N] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 0, 0, 1, 1, 2, 3, 0, 1};
const int RT_J[RT_N] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 4, 0, 4, 4, 4, 0, 1};

struct RingArgs {
    int gcap;            // staged elements per cluster (capacity of a coefficient plane)
    int zero;            // always 0: keeps the table loads next to their DFMA (see k_rows_cl)
    const int* clist;    // NULL: CTA b works on cluster b; else cluster clist[b]
    const int* cs;       // [ncl+1]
    const int* eptr;     // [ncl+1]
    const unsigned* elist;
    const long long* sptr;
    const unsigned* hdr;
    const unsigned* steps;
    const int* dptr;
    const RingRowDesc* desc;
    const int* vimg;     // [2*ncl]
    const int* xptr;
    const unsigned* xpos;
    const long long* xbase;
    const double* gbuf;  // Morton order, 8 doubles per element
    double* val;
    double* rhs;
    double* scratch;     // 4 doubles per (slice, lane)
    int accumulate;
    double drop;
    int* status;
};

__device__ __forceinline__ void sts64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// slot of an element record inside a coefficient plane (bank-group hash, as in k_rows_cl)
__device__ __forceinline__ unsigned rec_slot(unsigned el1) { return el1 ^ (((el1 >> 3) ^ (el1 >> 6) ^ (el1 >> 9)) & 7u); }

template <bool HASM, bool HASF>
__global__ void __launch_bounds__(256) k_rings(const __grid_constant__ RingTab T, const RingArgs p) {
    constexpr int PARTS = (HASM || HASF) ? 4 : 3;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int c = p.clist ? p.clist[blockIdx.x] : (int)blockIdx.x;
    const int e0 = p.eptr[c], ne = p.eptr[c + 1] - e0;
    const unsigned g_a = (unsigned)__cvta_generic_to_shared(smraw);
    const unsigned plane = (((unsigned)p.gcap + 8u) & ~7u) * 16u;
    // ---- 1. stage the coefficient records: record 0 stays zero (steps without a tet read it)
    {
        const char* gsrc = reinterpret_cast<const char*>(p.gbuf);
        if (threadIdx.x < PARTS) {
            double2 zz; zz.x = 0.0; zz.y = 0.0;
            *reinterpret_cast<double2*>(smraw + (size_t)threadIdx.x * plane) = zz;
        }
        for (int el = threadIdx.x; el < ne; el += blockDim.x) {
            const unsigned id = __ldg(p.elist + e0 + el);
            const unsigned dst = g_a + rec_slot((unsigned)el + 1u) * 16;
#pragma unroll
            for (int part = 0; part < PARTS; ++part) cp_async16(dst + part * plane, gsrc + ((size_t)id * 4 + part) * 16);
        }
        cp_async_commit();
    }
    const unsigned img_a = g_a + plane * PARTS;   // cluster image (doubles)
    const int sl0 = p.cs[c], sl1 = p.cs[c + 1];
    const double drop = p.drop;
    double chk = 0.0;   // NaN iff some coefficient of a visited tet is NaN or +-Inf
    cp_async_wait_all();
    __syncthreads();

    // ---- 2. ring phase
    for (int s = sl0 + warp; s < sl1; s += nwarps) {
        const unsigned* H = p.hdr + (size_t)s * (RING_HW * 32) + lane;
        const unsigned h0 = __ldg(H), h1 = __ldg(H + 32), h2 = __ldg(H + 64), h3 = __ldg(H + 96), h4 = __ldg(H + 128);
        const long long st0 = __ldg(p.sptr + s);
        // warp-uniform trip count in a uniform register (the compiler cannot see that s is the same for the whole warp): the loop
        // counter and with it the table offsets below stay uniform, which is what keeps the table operands on LDCU
        const int nst = __reduce_max_sync(0xffffffffu, (int)(__ldg(p.sptr + s + 1) - st0));
        const unsigned* W = p.steps + (size_t)st0 * (RING_SW * 32) + lane;
        const unsigned ebase = img_a + (h0 & 0xffffu) * 8u;
        // plan words two steps ahead
        unsigned wa0 = 0, wa1 = 0, wa2 = 0, wb0 = 0, wb1 = 0, wb2 = 0;
        if (nst > 0) { wa0 = __ldg(W); wa1 = __ldg(W + 32); wa2 = __ldg(W + 64); }
        if (nst > 1) { wb0 = __ldg(W + 96); wb1 = __ldg(W + 128); wb2 = __ldg(W + 160); }
        double Sa = 0, Sb = 0, Sab = 0, Vab = 0, Va4 = 0, Vba = 0, Vb4 = 0, Da = 0, Db = 0, Fab = 0, Fa = 0, Fb = 0;
        double C0 = 0, C1 = 0, C2 = 0, C3 = 0, K0 = 0, K1 = 0, K2 = 0, K3 = 0;
        for (int st = 0; st < nst; ++st) {
            const unsigned w0 = wa0, w1 = wa1, w2 = wa2;
            wa0 = wb0; wa1 = wb1; wa2 = wb2;
            if (st + 2 < nst) {
                const unsigned* Wn = W + (size_t)(st + 2) * (RING_SW * 32);
                wb0 = __ldg(Wn); wb1 = __ldg(Wn + 32); wb2 = __ldg(Wn + 64);
            }
            // z2 is 0 at run time but loop-variant for the compiler: the table entries stay uniform constant loads (LDCU) next to
            // their DFMA instead of being hoisted out of the loop into vector registers
            const int z2 = (st & p.zero) * 2;
            const unsigned eloc = w0 & ((1u << RW0_EL_BITS) - 1u);
            const unsigned rec = g_a + rec_slot(eloc) * 16;
            double G[6], gm = 0.0, gf = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const unsigned tau = (w0 >> (RW0_TAU_SHIFT + 2 * q)) & 3u;
                const double2 d = lds128(rec + tau * plane);
                const bool sw = (w0 >> (RW0_SWAP_SHIFT + q)) & 1u;
                G[2 * q] = sw ? d.y : d.x;
                G[2 * q + 1] = sw ? d.x : d.y;
            }
            if (HASM || HASF) {
                const double2 d = lds128(rec + 3 * plane);
                gm = d.x; gf = d.y;
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) chk = fma(G[q], 0.0, chk);
            if (HASM) chk = fma(gm, 0.0, chk);
            if (HASF) chk = fma(gf, 0.0, chk);
            // entry k of the table: |A_e(i,j)| > drop_val is the reference's rule (assembler.inl:416)
            auto E = [&](int k) -> double {
                double x = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) x = fma(T.A[k][q + z2], G[q], x);
                if (HASM) x = fma(T.A[k][6 + z2], gm, x);
                return (fabs(x) <= drop) ? 0.0 : x;
            };
            Sa += E(0); Sb += E(1); Sab += E(4);
            Vab += E(10); Va4 += E(11); Vba += E(12); Vb4 += E(13);
            {
                const double da = E(16), db = E(17);
                if (w0 & RW0_FLAGA) Da += da;
                if (w0 & RW0_FLAGB) Db += db;
            }
            if (HASF) {
                Fab = fma(T.F[0 + z2], gf, Fab);
                if (w0 & RW0_FLAGA) Fa = fma(T.F[1 + z2], gf, Fa);
                if (w0 & RW0_FLAGB) Fb = fma(T.F[2 + z2], gf, Fb);
            }
            if (eloc) sts64(ebase + (w1 >> 24) * 8u, E(9));
            double o0 = C0 + E(2), o1 = C1 + E(5), o2 = C2 + E(7), o3 = C3 + E(14);
            if (w0 & RW0_HOLDF) { K0 = o0; K1 = o1; K2 = o2; K3 = o3; }   // step 0 of a closed ring: C is zero, o = this tet's part
            if (w0 & RW0_ADDF) { o0 += K0; o1 += K1; o2 += K2; o3 += K3; }
            if (w0 & RW0_EMITR) {
                sts64(ebase + (w1 & 0xffu) * 8u, o0);
                sts64(ebase + ((w1 >> 8) & 0xffu) * 8u, o1);
                sts64(ebase + ((w1 >> 16) & 0xffu) * 8u, o2);
                sts64(img_a + (w2 & 0xffffu) * 8u, o3);
            }
            C0 = E(3); C1 = E(6); C2 = E(8); C3 = E(15);
        }
        if (h4 != 0xffffffffu) {
            sts64(ebase + (h1 & 0xffu) * 8u, Sa);
            sts64(ebase + ((h1 >> 8) & 0xffu) * 8u, Sb);
            sts64(ebase + ((h1 >> 16) & 0xffu) * 8u, Sab);
            sts64(img_a + (h2 & 0xffffu) * 8u, Vab);
            sts64(img_a + (h2 >> 16) * 8u, Va4);
            sts64(img_a + (h3 & 0xffffu) * 8u, Vba);
            sts64(img_a + (h3 >> 16) * 8u, Vb4);
            double2* sc = reinterpret_cast<double2*>(p.scratch + ((size_t)s * 32 + lane) * 4);
            double2 d0, d1;
            d0.x = Da; d0.y = Db; d1.x = Fa; d1.y = Fb;
            sc[0] = d0; sc[1] = d1;
            if (HASF) { if (p.accumulate) p.rhs[h4] += Fab; else p.rhs[h4] = Fab; }
        }
    }
    __syncthreads();

    // ---- 3. copy-out: complete edge rows (one row per warp step, lanes <-> entries), then the vertex-row entries
    {
        const double* img = reinterpret_cast<const double*>(smraw + (size_t)plane * PARTS);
        const int d0 = p.dptr[c], d1 = p.dptr[c + 1];
        for (int d = d0 + warp; d < d1; d += nwarps) {
            const RingRowDesc R = p.desc[d];
            double* dst = p.val + R.p0;
            const double* src = img + R.off;
            for (int k = lane; k < (int)R.len; k += 32) {
                if (p.accumulate) dst[k] += src[k]; else dst[k] = src[k];
            }
        }
        const int x0 = p.xptr[c], nx = p.xptr[c + 1] - x0;
        const double* vsrc = img + p.vimg[2 * c];
        double* vdst = p.val + p.xbase[c];
        for (int k = threadIdx.x; k < nx; k += blockDim.x) {
            const unsigned off = __ldg(p.xpos + x0 + k);
            if (p.accumulate) vdst[off] += vsrc[k]; else vdst[off] = vsrc[k];
        }
    }
    if (chk != chk) *p.status = 1;  // benign race: every writer stores the same value
}

// vertex diagonals and vertex loads: partial sums of the rings around the vertex, in plan order
__global__ void k_ring_vertices(long long v0, long long v1, const long long* __restrict__ vptr, const unsigned* __restrict__ vlist,
                                const long long* __restrict__ vdpos, const unsigned* __restrict__ vrow, const double* __restrict__ scratch,
                                double* val, double* rhs, int accumulate) {
    for (long long v = v0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; v < v1; v += (long long)gridDim.x * blockDim.x) {
        double d = 0.0, f = 0.0;
        for (long long k = vptr[v]; k < vptr[v + 1]; ++k) {
            const unsigned ix = vlist[k];
            d += scratch[ix];
            f += scratch[ix + 2];
        }
        if (val) { if (accumulate) val[vdpos[v]] += d; else val[vdpos[v]] = d; }
        if (rhs) { if (accumulate) rhs[vrow[v]] += f; else rhs[vrow[v]] = f; }
    }
}

__global__ void k_zero_positions(long long n, const long long* __restrict__ posn, double* val) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) val[posn[i]] = 0.0;
}

template <typename T>
int upload(afb_ctx* ctx, DevBuf& b, const std::vector<T>& v) {
    cudaError_t e = b.reserve(std::max<size_t>(16, v.size() * sizeof(T)));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "ring plan upload (reserve)");
    if (!v.empty()) {
        e = cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "ring plan upload");
    }
    return 0;
}

size_t rings_smem(const afb_ctx* ctx, int parts) {
    const size_t plane = (((size_t)ctx->rg_gcap + 8) & ~(size_t)7) * 16;
    return plane * parts + (size_t)ctx->rg_imgcap * 8 + 16;
}

}  // namespace

namespace afb {

// Builds the ring plan for the current pattern (after build_rows_plan: it reuses the Morton order of the elements).  The plan is
// simply absent (has_ring_plan = false -> row gather k_rows_cl) when the problem is not a square P2 problem on a conforming mesh.
int build_ring_plan(afb_ctx* ctx) {
    ctx->has_ring_plan = false;
    ctx->rg_prio_valid = false;
    if (getenv("AFB_DISABLE_RING_PLAN")) return 0;
    if (!ctx->has_rows_plan || ctx->nrow_loc != 10 || ctx->ncol_loc != 10) return 0;
    if (ctx->pos_bytes != 1 || ctx->has_signs || ctx->pos_has_dup) return 0;
    const long long nrows = ctx->row_end - ctx->row_begin, ntet = ctx->ntet, nadj = ctx->n_adj;
    if (nrows <= 0 || ntet <= 0 || nadj != 10 * ntet) return 0;   // some local row is skipped (code 0): the rings would be incomplete
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> v[4], e2r((size_t)10 * ntet);
    std::vector<long long> rowptr(nrows + 1), radj_ptr(nrows + 1);
    std::vector<unsigned> radj((size_t)nadj), old2new((size_t)ntet);
    std::vector<unsigned char> pos((size_t)nadj * 10);
    for (int k = 0; k < 4; ++k) {
        v[k].resize(ntet);
        AFB_CUDA(ctx, cudaMemcpyAsync(v[k].data(), ctx->v[k].p, ntet * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    AFB_CUDA(ctx, cudaMemcpyAsync(e2r.data(), ctx->e2r.p, e2r.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaMemcpyAsync(rowptr.data(), ctx->rowptr.p, rowptr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaMemcpyAsync(radj_ptr.data(), ctx->radj_ptr.p, radj_ptr.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaMemcpyAsync(radj.data(), ctx->radj.p, radj.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaMemcpyAsync(old2new.data(), ctx->rp_old2new.p, old2new.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaMemcpyAsync(pos.data(), ctx->pos.p, pos.size(), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    for (size_t k = 0; k < e2r.size(); ++k)
        if (e2r[k] <= 0) return 0;
    RingPlanIn in;
    in.ntet = ntet; in.nrows = nrows;
    for (int k = 0; k < 4; ++k) in.v[k] = v[k].data();
    in.e2r = e2r.data(); in.rowptr = rowptr.data(); in.radj_ptr = radj_ptr.data(); in.radj = radj.data(); in.pos = pos.data();
    in.old2new = old2new.data();
    in.nthreads = (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    if (const char* tv = getenv("AFB_PLAN_THREADS")) in.nthreads = std::max(1, atoi(tv));
    // two clusters per SM: image + staged records <= ~110 KB
    int ec = 256;
    if (const char* ev = getenv("AFB_RING_EDGES")) ec = std::max(32, atoi(ev));
    RingPlan pl;
    for (int attempt = 0; attempt < 3; ++attempt, ec /= 2) {
        in.edges_per_cluster = ec;
        in.max_staged = 1400;
        in.max_image_doubles = (220 * 1024 - 64 * (in.max_staged + 8)) / 8;
        ring_plan_build(in, pl);
        if (pl.ok) break;
        if (pl.why.find("cluster") == std::string::npos) break;   // not a size problem: smaller clusters will not help
    }
    if (!pl.ok) {
        if (getenv("AFB_VERBOSE")) fprintf(stderr, "[afb] ring plan not built: %s\n", pl.why.c_str());
        return 0;
    }
    int rc = 0;
    rc = rc ? rc : upload(ctx, ctx->rg_cs, pl.cs);
    rc = rc ? rc : upload(ctx, ctx->rg_eptr, pl.eptr);
    rc = rc ? rc : upload(ctx, ctx->rg_elist, pl.elist);
    rc = rc ? rc : upload(ctx, ctx->rg_sptr, pl.sptr);
    rc = rc ? rc : upload(ctx, ctx->rg_hdr, pl.hdr);
    rc = rc ? rc : upload(ctx, ctx->rg_steps, pl.steps);
    rc = rc ? rc : upload(ctx, ctx->rg_dptr, pl.dptr);
    rc = rc ? rc : upload(ctx, ctx->rg_desc, pl.desc);
    rc = rc ? rc : upload(ctx, ctx->rg_vimg, pl.vimg);
    rc = rc ? rc : upload(ctx, ctx->rg_xptr, pl.xptr);
    rc = rc ? rc : upload(ctx, ctx->rg_xpos, pl.xpos);
    rc = rc ? rc : upload(ctx, ctx->rg_xbase, pl.xbase);
    rc = rc ? rc : upload(ctx, ctx->rg_vptr, pl.vptr);
    rc = rc ? rc : upload(ctx, ctx->rg_vlist, pl.vlist);
    rc = rc ? rc : upload(ctx, ctx->rg_vdpos, pl.vdpos);
    rc = rc ? rc : upload(ctx, ctx->rg_vrow, pl.vrow);
    rc = rc ? rc : upload(ctx, ctx->rg_zlist, pl.zlist);
    if (rc) return rc;
    AFB_CUDA(ctx, ctx->rg_scratch.reserve(std::max<size_t>(16, (size_t)pl.nslices * 32 * 4 * sizeof(double))));
    AFB_CUDA(ctx, cudaMemsetAsync(ctx->rg_scratch.p, 0, (size_t)pl.nslices * 32 * 4 * sizeof(double), st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->rg_ncl = pl.ncl; ctx->rg_nslices = pl.nslices; ctx->rg_nsteps = pl.nsteps; ctx->rg_nvert = pl.nvert;
    ctx->rg_gcap = pl.gcap; ctx->rg_imgcap = pl.imgcap; ctx->rg_nz = (long long)pl.zlist.size();
    ctx->rg_maxrow = pl.cl_maxrow;
    ctx->rg_vrow_host = pl.vrow;
    if (rings_smem(ctx, 4) > 227 * 1024) return 0;
    ctx->has_ring_plan = true;
    if (getenv("AFB_VERBOSE"))
        fprintf(stderr, "[afb] ring plan: %lld edges in %lld clusters of <= %d, %lld slices, %lld steps (%.1f%% lanes busy), %lld staged elements (x%.2f), "
                        "gcap %d, image <= %d doubles, %zu B shared per CTA, %lld unproduced entries\n",
                pl.nedges, pl.ncl, pl.edges_per_cluster, pl.nslices, pl.nsteps, 100.0 * (6.0 * ntet + pl.nedges) / (32.0 * std::max<long long>(1, pl.nsteps)),
                pl.nstaged, (double)pl.nstaged / ntet, pl.gcap, pl.imgcap, rings_smem(ctx, 4), ctx->rg_nz);
    return 0;
}

// clusters that write to a row >= first_priority_row go first (phased assembly, afb_assemble_phase)
int rings_priority_build(afb_ctx* ctx, long long first_priority_row) {
    ctx->rg_prio_valid = false;
    if (!ctx->has_ring_plan) return 0;
    std::vector<int> first, rest;
    for (long long c = 0; c < ctx->rg_ncl; ++c) ((long long)ctx->rg_maxrow[c] >= first_priority_row ? first : rest).push_back((int)c);
    ctx->rg_nprio = (long long)first.size();
    first.insert(first.end(), rest.begin(), rest.end());
    const int rc = upload(ctx, ctx->rg_clist, first);
    if (rc) return rc;
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // vertex rows are numbered in ascending row order: the priority vertices are a suffix
    ctx->rg_vsplit = (long long)(std::lower_bound(ctx->rg_vrow_host.begin(), ctx->rg_vrow_host.end(), (unsigned)std::max<long long>(0, first_priority_row)) -
                                 ctx->rg_vrow_host.begin());
    ctx->rg_prio_valid = true;
    return 0;
}

bool rings_supports(const afb_ctx* ctx, int nstiff, int nmass, int nload) {
    if (!ctx->has_ring_plan || getenv("AFB_DISABLE_RING_KERNEL")) return false;
    return nstiff == 1 && nmass <= 1 && nload <= 1;
}

// TG: [6][10][10] canonical stiffness table (ring_table_from_M), Tm: [10][10] mass table or NULL, Tf: [10] load table or NULL.
// gbuf: Morton order, 8 doubles per element (G01,G23,G02,G13,G03,G12, mass coefficient, load coefficient).
// 1 = launched, < 0 error.
int launch_rings(afb_ctx* ctx, const double* TG, const double* Tm, const double* Tf, const double* gbuf, double* val, double* rhs,
                 int accumulate, double drop_val, int* status, int phase) {
    std::unique_ptr<RingTab> Tp(new RingTab());
    std::memset(Tp.get(), 0, sizeof(RingTab));
    for (int k = 0; k < RT_N; ++k) {
        for (int q = 0; q < 6; ++q) Tp->A[k][q] = TG[((size_t)q * 10 + RT_I[k]) * 10 + RT_J[k]];
        if (Tm) Tp->A[k][6] = Tm[RT_I[k] * 10 + RT_J[k]];
    }
    if (Tf) { Tp->F[0] = Tf[4]; Tp->F[1] = Tf[0]; Tp->F[2] = Tf[1]; }
    RingArgs p;
    p.gcap = ctx->rg_gcap; p.zero = 0; p.clist = nullptr;
    p.cs = ctx->rg_cs.as<int>(); p.eptr = ctx->rg_eptr.as<int>(); p.elist = ctx->rg_elist.as<unsigned>();
    p.sptr = ctx->rg_sptr.as<long long>(); p.hdr = ctx->rg_hdr.as<unsigned>(); p.steps = ctx->rg_steps.as<unsigned>();
    p.dptr = ctx->rg_dptr.as<int>(); p.desc = ctx->rg_desc.as<RingRowDesc>(); p.vimg = ctx->rg_vimg.as<int>();
    p.xptr = ctx->rg_xptr.as<int>(); p.xpos = ctx->rg_xpos.as<unsigned>(); p.xbase = ctx->rg_xbase.as<long long>();
    p.gbuf = gbuf; p.val = val; p.rhs = rhs; p.scratch = ctx->rg_scratch.as<double>();
    p.accumulate = accumulate; p.drop = drop_val; p.status = status;
    long long nblocks = ctx->rg_ncl, v0 = 0, v1 = ctx->rg_nvert;
    if (phase != 0 && ctx->rg_prio_valid) {
        nblocks = phase == 1 ? ctx->rg_nprio : ctx->rg_ncl - ctx->rg_nprio;
        p.clist = ctx->rg_clist.as<int>() + (phase == 1 ? 0 : ctx->rg_nprio);
        if (phase == 1) v0 = ctx->rg_vsplit; else v1 = ctx->rg_vsplit;
    } else if (phase == 2) { nblocks = 0; v1 = 0; }
    cudaStream_t st = ctx->stream;
    if (!accumulate && ctx->rg_nz > 0 && phase != 2) {
        k_zero_positions<<<grid_for(ctx->rg_nz), 256, 0, st>>>(ctx->rg_nz, ctx->rg_zlist.as<long long>(), val);
        ctx->launches++;
    }
    const bool hasm = Tm != nullptr, hasf = Tf != nullptr && rhs != nullptr;
    const size_t smem = rings_smem(ctx, (hasm || hasf) ? 4 : 3);
    if (nblocks > 0) {
#define LAUNCH(M, F)                                                                                                         \
    do {                                                                                                                     \
        cudaError_t e = cudaFuncSetAttribute(k_rings<M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_rings)");                                     \
        k_rings<M, F><<<(unsigned)nblocks, 256, smem, st>>>(*Tp, p);                                                         \
    } while (0)
        if (hasm && hasf) LAUNCH(true, true);
        else if (hasm) LAUNCH(true, false);
        else if (hasf) LAUNCH(false, true);
        else LAUNCH(false, false);
#undef LAUNCH
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(ctx, e, "k_rings launch");
    }
    if (v1 > v0) {
        k_ring_vertices<<<grid_for(v1 - v0), 256, 0, st>>>(v0, v1, ctx->rg_vptr.as<long long>(), ctx->rg_vlist.as<unsigned>(), ctx->rg_vdpos.as<long long>(),
                                                        ctx->rg_vrow.as<unsigned>(), ctx->rg_scratch.as<double>(), val, hasf ? rhs : nullptr, accumulate);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(ctx, e, "k_ring_vertices launch");
    }
    return 1;
}

}  // namespace afb
