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
#include <chrono>
#include <cmath>
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
    int imgcap, stepcap, slicecap, edgecap, xcap;   // capacities of the other shared-memory regions (ring_smem_bytes)
    int zero;            // always 0: keeps the table loads next to their DFMA (see k_rows_cl)
    long long ncl;       // clusters of this launch
    const RingCluster* cinfo;   // [ncl] cluster records in processing order
    const unsigned* elist;
    const unsigned* hdr;
    const unsigned* steps;
    const RingRowDesc* desc;
    const unsigned* xpos;
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
// slot of an element record inside a coefficient plane (bank-group hash, as in k_rows_cl)
__device__ __forceinline__ unsigned rec_slot(unsigned el1) { return el1 ^ (((el1 >> 3) ^ (el1 >> 6) ^ (el1 >> 9)) & 7u); }

__device__ __forceinline__ void cp_async4(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Entries a ring thread evaluates per tet.  The tables of the supported forms are symmetric (T[q][i][j] == T[q][j][i]: symmetric
// coefficient, same space on both sides; checked on the host), so (a,ab) = (ab,a), (b,ab) = (ab,b), (b,a) = (a,b) and
// (r,ab) = (ab,r) are not evaluated a second time: 13 entries, k = 0..9 (row ab), 10 (a,b), 16 (a,a), 17 (b,b).
// RING_MASK_P2STIFF: coefficients q (bit q: G01,G23,G02,G13,G03,G12) with a non-zero table value for the P2 stiffness matrix --
// grad phi_vertex is parallel to one barycentric gradient, grad phi_edge lies in the span of two, so 34 of the 78 products
// remain.  The host compares the actual table with the mask and falls back to the dense instantiation if they disagree.
__device__ constexpr unsigned RING_MASK_P2STIFF[RT_N] = {0x15, 0x29, 0x24, 0x18, 0x3d, 0x30, 0x0c, 0x0c, 0x30, 0x3c, 0x01, 0, 0, 0, 0, 0, 0x15, 0x29};
__device__ constexpr unsigned RING_MASK_DENSE[RT_N] = {0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0x3f, 0, 0, 0, 0, 0, 0x3f, 0x3f};

// DROPF: apply the drop rule |A_e(i,j)| > drop_val per contribution.  Off for drop_val <= 1e-100 (the reference's default,
// assembler.h:212): a dropped contribution is then below 1e-100 in magnitude, i.e. the sums differ by less than 1e-98 absolutely.
//
// Persistent CTAs (a few per SM) walk the clusters with a software pipeline that keeps memory latency off the critical path:
//   iteration i:  wait A_i | barrier | prefetch ids of i+1 | RING PHASE i | wait ids | barrier | issue A_{i+1} | COPY-OUT i | barrier | issue B_{i+1}
// A = coefficient records + plan words + headers (needed by the ring phase), B = row descriptors + CSR offsets (needed by the
// copy-out); the cluster record of i+1 (one 64-byte load) is fetched at the top of iteration i.  [The first version did the
// loads of a cluster at the start of its own CTA: 40 % of the stall samples sat on that chain of three dependent global loads,
// profiles/r02b_rings.md.]
template <bool HASM, bool HASF, bool DROPF, bool SPARSE>
__global__ void __launch_bounds__(256) k_rings(const __grid_constant__ RingTab T, const RingArgs p) {
    constexpr int PARTS = (HASM || HASF) ? 4 : 3;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // shared-memory layout (ring_smem_bytes): coefficient planes | image | step words | slice headers | row descriptors | offsets | ids
    const unsigned g_a = (unsigned)__cvta_generic_to_shared(smraw);
    const unsigned plane = (((unsigned)p.gcap + 8u) & ~7u) * 16u;
    const unsigned img_a = g_a + plane * PARTS;
    const unsigned stp_a = img_a + (((unsigned)p.imgcap * 8u + 15u) & ~15u);
    const unsigned hdr_a = stp_a + (unsigned)p.stepcap * (RING_SW * 128);
    const unsigned dsc_a = hdr_a + (unsigned)p.slicecap * (RING_HW * 128);
    const unsigned xps_a = dsc_a + (unsigned)p.edgecap * 16u;
    const unsigned ids_a = xps_a + (((unsigned)p.xcap * 4u + 15u) & ~15u);
    const double* img = reinterpret_cast<const double*>(smraw + (size_t)(img_a - g_a));
    const RingRowDesc* sdesc = reinterpret_cast<const RingRowDesc*>(smraw + (size_t)(dsc_a - g_a));
    const unsigned* sxpos = reinterpret_cast<const unsigned*>(smraw + (size_t)(xps_a - g_a));
    const unsigned* sids = reinterpret_cast<const unsigned*>(smraw + (size_t)(ids_a - g_a));
    const double drop = p.drop;
    double chk = 0.0;   // NaN iff some coefficient of a visited tet is NaN or +-Inf
    const char* gsrc = reinterpret_cast<const char*>(p.gbuf);
    // p.zero through a warp reduction: the value lives in a uniform register, so that (st & zero_u) below is computed on the uniform
    // datapath and the table operands are fetched by LDCU.  [Read from the parameter bank inside the loop it ended up in a vector
    // register and every table value became a per-thread LDC.64 on the ADU pipe.]
    const int zero_u = __reduce_or_sync(0xffffffffu, p.zero);

    auto issue_ids = [&](const RingCluster& ci) {   // element ids of the cluster (ranges are padded to 4 entries)
        const char* src = reinterpret_cast<const char*>(p.elist + ci.e0);
        for (int k = threadIdx.x; k < (ci.ne + 3) / 4; k += blockDim.x) cp_async16(ids_a + k * 16, src + (size_t)k * 16);
    };
    auto issue_A = [&](const RingCluster& ci) {     // the ids must be in shared memory
        for (int el = threadIdx.x; el < ci.ne; el += blockDim.x) {
            const unsigned id = sids[el];
            const unsigned dst = g_a + rec_slot((unsigned)el + 1u) * 16;
#pragma unroll
            for (int part = 0; part < PARTS; ++part) cp_async16(dst + part * plane, gsrc + ((size_t)id * 4 + part) * 16);
        }
        const char* ssrc = reinterpret_cast<const char*>(p.steps + (size_t)ci.stc0 * (RING_SW * 32));
        for (int k = threadIdx.x; k < ci.nstc * (RING_SW * 8); k += blockDim.x) cp_async16(stp_a + k * 16, ssrc + (size_t)k * 16);
        const char* hsrc = reinterpret_cast<const char*>(p.hdr + (size_t)ci.sl0 * (RING_HW * 32));
        for (int k = threadIdx.x; k < ci.nsl * (RING_HW * 8); k += blockDim.x) cp_async16(hdr_a + k * 16, hsrc + (size_t)k * 16);
    };
    auto issue_B = [&](const RingCluster& ci) {
        const char* dsrc = reinterpret_cast<const char*>(p.desc + ci.d0);
        for (int k = threadIdx.x; k < ci.nd; k += blockDim.x) cp_async16(dsc_a + k * 16, dsrc + (size_t)k * 16);
        const char* xsrc = reinterpret_cast<const char*>(p.xpos + ci.x0);   // ranges are padded to 4 entries: whole 16-byte pieces
        for (int k = threadIdx.x; k < (ci.nx + 3) / 4; k += blockDim.x) cp_async16(xps_a + k * 16, xsrc + (size_t)k * 16);
    };
    auto load_cluster = [&](long long kc) -> RingCluster {   // one 64-byte record, the same address for every thread
        const int4* q = reinterpret_cast<const int4*>(p.cinfo + kc);
        union { int4 v[4]; RingCluster c; } u;
        u.v[0] = __ldg(q); u.v[1] = __ldg(q + 1); u.v[2] = __ldg(q + 2); u.v[3] = __ldg(q + 3);
        return u.c;
    };

    // record 0 of every coefficient plane stays zero (steps without a tet read it); no asynchronous copy ever targets it
    if (threadIdx.x < PARTS) {
        double2 zz; zz.x = 0.0; zz.y = 0.0;
        *reinterpret_cast<double2*>(smraw + (size_t)threadIdx.x * plane) = zz;
    }
    long long kc = blockIdx.x;
    if (kc >= p.ncl) return;
    // cluster records two iterations ahead (clamped index: the loads are unconditional, nothing waits for them until they are used)
    RingCluster cur = load_cluster(kc), nxt = load_cluster(min(kc + (long long)gridDim.x, p.ncl - 1));
    issue_ids(cur);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    issue_A(cur);
    cp_async_commit();
    issue_B(cur);
    cp_async_commit();

    for (; kc < p.ncl; kc += gridDim.x) {
        const bool more = kc + gridDim.x < p.ncl;
        const RingCluster nn = load_cluster(min(kc + 2LL * gridDim.x, p.ncl - 1));
        if (cur.pad & 1) {   // superset pattern (entries only other ranks contribute to): those image slots are never written
            for (int k = threadIdx.x * 2; k < cur.vim0; k += blockDim.x * 2) {
                asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(img_a + (unsigned)k * 8u), "d"(0.0) : "memory");
            }
        }
        cp_async_wait<1>();   // A of this cluster (its B may still be in flight)
        __syncthreads();
        if (more) issue_ids(nxt);   // the id buffer is free: A of this cluster was issued from it before the last barrier
        cp_async_commit();

        // ---- ring phase: no global loads
        for (int sl = warp; sl < cur.nsl; sl += nwarps) {
            const unsigned H = hdr_a + (unsigned)sl * (RING_HW * 128) + lane * 4;
            const unsigned h0 = lds32(H), h1 = lds32(H + 128), h2 = lds32(H + 256), h3 = lds32(H + 384), h4 = lds32(H + 512), h5 = lds32(H + 640);
            // warp-uniform trip count in a uniform register (the compiler cannot see that the slice is the same for the whole
            // warp): the loop counter and with it the table offsets below stay uniform, which keeps the table operands on LDCU
            const int nst = __reduce_max_sync(0xffffffffu, (int)(h5 >> 16));
            unsigned W = stp_a + (h5 & 0xffffu) * (RING_SW * 128) + lane * 4;
            const unsigned ebase = img_a + (h0 & 0xffffu) * 8u;
            double Sa = 0, Sb = 0, Sab = 0, Vab = 0, Da = 0, Db = 0, Fab = 0, Fa = 0, Fb = 0;
            double C0 = 0, C1 = 0, C2 = 0, K0 = 0, K1 = 0, K2 = 0;
            // two steps per trip: the arithmetic of consecutive ring tets is independent (only the three carried sums link them), so
            // the two coefficient fetches and DFMA groups interleave -- the kernel runs with ~3 warps per scheduler, instruction-level
            // parallelism is what hides the latencies.  A missing second step (odd count) is an empty step.
            for (int st = 0; st < nst; st += 2, W += 2 * RING_SW * 128) {
                const bool two = st + 1 < nst;   // uniform
                const unsigned wa0 = lds32(W), wa1 = lds32(W + 128), wa2 = lds32(W + 256);
                unsigned wb0 = lds32(W + 384), wb1 = lds32(W + 512), wb2 = lds32(W + 640);
                if (!two) { wb0 = 0u; wb1 = 0u; wb2 = 0u; }
                // z2 is 0 at run time but loop-variant for the compiler: the table entries stay uniform constant loads (LDCU) next
                // to their DFMA instead of being hoisted out of the loop into vector registers
                const int z2 = (st & zero_u) * 2;
                const unsigned ela = wa0 & ((1u << RW0_EL_BITS) - 1u), elb = wb0 & ((1u << RW0_EL_BITS) - 1u);
                double xa[RT_N], xb[RT_N], gfa = 0.0, gfb = 0.0;
#pragma unroll
                for (int k = 0; k < RT_N; ++k) { xa[k] = 0.0; xb[k] = 0.0; }
                if (__any_sync(0xffffffffu, (ela | elb) != 0u)) {   // terminal / padding steps of the whole slice skip the arithmetic
                    double Ga[6], Gb[6], gma = 0.0, gmb = 0.0;
                    const unsigned reca = g_a + rec_slot(ela) * 16, recb = g_a + rec_slot(elb) * 16;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double2 da = lds128(reca + ((wa0 >> (RW0_TAU_SHIFT + 2 * q)) & 3u) * plane);
                        const double2 db = lds128(recb + ((wb0 >> (RW0_TAU_SHIFT + 2 * q)) & 3u) * plane);
                        const bool swa = (wa0 >> (RW0_SWAP_SHIFT + q)) & 1u, swb = (wb0 >> (RW0_SWAP_SHIFT + q)) & 1u;
                        Ga[2 * q] = swa ? da.y : da.x; Ga[2 * q + 1] = swa ? da.x : da.y;
                        Gb[2 * q] = swb ? db.y : db.x; Gb[2 * q + 1] = swb ? db.x : db.y;
                    }
                    if (HASM || HASF) {
                        const double2 da = lds128(reca + 3 * plane), db = lds128(recb + 3 * plane);
                        gma = da.x; gfa = da.y; gmb = db.x; gfb = db.y;
                    }
                    {   // non-finite test of the coefficients (x*0 is NaN for NaN and +-Inf)
                        const double sa = (Ga[0] * 0.0 + Ga[1] * 0.0) + (Ga[2] * 0.0 + Ga[3] * 0.0) + (Ga[4] * 0.0 + Ga[5] * 0.0);
                        const double sb = (Gb[0] * 0.0 + Gb[1] * 0.0) + (Gb[2] * 0.0 + Gb[3] * 0.0) + (Gb[4] * 0.0 + Gb[5] * 0.0);
                        chk += (sa + sb) + ((HASM ? gma * 0.0 + gmb * 0.0 : 0.0) + (HASF ? gfa * 0.0 + gfb * 0.0 : 0.0));
                    }
                    // the 13 entries of both tets, coefficient-major so that the DFMA chains of different entries interleave.
                    // |A_e(i,j)| > drop_val is the reference's rule (assembler.inl:416)
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
#pragma unroll
                        for (int k = 0; k < RT_N; ++k)
                            if (((SPARSE ? RING_MASK_P2STIFF[k] : RING_MASK_DENSE[k]) >> q) & 1u) {
                                const double t = T.A[k][q + z2];
                                xa[k] = fma(t, Ga[q], xa[k]);
                                xb[k] = fma(t, Gb[q], xb[k]);
                            }
                    }
                    if (HASM) {
#pragma unroll
                        for (int k = 0; k < RT_N; ++k)
                            if (RING_MASK_DENSE[k]) { const double t = T.A[k][6 + z2]; xa[k] = fma(t, gma, xa[k]); xb[k] = fma(t, gmb, xb[k]); }
                    }
                    if (DROPF) {
#pragma unroll
                        for (int k = 0; k < RT_N; ++k)
                            if (RING_MASK_DENSE[k]) { xa[k] = (fabs(xa[k]) <= drop) ? 0.0 : xa[k]; xb[k] = (fabs(xb[k]) <= drop) ? 0.0 : xb[k]; }
                    }
                }
                // ring sums: tet a, then tet b (fixed order)
                Sa = (Sa + xa[0]) + xb[0]; Sb = (Sb + xa[1]) + xb[1]; Sab = (Sab + xa[4]) + xb[4]; Vab = (Vab + xa[10]) + xb[10];
                if (wa0 & RW0_FLAGA) Da += xa[16];
                if (wa0 & RW0_FLAGB) Db += xa[17];
                if (wb0 & RW0_FLAGA) Da += xb[16];
                if (wb0 & RW0_FLAGB) Db += xb[17];
                if (HASF) {
                    const double tf0 = T.F[0 + z2], tf1 = T.F[1 + z2], tf2 = T.F[2 + z2];
                    Fab = fma(tf0, gfa, Fab); Fab = fma(tf0, gfb, Fab);
                    if (wa0 & RW0_FLAGA) Fa = fma(tf1, gfa, Fa);
                    if (wa0 & RW0_FLAGB) Fb = fma(tf2, gfa, Fb);
                    if (wb0 & RW0_FLAGA) Fa = fma(tf1, gfb, Fa);
                    if (wb0 & RW0_FLAGB) Fb = fma(tf2, gfb, Fb);
                }
                if (ela) sts64(ebase + (wa1 >> 24) * 8u, xa[9]);
                if (elb) sts64(ebase + (wb1 >> 24) * 8u, xb[9]);
                // step a: the group of its ring vertex r is complete (carry + this tet), the group of s is carried on
                {
                    double o0 = C0 + xa[2], o1 = C1 + xa[5], o2 = C2 + xa[7];
                    if (wa0 & RW0_HOLDF) { K0 = o0; K1 = o1; K2 = o2; }   // step 0 of a closed ring: the carry is zero, o = this tet's part
                    if (wa0 & RW0_ADDF) { o0 += K0; o1 += K1; o2 += K2; }
                    if (wa0 & RW0_EMITR) {
                        sts64(ebase + (wa1 & 0xffu) * 8u, o0);
                        sts64(ebase + ((wa1 >> 8) & 0xffu) * 8u, o1);
                        sts64(ebase + ((wa1 >> 16) & 0xffu) * 8u, o2);
                        sts64(img_a + (wa2 & 0xffffu) * 8u, o0);   // (r, ab) = (ab, r)
                    }
                    C0 = xa[3]; C1 = xa[6]; C2 = xa[8];
                }
                {
                    double o0 = C0 + xb[2], o1 = C1 + xb[5], o2 = C2 + xb[7];
                    if (wb0 & RW0_HOLDF) { K0 = o0; K1 = o1; K2 = o2; }
                    if (wb0 & RW0_ADDF) { o0 += K0; o1 += K1; o2 += K2; }
                    if (wb0 & RW0_EMITR) {
                        sts64(ebase + (wb1 & 0xffu) * 8u, o0);
                        sts64(ebase + ((wb1 >> 8) & 0xffu) * 8u, o1);
                        sts64(ebase + ((wb1 >> 16) & 0xffu) * 8u, o2);
                        sts64(img_a + (wb2 & 0xffffu) * 8u, o0);
                    }
                    if (two) { C0 = xb[3]; C1 = xb[6]; C2 = xb[8]; }
                }
            }
            if (h4 != 0xffffffffu) {
                sts64(ebase + (h1 & 0xffu) * 8u, Sa);
                sts64(ebase + ((h1 >> 8) & 0xffu) * 8u, Sb);
                sts64(ebase + ((h1 >> 16) & 0xffu) * 8u, Sab);
                sts64(img_a + (h2 & 0xffffu) * 8u, Vab);
                sts64(img_a + (h2 >> 16) * 8u, Sa);      // (a, ab) = (ab, a)
                sts64(img_a + (h3 & 0xffffu) * 8u, Vab);  // (b, a) = (a, b)
                sts64(img_a + (h3 >> 16) * 8u, Sb);      // (b, ab) = (ab, b)
                double2* sc = reinterpret_cast<double2*>(p.scratch + ((size_t)(cur.sl0 + sl) * 32 + lane) * 4);
                double2 d0v, d1v;
                d0v.x = Da; d0v.y = Db; d1v.x = Fa; d1v.y = Fb;
                sc[0] = d0v; sc[1] = d1v;
                if (HASF) { if (p.accumulate) p.rhs[h4] += Fab; else p.rhs[h4] = Fab; }
            }
        }
        cp_async_wait<0>();   // B of this cluster and the ids of the next one
        __syncthreads();
        if (more) issue_A(nxt);   // coefficient planes, step words and headers are free; the copies fly during the copy-out
        cp_async_commit();

        // ---- copy-out: complete edge rows (8 lanes per row, 4 rows per warp pass; rows of up to 32 entries without a loop), then
        //      the vertex-row entries.  [A plain k-loop per row was 23 % of all instructions of the kernel, profiles/r02b_rings.md.]
        {
            const int sub = lane >> 3, l8 = lane & 7;
            for (int d = warp * 4 + sub; d < cur.nd; d += nwarps * 4) {
                const RingRowDesc R = sdesc[d];
                double* dst = p.val + R.p0 + l8;
                const unsigned sa = img_a + ((unsigned)R.off + (unsigned)l8) * 8u;
                const int rem = (int)R.len - l8;   // entries l8, l8+8, ... < len
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                if (rem > 0) v0 = lds64(sa);
                if (rem > 8) v1 = lds64(sa + 64);
                if (rem > 16) v2 = lds64(sa + 128);
                if (rem > 24) v3 = lds64(sa + 192);
                if (p.accumulate) {
                    if (rem > 0) dst[0] += v0;
                    if (rem > 8) dst[8] += v1;
                    if (rem > 16) dst[16] += v2;
                    if (rem > 24) dst[24] += v3;
                    for (int k = 32; k < rem; k += 8) dst[k] += lds64(sa + k * 8);
                } else {
                    if (rem > 0) dst[0] = v0;
                    if (rem > 8) dst[8] = v1;
                    if (rem > 16) dst[16] = v2;
                    if (rem > 24) dst[24] = v3;
                    for (int k = 32; k < rem; k += 8) dst[k] = lds64(sa + k * 8);
                }
            }
            const double* vsrc = img + cur.vim0;
            double* vdst = p.val + cur.xbase;
            if (p.accumulate) { for (int k = threadIdx.x; k < cur.nx; k += blockDim.x) vdst[sxpos[k]] += vsrc[k]; }
            else {
#pragma unroll 4
                for (int k = threadIdx.x; k < cur.nx; k += blockDim.x) vdst[sxpos[k]] = vsrc[k];
            }
        }
        __syncthreads();   // image, descriptors and offsets are free
        if (more) issue_B(nxt);
        cp_async_commit();
        cur = nxt;
        nxt = nn;
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
    return ring_smem_bytes(ctx->rg_gcap, ctx->rg_imgcap, ctx->rg_stepcap, ctx->rg_edges, ctx->rg_xcap, parts);
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
    // one warp per slice of 32 edges; several clusters resident per SM (about 72 KB of shared memory each at 128 edges)
    int ec = 128;
    if (const char* ev = getenv("AFB_RING_EDGES")) ec = std::max(32, std::min(256, atoi(ev) / 32 * 32));
    RingPlan pl;
    for (int attempt = 0; attempt < 3 && ec >= 32; ++attempt, ec /= 2) {
        in.edges_per_cluster = ec;
        in.max_smem_bytes = 227 * 1024;
        ring_plan_build(in, pl);
        if (pl.ok) break;
        if (pl.why.find("cluster") == std::string::npos) break;   // not a size problem: smaller clusters will not help
    }
    if (!pl.ok) {
        if (getenv("AFB_VERBOSE")) fprintf(stderr, "[afb] ring plan not built: %s\n", pl.why.c_str());
        return 0;
    }
    int rc = 0;
    rc = rc ? rc : upload(ctx, ctx->rg_cinfo, pl.cinfo);
    rc = rc ? rc : upload(ctx, ctx->rg_elist, pl.elist);
    rc = rc ? rc : upload(ctx, ctx->rg_hdr, pl.hdr);
    rc = rc ? rc : upload(ctx, ctx->rg_steps, pl.steps);
    rc = rc ? rc : upload(ctx, ctx->rg_desc, pl.desc);
    rc = rc ? rc : upload(ctx, ctx->rg_xpos, pl.xpos);
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
    ctx->rg_stepcap = pl.stepcap; ctx->rg_xcap = pl.xcap; ctx->rg_edges = pl.edges_per_cluster;
    ctx->rg_maxrow = pl.cl_maxrow;
    ctx->rg_cinfo_host.assign(reinterpret_cast<const unsigned char*>(pl.cinfo.data()), reinterpret_cast<const unsigned char*>(pl.cinfo.data() + pl.cinfo.size()));
    ctx->rg_vrow_host = pl.vrow;
    if (rings_smem(ctx, 4) > 227 * 1024) return 0;
    ctx->has_ring_plan = true;
    if (getenv("AFB_VERBOSE"))
        fprintf(stderr, "[afb] ring plan: %lld edges in %lld clusters of <= %d, %lld slices, %lld steps (%.1f%% lanes busy), %lld staged elements (x%.2f), "
                        "gcap %d, image <= %d doubles, %zu B shared per CTA, %lld unproduced entries\n",
                pl.nedges, pl.ncl, pl.edges_per_cluster, pl.nslices, pl.nsteps, 100.0 * (6.0 * ntet + pl.nedges) / (32.0 * std::max<long long>(1, pl.nsteps)),
                pl.nstaged, (double)pl.nstaged / ntet, pl.gcap, pl.imgcap, rings_smem(ctx, 4), ctx->rg_nz);
    if (getenv("AFB_VERBOSE")) fprintf(stderr, "[afb] ring plan: %d steps and %d vertex-row entries per cluster at most\n", pl.stepcap, pl.xcap);
    return 0;
}

int ensure_ring_plan(afb_ctx* ctx) {
    if (ctx->ring_plan_tried || ctx->is_sub) return 0;
    ctx->ring_plan_tried = true;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = build_ring_plan(ctx);
    if (!rc && ctx->has_ring_plan && ctx->priority_row >= 0) rc = rings_priority_build(ctx, ctx->priority_row);
    ctx->ring_plan_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("AFB_VERBOSE")) fprintf(stderr, "[afb] ring plan build: %.1f ms (%s)\n", ctx->ring_plan_ms, ctx->has_ring_plan ? "in use" : "not applicable");
    return rc;
}

// clusters that write to a row >= first_priority_row go first (phased assembly, afb_assemble_phase)
int rings_priority_build(afb_ctx* ctx, long long first_priority_row) {
    ctx->rg_prio_valid = false;
    if (!ctx->has_ring_plan) return 0;
    std::vector<int> first, rest;
    for (long long c = 0; c < ctx->rg_ncl; ++c) ((long long)ctx->rg_maxrow[c] >= first_priority_row ? first : rest).push_back((int)c);
    ctx->rg_nprio = (long long)first.size();
    first.insert(first.end(), rest.begin(), rest.end());
    const RingCluster* all = reinterpret_cast<const RingCluster*>(ctx->rg_cinfo_host.data());
    std::vector<RingCluster> perm(first.size());
    for (size_t k = 0; k < first.size(); ++k) perm[k] = all[first[k]];
    const int rc = upload(ctx, ctx->rg_clist, perm);
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
    p.gcap = ctx->rg_gcap; p.zero = 0;
    p.imgcap = ctx->rg_imgcap; p.stepcap = ctx->rg_stepcap; p.slicecap = (ctx->rg_edges + 31) / 32; p.edgecap = ctx->rg_edges; p.xcap = ctx->rg_xcap;
    p.cinfo = ctx->rg_cinfo.as<RingCluster>();
    p.elist = ctx->rg_elist.as<unsigned>(); p.hdr = ctx->rg_hdr.as<unsigned>(); p.steps = ctx->rg_steps.as<unsigned>();
    p.desc = ctx->rg_desc.as<RingRowDesc>(); p.xpos = ctx->rg_xpos.as<unsigned>();
    p.gbuf = gbuf; p.val = val; p.rhs = rhs; p.scratch = ctx->rg_scratch.as<double>();
    p.accumulate = accumulate; p.drop = drop_val; p.status = status;
    long long nblocks = ctx->rg_ncl, v0 = 0, v1 = ctx->rg_nvert;
    if (phase != 0 && ctx->rg_prio_valid) {
        nblocks = phase == 1 ? ctx->rg_nprio : ctx->rg_ncl - ctx->rg_nprio;
        p.cinfo = ctx->rg_clist.as<RingCluster>() + (phase == 1 ? 0 : ctx->rg_nprio);   // records permuted: priority clusters first
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
        const int nthreads = 32 * std::max(1, (ctx->rg_edges + 31) / 32);   // one warp per slice
        p.ncl = nblocks;
        // persistent CTAs: as many as fit on the device (shared memory bound), each walks clusters blockIdx, blockIdx + grid, ...
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(227 * 1024) / (smem + 1024)));
        if (const char* pv = getenv("AFB_RING_CTAS")) per_sm = std::max(1, atoi(pv));
        const unsigned grid = (unsigned)std::min<long long>(nblocks, 148LL * per_sm);
        const bool dropf = drop_val > 1e-100;
        // sparse instantiation: every table value outside the P2 stiffness mask must vanish (to rounding of the table builder)
        bool sparse = !getenv("AFB_RING_DENSE");
        {
            double mx = 0.0;
            for (int k = 0; k < RT_N; ++k) for (int q = 0; q < 6; ++q) mx = std::max(mx, std::fabs(Tp->A[k][q]));
            const unsigned mask[RT_N] = {0x15, 0x29, 0x24, 0x18, 0x3d, 0x30, 0x0c, 0x0c, 0x30, 0x3c, 0x01, 0, 0, 0, 0, 0, 0x15, 0x29};
            const bool used[RT_N] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1};
            for (int k = 0; k < RT_N; ++k)
                for (int q = 0; q < 6; ++q)
                    if (used[k] && !((mask[k] >> q) & 1u) && std::fabs(Tp->A[k][q]) > 1e-13 * mx) sparse = false;
        }
#define LAUNCH(M, F, D, S)                                                                                                   \
    do {                                                                                                                     \
        cudaError_t e = cudaFuncSetAttribute(k_rings<M, F, D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncSetAttribute(k_rings)");                                     \
        k_rings<M, F, D, S><<<grid, nthreads, smem, st>>>(*Tp, p);                                                           \
    } while (0)
#define LAUNCH_D(M, F)                                                                      \
    do {                                                                                    \
        if (dropf) { if (sparse) LAUNCH(M, F, true, true); else LAUNCH(M, F, true, false); } \
        else { if (sparse) LAUNCH(M, F, false, true); else LAUNCH(M, F, false, false); }     \
    } while (0)
        if (hasm && hasf) LAUNCH_D(true, true);
        else if (hasm) LAUNCH_D(true, false);
        else if (hasf) LAUNCH_D(false, true);
        else LAUNCH_D(false, false);
#undef LAUNCH_D
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
