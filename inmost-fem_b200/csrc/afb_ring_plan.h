// Plan of the ring-traversal assembly of square P2 problems (k_rings, afb_rings.cu).  Pure host code (no CUDA), so that the
// plan and the algorithm it drives are testable on the CPU (tests/cxx/test_ring_plan.cpp emulates the kernel on this plan).
//
// Idea.  Every entry (r, c) of a P2 matrix is a sum over the elements that contain the mesh entities of r and c, i.e. over the
// star of the join of the two entities.  For all entries except the vertex diagonals that join contains an EDGE of the mesh:
//   * row of edge ab:     all 3 + 4n entries (n = tets around ab) are sums over sub-chains of the ring of tets around ab;
//   * vertex rows:        (a,b), (a,ab) are ring sums of edge ab;  (r, ab) with r opposite to ab in a face is the sum over the
//                         two ring tets that share the face (a, b, r).
// So one thread that walks the ring of an edge IN RING ORDER keeps every partial sum in registers: 7 ring sums, the four sums of
// the ring vertex shared with the previous tet (carried), one entry that only this tet contributes to.  Each matrix entry is
// produced exactly once, by one thread, in a fixed order (deterministic, no atomics), and leaves the thread when it is complete:
// no read-modify-write accumulators at all.  The vertex diagonals (sums over the ball of a vertex) are split into one partial
// sum per incident edge (every tet is assigned to one of its three edges at the vertex) and added by a small second kernel.
// Reference semantics reproduced: the scatter of AssemblerT::Assemble (inmost_interface/assembler.inl:397-481) for the local
// matrices of fem3Dtet (fem/operations/core.inl:277-367) in their tensor representation (afb_tensor.cu).
//
// Canonical frame.  In the frame (a, b, r, s) of a ring tet (a < b the edge by mesh vertex id, r the ring vertex shared with
// the previous tet, s with the next one) the local dof of every produced entry is fixed: a=0, b=1, r=2, s=3, ab=4, ar=5, as=6,
// br=7, bs=8, rs=9.  The element table T[c][i][j] is invariant under relabelling the vertices when the six off-diagonal
// barycentric coefficients G_xy = |T| grad l_x . K grad l_y travel with the labels, so the thread reads the coefficient record
// of the tet through the permutation of its frame and uses ONE set of table rows for every tet of every ring.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace afb {

constexpr int RING_HW = 6;   // header words per (slice, lane)
constexpr int RING_SW = 3;   // plan words per (step, lane)
constexpr unsigned RING_NOIMG = 0xFFFFu;

// step word 0
constexpr unsigned RW0_EL_BITS = 12;          // local element + 1 inside the cluster's staged list (0 = none: zero record)
constexpr unsigned RW0_TAU_SHIFT = 12;        // 3 x 2 bits: record piece feeding canonical piece p
constexpr unsigned RW0_SWAP_SHIFT = 18;       // 3 bits: canonical piece p takes the two halves of its record piece swapped
constexpr unsigned RW0_FLAGA = 1u << 21;      // this tet's diagonal / load contribution of vertex a belongs to this ring
constexpr unsigned RW0_FLAGB = 1u << 22;
constexpr unsigned RW0_EMITR = 1u << 23;      // the group of ring vertex r is complete with this step: store it
constexpr unsigned RW0_HOLDF = 1u << 24;      // first step of a closed ring: keep the r-group, it is completed by the last tet
constexpr unsigned RW0_ADDF = 1u << 25;       // terminal step of a closed ring: add the kept group
// step word 1: slot bytes of (ab, r), (ab, ar), (ab, br), (ab, rs) in row ab;  step word 2: low 16 bits = image offset of (r, ab)
// header words: H0 = image offset of row ab | steps << 16 | valid << 24;  H1 = slot bytes of (ab,a), (ab,b), (ab,ab);
// H2 = image offsets of (a,b) | (a,ab) << 16;  H3 = (b,a) | (b,ab) << 16;  H4 = row of ab (0xFFFFFFFF: empty lane);
// H5 (the same in every lane) = first step of the slice relative to the cluster's first step | steps of the slice << 16

struct RingRowDesc {       // one row image of a cluster (copy-out)
    long long p0;          // first CSR entry of the row
    unsigned short off;    // offset of the image inside the cluster image (doubles) -- images are < 2^16 doubles
    unsigned short len;    // row length
    unsigned short pad0, pad1;
};

struct alignas(16) RingCluster {   // everything a CTA needs to know about a cluster: one 64-byte load
    int e0, ne;            // staged element ids: elist[e0 .. e0+ne)   (e0 is a multiple of 4)
    int sl0, nsl;          // slices
    long long stc0;        // first step
    int nstc;              // steps
    int d0, nd;            // row descriptors (complete edge rows)
    int x0;                // first CSR offset of the vertex-row entries (multiple of 4)
    int vim0, nx;          // image offset / number of the vertex-row entries
    long long xbase;       // CSR position the offsets are relative to
    long long pad;         // bit 0: some edge row of the cluster has entries no local element contributes to (superset pattern):
                           // the kernel zeroes the edge-row part of the image before the ring phase
};
static_assert(sizeof(RingCluster) == 64, "cluster record is one 64-byte line");

struct RingPlanIn {
    long long ntet = 0, nrows = 0;
    const int32_t* v[4] = {nullptr, nullptr, nullptr, nullptr};   // mesh vertices [ntet]
    const int32_t* e2r = nullptr;          // [10*ntet] row codes, all > 0 (local row + 1)
    const long long* rowptr = nullptr;     // [nrows+1]
    const long long* radj_ptr = nullptr;   // [nrows+1]
    const unsigned* radj = nullptr;        // e*10 + i, ascending per row
    const unsigned char* pos = nullptr;    // [n_adj*10] slot of local column j of adjacency entry a in its row
    const unsigned* old2new = nullptr;     // Morton id of every element
    int edges_per_cluster = 256;
    int max_smem_bytes = 227 * 1024;       // shared-memory budget of one cluster (ring_smem_bytes)
    int nthreads = 1;
};

struct RingPlan {
    bool ok = false;
    std::string why;               // reason when !ok (the caller falls back to the row gather)
    long long ncl = 0, nslices = 0, nsteps = 0, nedges = 0, nvert = 0, nstaged = 0;
    int gcap = 0;                  // largest staged element list
    int imgcap = 0;                // largest cluster image (doubles)
    int stepcap = 0;               // most steps of one cluster
    int xcap = 0;                  // most vertex-row entries of one cluster
    size_t smem_bytes = 0;         // shared memory of one CTA with 4 record pieces (ring_smem_bytes)
    int edges_per_cluster = 0;
    std::vector<RingCluster> cinfo;   // [ncl]
    std::vector<int> cs;           // [ncl+1] slices of a cluster
    std::vector<int> eptr;         // [ncl+1] staged element list of a cluster
    std::vector<unsigned> elist;   // Morton ids
    std::vector<long long> sptr;   // [nslices+1] first step of a slice
    std::vector<unsigned> hdr;     // [(slice*RING_HW + w)*32 + lane]
    std::vector<unsigned> steps;   // [(step*RING_SW + w)*32 + lane]
    std::vector<int> dptr;         // [ncl+1] row images of a cluster
    std::vector<RingRowDesc> desc;
    std::vector<int> vimg;         // [2*ncl]: first double of the vertex-row entries, total doubles of the image
    // vertex-row entries (a,b), (a,ab), (r,ab) produced by a cluster: stored compactly behind its edge rows in (row, slot) order
    std::vector<int> xptr;         // [ncl+1]
    std::vector<unsigned> xpos;    // CSR position of the entry relative to xbase[cluster]
    std::vector<long long> xbase;  // [ncl]
    // vertex diagonals + loads: per vertex row the partial sums (indices into the scratch array, 4 doubles per (slice, lane):
    // D_a, D_b, F_a, F_b) in a fixed order
    std::vector<long long> vptr;   // [nvert+1]
    std::vector<unsigned> vlist;   // scratch index of the D partial; the F partial sits 2 doubles further
    std::vector<long long> vdpos;  // [nvert] CSR position of the diagonal entry
    std::vector<unsigned> vrow;    // [nvert] row
    std::vector<long long> zlist;  // CSR positions no element contributes to (superset patterns): zeroed by non-accumulating assemblies
    std::vector<unsigned char> cl_minrow_prio;  // scratch for phased assembly, filled by ring_plan_priority
    std::vector<unsigned> cl_maxrow;            // [ncl] largest row a cluster writes to
};

// shared memory of one k_rings CTA: coefficient planes (parts 16-byte pieces per element + the zero record, capacity rounded to 8),
// cluster image, plan words of every step, slice headers, row descriptors, CSR offsets of the vertex-row entries, element ids
inline size_t ring_smem_bytes(int gcap, int imgcap, int stepcap, int edges_per_cluster, int xcap, int parts) {
    const size_t plane = (((size_t)gcap + 8) & ~(size_t)7) * 16;
    return plane * parts + (((size_t)imgcap * 8 + 15) & ~(size_t)15) + (size_t)stepcap * RING_SW * 128 + (size_t)((edges_per_cluster + 31) / 32) * RING_HW * 128 +
           (size_t)edges_per_cluster * 16 + (((size_t)xcap * 4 + 15) & ~(size_t)15) + (((size_t)gcap * 4 + 15) & ~(size_t)15) + 16;
}

// 0 ok (plan.ok tells whether the mesh / numbering is covered)
int ring_plan_build(const RingPlanIn& in, RingPlan& out);

// Table of the ring kernel from the table of the row gather: TM[c][i][j], c over the symmetric 3x3 coefficient M = (M00, M11, M22,
// M01, M02, M12) in the reference-gradient basis (afb_tensor.cu), to TG[q][i][j], q over the off-diagonal barycentric coefficients
// in record order (G01, G23, G02, G13, G03, G12): M_ab = G_{a+1,b+1}, G_xx = -sum_{y != x} G_xy (the barycentric gradients sum to 0).
void ring_table_from_M(const double* TM, int nloc, double* TG);
// largest violation of TG[pi(q)][pi(i)][pi(j)] == TG[q][i][j] over the 24 vertex relabellings pi (P2 local dofs), relative to max |TG|
double ring_table_symmetry_defect(const double* TG);

// local P2 dof of the edge (x, y), x != y
inline int ring_eidx(int x, int y) {
    static const int t[4][4] = {{-1, 4, 5, 6}, {4, -1, 7, 8}, {5, 7, -1, 9}, {6, 8, 9, -1}};
    return t[x][y];
}
// coefficient record: pieces {G01,G23}, {G02,G13}, {G03,G12}: piece and half of the pair {x,y}
inline void ring_pair_piece(int x, int y, int* piece, int* half) {
    if (x > y) { const int t = x; x = y; y = t; }
    static const int pc[4][4] = {{-1, 0, 1, 2}, {-1, -1, 2, 1}, {-1, -1, -1, 0}, {-1, -1, -1, -1}};
    static const int hf[4][4] = {{-1, 0, 0, 0}, {-1, -1, 1, 1}, {-1, -1, -1, 1}, {-1, -1, -1, -1}};
    *piece = pc[x][y];
    *half = hf[x][y];
}

}  // namespace afb
