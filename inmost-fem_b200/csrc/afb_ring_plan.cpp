// Host builder of the ring plan (afb_ring_plan.h).  Inputs are host copies of structures the pattern builder made on the
// device (row adjacency, slot table) plus the mesh; everything here is integer work, done once per pattern (setup).
#include "afb_ring_plan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>

namespace afb {

namespace {

template <typename F>
void parallel_for(long long n, int nthreads, F f) {   // f(thread, begin, end) on contiguous ranges
    nthreads = (int)std::max<long long>(1, std::min<long long>(nthreads, n));
    if (nthreads == 1) { f(0, 0LL, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([=]() { f(t, n * t / nthreads, n * (t + 1) / nthreads); });
    for (auto& x : th) x.join();
}

const int EDGE_V[10][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

enum Fail { F_OK = 0, F_MIXED_ROW, F_RING_LONG, F_EDGE_ENDS, F_NONMANIFOLD, F_NOADJ, F_DUP, F_RANGE };
const char* fail_text(int f) {
    switch (f) {
        case F_MIXED_ROW: return "a row mixes vertex and edge dofs";
        case F_RING_LONG: return "more than 62 tets around an edge";
        case F_EDGE_ENDS: return "the tets of an edge row do not share one vertex pair";
        case F_NONMANIFOLD: return "the tets around an edge do not form one chain or cycle";
        case F_NOADJ: return "adjacency entry of an element missing in a vertex row";
        case F_DUP: return "a vertex-row entry would be produced twice";
        case F_RANGE: return "vertex-row entries of a cluster span more than 2^32 CSR positions";
    }
    return "";
}

struct Frame { int la, lb, lr, ls; };
inline unsigned short pack_vis(int korig, const Frame& f) { return (unsigned short)(korig | (f.la << 8) | (f.lb << 10) | (f.lr << 12) | (f.ls << 14)); }
inline void unpack_vis(unsigned short w, int* korig, Frame* f) {
    *korig = w & 0xff; f->la = (w >> 8) & 3; f->lb = (w >> 10) & 3; f->lr = (w >> 12) & 3; f->ls = (w >> 14) & 3;
}

// adjacency entry (position in radj) of (element e, local dof i) in row `row`; -1 if missing
inline long long find_adj(const RingPlanIn& in, long long row, unsigned want) {
    long long lo = in.radj_ptr[row], hi = in.radj_ptr[row + 1];
    const long long end = hi;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (in.radj[mid] < want) lo = mid + 1; else hi = mid; }
    return (lo < end && in.radj[lo] == want) ? lo : -1;
}

struct ClusterLayout {
    std::vector<unsigned> edges;     // edge rows in slice / lane order (0xFFFFFFFF pads the last slice)
    std::vector<int> slice_steps;    // steps of every slice
    std::vector<unsigned> elist;     // staged elements (Morton ids, ascending)
    std::vector<int> ebase;          // image offsets of the edge rows (doubles)
    int image = 0, vimg0 = 0, nvent = 0;   // image size, first / number of vertex-row entries (stored compactly after the edge rows)
};

}  // namespace

void ring_table_from_M(const double* TM, int nloc, double* TG) {
    const size_t n = (size_t)nloc * nloc;
    const double *M00 = TM, *M11 = TM + n, *M22 = TM + 2 * n, *M01 = TM + 3 * n, *M02 = TM + 4 * n, *M12 = TM + 5 * n;
    double *G01 = TG, *G23 = TG + n, *G02 = TG + 2 * n, *G13 = TG + 3 * n, *G03 = TG + 4 * n, *G12 = TG + 5 * n;
    for (size_t k = 0; k < n; ++k) {
        G01[k] = -M00[k]; G02[k] = -M11[k]; G03[k] = -M22[k];
        G12[k] = M01[k] - M00[k] - M11[k];
        G13[k] = M02[k] - M00[k] - M22[k];
        G23[k] = M12[k] - M11[k] - M22[k];
    }
}

double ring_table_symmetry_defect(const double* TG) {
    static const int QP[6][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}, {0, 3}, {1, 2}};   // record order of the pairs
    int perm[4] = {0, 1, 2, 3};
    double worst = 0, scale = 0;
    for (int k = 0; k < 600; ++k) scale = std::max(scale, std::fabs(TG[k]));
    do {
        int dofp[10];
        for (int x = 0; x < 4; ++x) dofp[x] = perm[x];
        for (int i = 4; i < 10; ++i) dofp[i] = ring_eidx(perm[EDGE_V[i][0]], perm[EDGE_V[i][1]]);
        for (int q = 0; q < 6; ++q) {
            int pc, hf;
            ring_pair_piece(perm[QP[q][0]], perm[QP[q][1]], &pc, &hf);
            const int qp = 2 * pc + hf;
            for (int i = 0; i < 10; ++i)
                for (int j = 0; j < 10; ++j)
                    worst = std::max(worst, std::fabs(TG[(qp * 10 + dofp[i]) * 10 + dofp[j]] - TG[(q * 10 + i) * 10 + j]));
        }
    } while (std::next_permutation(perm, perm + 4));
    return scale > 0 ? worst / scale : 0.0;
}

int ring_plan_build(const RingPlanIn& in, RingPlan& out) {
    out = RingPlan();
    const long long nrows = in.nrows, ntet = in.ntet;
    const int nth = std::max(1, in.nthreads);
    if (nrows <= 0 || ntet <= 0) { out.why = "empty problem"; return 0; }
    const long long nadj = in.radj_ptr[nrows];

    // ---- pass 0/1: row types, ring order of every edge row
    std::vector<unsigned char> rtype(nrows, 0), closed(nrows, 0);
    std::vector<unsigned short> ringvis((size_t)nadj);
    std::vector<unsigned> minmort(nrows, 0xffffffffu);
    std::atomic<int> fail(F_OK);
    parallel_for(nrows, nth, [&](int, long long r0, long long r1) {
        for (long long r = r0; r < r1; ++r) {
            const long long a0 = in.radj_ptr[r], a1 = in.radj_ptr[r + 1];
            const int n = (int)(a1 - a0);
            if (n == 0) continue;
            const bool edge = (in.radj[a0] % 10u) >= 4;
            for (long long a = a0; a < a1; ++a)
                if (((in.radj[a] % 10u) >= 4) != edge) { fail = F_MIXED_ROW; return; }
            rtype[r] = edge ? 2 : 1;
            unsigned mm = 0xffffffffu;
            for (long long a = a0; a < a1; ++a) mm = std::min(mm, in.old2new[in.radj[a] / 10u]);
            minmort[r] = mm;
            if (!edge) continue;
            if (n > 62) { fail = F_RING_LONG; return; }
            Frame fr[62];
            int32_t pr[62], ps[62];   // ring vertices (mesh ids) at local positions lr, ls
            int32_t va = -1, vb = -1;
            for (int k = 0; k < n; ++k) {
                const unsigned t = in.radj[a0 + k];
                const long long e = t / 10u;
                const int i = (int)(t % 10u);
                int la = EDGE_V[i][0], lb = EDGE_V[i][1];
                if (in.v[la][e] > in.v[lb][e]) std::swap(la, lb);
                if (k == 0) { va = in.v[la][e]; vb = in.v[lb][e]; }
                else if (in.v[la][e] != va || in.v[lb][e] != vb) { fail = F_EDGE_ENDS; return; }
                int o[2], no = 0;
                for (int x = 0; x < 4; ++x) if (x != la && x != lb) o[no++] = x;
                fr[k] = Frame{la, lb, o[0], o[1]};
                pr[k] = in.v[o[0]][e]; ps[k] = in.v[o[1]][e];
            }
            // occurrences of every ring vertex
            int nends = 0;
            int32_t endv[2] = {0, 0};
            for (int k = 0; k < n; ++k)
                for (int side = 0; side < 2; ++side) {
                    const int32_t w = side ? ps[k] : pr[k];
                    int cnt = 0;
                    for (int m = 0; m < n; ++m) cnt += (pr[m] == w) + (ps[m] == w);
                    if (cnt > 2) { fail = F_NONMANIFOLD; return; }
                    if (cnt == 1) { if (nends < 2) endv[nends] = w; ++nends; }
                }
            if (nends != 0 && nends != 2) { fail = F_NONMANIFOLD; return; }
            closed[r] = nends == 0;
            // start: open chain at its end with the smaller vertex id; closed cycle at the first visit
            int cur = 0;
            int32_t rv;
            if (nends == 2) {
                const int32_t w = std::min(endv[0], endv[1]);
                for (int k = 0; k < n; ++k) if (pr[k] == w || ps[k] == w) { cur = k; break; }
                rv = w;
            } else rv = pr[0];
            bool used[62];
            for (int k = 0; k < n; ++k) used[k] = false;
            for (int step = 0; step < n; ++step) {
                if (cur < 0 || used[cur]) { fail = F_NONMANIFOLD; return; }
                used[cur] = true;
                Frame f = fr[cur];
                if (ps[cur] == rv) { std::swap(f.lr, f.ls); std::swap(pr[cur], ps[cur]); }   // orient: r = shared with the previous tet
                ringvis[a0 + step] = pack_vis(cur, f);
                const int32_t sv = ps[cur];
                int nxt = -1;
                for (int m = 0; m < n; ++m) if (!used[m] && (pr[m] == sv || ps[m] == sv)) { nxt = m; break; }
                cur = nxt;
                rv = sv;
            }
        }
    });
    if (fail.load() != F_OK) { out.why = fail_text(fail.load()); return 0; }

    const bool verbose = getenv("RING_PLAN_VERBOSE") != nullptr;
    auto tlast = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[ring plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - tlast).count());
        tlast = t;
    };
    lap("ring order");
    // ---- edges in Morton order of their first tet; vertex rows
    std::vector<unsigned long long> ekey;
    std::vector<unsigned> vert_index(nrows, 0xffffffffu);
    long long nvert = 0;
    for (long long r = 0; r < nrows; ++r) {
        if (rtype[r] == 2) ekey.push_back(((unsigned long long)minmort[r] << 32) | (unsigned long long)r);
        else if (rtype[r] == 1) vert_index[r] = (unsigned)nvert++;
    }
    std::sort(ekey.begin(), ekey.end());
    const long long nedges = (long long)ekey.size();
    if (nedges == 0) { out.why = "no edge rows"; return 0; }
    const int EC = std::max(32, in.edges_per_cluster);
    const long long ncl = (nedges + EC - 1) / EC;
    if (ncl >= (1LL << 30)) { out.why = "too many clusters"; return 0; }

    // layout of one cluster (deterministic; computed twice: sizes, then fill)
    auto layout = [&](long long c, ClusterLayout& L) {
        const long long e0 = c * EC, e1 = std::min(nedges, e0 + EC);
        std::vector<std::pair<int, unsigned>> byn;   // (-steps, row)
        for (long long k = e0; k < e1; ++k) {
            const unsigned r = (unsigned)(ekey[k] & 0xffffffffu);
            byn.push_back({-(int)(in.radj_ptr[r + 1] - in.radj_ptr[r] + 1), r});
        }
        std::stable_sort(byn.begin(), byn.end(), [](const std::pair<int, unsigned>& x, const std::pair<int, unsigned>& y) { return x.first < y.first; });
        const int ns = (int)((byn.size() + 31) / 32);
        L.edges.assign((size_t)ns * 32, 0xffffffffu);
        L.slice_steps.assign(ns, 0);
        for (size_t k = 0; k < byn.size(); ++k) {
            L.edges[k] = byn[k].second;
            L.slice_steps[k / 32] = std::max(L.slice_steps[k / 32], -byn[k].first);
        }
        L.elist.clear();
        L.nvent = 0;
        for (const auto& b : byn) {
            const unsigned r = b.second;
            const long long a0 = in.radj_ptr[r], a1 = in.radj_ptr[r + 1];
            for (long long a = a0; a < a1; ++a) L.elist.push_back(in.old2new[in.radj[a] / 10u]);
            L.nvent += 4 + (int)(a1 - a0) + (closed[r] ? 0 : 1);   // (a,b), (a,ab), (b,a), (b,ab) and one (ring vertex, ab) per ring vertex
        }
        std::sort(L.elist.begin(), L.elist.end());
        L.elist.erase(std::unique(L.elist.begin(), L.elist.end()), L.elist.end());
        int off = 0;
        L.ebase.assign(L.edges.size(), 0);
        for (size_t k = 0; k < L.edges.size(); ++k) {
            L.ebase[k] = off;
            if (L.edges[k] != 0xffffffffu) off += (int)(in.rowptr[L.edges[k] + 1] - in.rowptr[L.edges[k]]);
        }
        L.vimg0 = off;
        L.image = off + L.nvent;
    };

    lap("edge sort");
    // ---- pass 2a: sizes
    std::vector<int> c_slices(ncl), c_elist(ncl), c_desc(ncl), c_image(ncl), c_vent(ncl);
    std::vector<long long> c_steps(ncl);
    parallel_for(ncl, nth, [&](int, long long c0, long long c1) {
        ClusterLayout L;
        for (long long c = c0; c < c1; ++c) {
            layout(c, L);
            c_slices[c] = (int)L.slice_steps.size();
            long long st = 0;
            for (int s : L.slice_steps) st += s;
            c_steps[c] = st;
            c_elist[c] = (int)L.elist.size();
            int ne = 0;
            for (unsigned r : L.edges) ne += r != 0xffffffffu;
            c_desc[c] = ne;
            c_image[c] = L.image;
            c_vent[c] = L.nvent;
        }
    });
    int gcap = 0, imgcap = 0;
    for (long long c = 0; c < ncl; ++c) {
        gcap = std::max(gcap, c_elist[c]); imgcap = std::max(imgcap, c_image[c]);
        out.stepcap = std::max(out.stepcap, (int)c_steps[c]); out.xcap = std::max(out.xcap, c_vent[c]);
    }
    out.gcap = gcap; out.imgcap = imgcap; out.edges_per_cluster = EC;
    out.smem_bytes = ring_smem_bytes(gcap, imgcap, out.stepcap, EC, out.xcap, 4);
    if (gcap >= (1 << RW0_EL_BITS) - 1) { out.why = "cluster stages too many elements"; return 0; }
    if (imgcap >= 65535 || out.smem_bytes > (size_t)in.max_smem_bytes) { out.why = "cluster does not fit the shared-memory budget"; return 0; }

    out.cs.assign(ncl + 1, 0); out.eptr.assign(ncl + 1, 0); out.dptr.assign(ncl + 1, 0); out.xptr.assign(ncl + 1, 0);
    std::vector<long long> cstep0(ncl + 1, 0);
    for (long long c = 0; c < ncl; ++c) {
        out.cs[c + 1] = out.cs[c] + c_slices[c];
        out.eptr[c + 1] = out.eptr[c] + ((c_elist[c] + 3) & ~3);   // 16-byte aligned ranges; the count is in the cluster record
        out.dptr[c + 1] = out.dptr[c] + c_desc[c];
        out.xptr[c + 1] = out.xptr[c] + ((c_vent[c] + 3) & ~3);   // 16-byte aligned ranges (asynchronous copies); the count is vimg[2c+1] - vimg[2c]
        cstep0[c + 1] = cstep0[c] + c_steps[c];
    }
    lap("cluster sizes");
    const long long nslices = out.cs[ncl], nsteps = cstep0[ncl];
    if (nslices * 32 * 4 >= 0xfffffff0LL) { out.why = "too many edges for 32-bit scratch indices"; return 0; }
    out.elist.assign((size_t)out.eptr[ncl], 0);
    out.desc.assign((size_t)out.dptr[ncl], RingRowDesc{0, 0, 0, 0, 0});
    out.sptr.assign(nslices + 1, 0);
    out.hdr.assign((size_t)nslices * RING_HW * 32, 0);
    out.steps.assign((size_t)nsteps * RING_SW * 32, 0);
    out.vimg.assign((size_t)2 * ncl, 0);
    out.xpos.assign((size_t)out.xptr[ncl], 0);
    out.xbase.assign(ncl, 0);
    out.cl_maxrow.assign(ncl, 0);
    out.cinfo.assign(ncl, RingCluster());
    std::vector<unsigned> lane_va((size_t)nslices * 32, 0xffffffffu), lane_vb((size_t)nslices * 32, 0xffffffffu);

    lap("allocate");
    std::vector<unsigned char> row_holes;   // rows with entries no local element contributes to (superset patterns)
    // ---- entries of the pattern no element contributes to
    {
        std::vector<std::vector<long long>> zl(nth);
        row_holes.assign(nrows, 0);
        parallel_for(nrows, nth, [&](int t, long long r0, long long r1) {
            for (long long r = r0; r < r1; ++r) {
                const int len = (int)(in.rowptr[r + 1] - in.rowptr[r]);
                if (len == 0) continue;
                unsigned long long m[4] = {0, 0, 0, 0};
                for (long long a = in.radj_ptr[r]; a < in.radj_ptr[r + 1]; ++a)
                    for (int j = 0; j < 10; ++j) { const int sl = in.pos[(size_t)a * 10 + j]; m[sl >> 6] |= 1ULL << (sl & 63); }
                for (int sl = 0; sl < len; ++sl)
                    if (!((m[sl >> 6] >> (sl & 63)) & 1ULL)) { zl[t].push_back(in.rowptr[r] + sl); row_holes[r] = 1; }
            }
        });
        for (auto& z : zl) out.zlist.insert(out.zlist.end(), z.begin(), z.end());
    }
    lap("unproduced entries");
    // ---- pass 2b: fill
    parallel_for(ncl, nth, [&](int, long long c0, long long c1) {
        ClusterLayout L;
        for (long long c = c0; c < c1; ++c) {
            layout(c, L);
            std::copy(L.elist.begin(), L.elist.end(), out.elist.begin() + out.eptr[c]);
            out.vimg[2 * c] = L.vimg0; out.vimg[2 * c + 1] = L.image;
            {
                RingCluster& ci = out.cinfo[c];
                std::memset(&ci, 0, sizeof(ci));
                ci.e0 = out.eptr[c]; ci.ne = (int)L.elist.size();
                ci.sl0 = out.cs[c]; ci.nsl = (int)L.slice_steps.size();
                ci.stc0 = cstep0[c]; ci.nstc = (int)c_steps[c];
                ci.d0 = out.dptr[c]; ci.nd = out.dptr[c + 1] - out.dptr[c];
                ci.x0 = out.xptr[c]; ci.vim0 = L.vimg0; ci.nx = L.nvent;
            }
            unsigned maxrow = 0;
            // row images
            int d = out.dptr[c];
            for (size_t k = 0; k < L.edges.size(); ++k) {
                const unsigned r = L.edges[k];
                if (r == 0xffffffffu) continue;
                out.desc[d++] = RingRowDesc{in.rowptr[r], (unsigned short)L.ebase[k], (unsigned short)(in.rowptr[r + 1] - in.rowptr[r]), 0, 0};
                maxrow = std::max(maxrow, r);
                if (row_holes[r]) out.cinfo[c].pad = 1;   // the image of this cluster's edge rows must be zeroed first
            }
            // vertex-row entries this cluster produces, in emission order (per lane: the four ring sums of the edge's end points, then one
            // (ring vertex, ab) entry per step); they are stored compactly behind the edge rows, sorted by (row, slot)
            std::vector<unsigned long long> vk;
            vk.reserve(L.nvent);
            for (size_t k = 0; k < L.edges.size(); ++k) {
                const unsigned r = L.edges[k];
                if (r == 0xffffffffu) continue;
                const long long a0 = in.radj_ptr[r];
                const int n = (int)(in.radj_ptr[r + 1] - a0);
                auto key_of = [&](long long e, int lrow, int lcol) -> unsigned long long {
                    const unsigned row = (unsigned)(in.e2r[(long long)lrow * ntet + e] - 1);
                    const long long A = find_adj(in, row, (unsigned)(e * 10 + lrow));
                    if (A < 0) { fail = F_NOADJ; return 0; }
                    return ((unsigned long long)row << 8) | in.pos[(size_t)A * 10 + lcol];
                };
                {
                    int k0; Frame f;
                    unpack_vis(ringvis[a0], &k0, &f);
                    const unsigned t = in.radj[a0 + k0];
                    const long long e = t / 10u;
                    const int i = (int)(t % 10u);
                    vk.push_back(key_of(e, f.la, f.lb)); vk.push_back(key_of(e, f.la, i));
                    vk.push_back(key_of(e, f.lb, f.la)); vk.push_back(key_of(e, f.lb, i));
                }
                for (int st = closed[r] ? 1 : 0; st <= n; ++st) {   // step 0 of a closed ring stores nothing (its group ends at the terminal step)
                    int kk; Frame f;
                    unpack_vis(ringvis[a0 + (st == n ? n - 1 : st)], &kk, &f);
                    const unsigned t = in.radj[a0 + kk];
                    vk.push_back(key_of(t / 10u, st == n ? f.ls : f.lr, (int)(t % 10u)));
                }
            }
            if (fail.load() != F_OK) return;
            std::vector<unsigned long long> vs(vk);
            std::sort(vs.begin(), vs.end());
            if (std::adjacent_find(vs.begin(), vs.end()) != vs.end()) { fail = F_DUP; return; }
            if (!vs.empty()) {
                long long lo = in.rowptr[vs.front() >> 8] + (long long)(vs.front() & 0xff), hi = in.rowptr[vs.back() >> 8] + (long long)(vs.back() & 0xff);
                if (hi - lo >= 0xffffffffLL) { fail = F_RANGE; return; }
                out.xbase[c] = lo;
                out.cinfo[c].xbase = lo;
                for (size_t k = 0; k < vs.size(); ++k) {
                    out.xpos[(size_t)out.xptr[c] + k] = (unsigned)(in.rowptr[vs[k] >> 8] + (long long)(vs[k] & 0xff) - lo);
                    maxrow = std::max(maxrow, (unsigned)(vs[k] >> 8));
                }
            }
            out.cl_maxrow[c] = maxrow;
            size_t vnext = 0;   // walks vk in emission order
            auto next_vimg = [&]() -> unsigned {
                const unsigned long long key = vk[vnext++];
                return (unsigned)(L.vimg0 + (int)(std::lower_bound(vs.begin(), vs.end(), key) - vs.begin()));
            };
            auto eloc_of = [&](long long e) -> unsigned {
                const unsigned m = in.old2new[e];
                return (unsigned)(std::lower_bound(L.elist.begin(), L.elist.end(), m) - L.elist.begin()) + 1u;
            };
            long long step0 = cstep0[c];
            for (size_t s = 0; s < L.slice_steps.size(); ++s) {
                const long long sg = out.cs[c] + (long long)s;
                out.sptr[sg] = step0;
                const int nst = L.slice_steps[s];
                for (int lane = 0; lane < 32; ++lane) {
                    const unsigned r = L.edges[s * 32 + lane];
                    unsigned* H = out.hdr.data() + (size_t)sg * RING_HW * 32 + lane;
                    H[5 * 32] = (unsigned)(step0 - cstep0[c]) | ((unsigned)nst << 16);
                    if (r == 0xffffffffu) { H[4 * 32] = 0xffffffffu; continue; }
                    const long long a0 = in.radj_ptr[r];
                    const int n = (int)(in.radj_ptr[r + 1] - a0);
                    const bool cyc = closed[r] != 0;
                    const int ebase = L.ebase[s * 32 + lane];
                    // static targets from the first ring tet
                    {
                        int k0; Frame f;
                        unpack_vis(ringvis[a0], &k0, &f);
                        const long long A = a0 + k0;
                        const unsigned t = in.radj[A];
                        const long long e = t / 10u;
                        const int i = (int)(t % 10u);
                        const unsigned char* pa = in.pos + (size_t)A * 10;
                        const unsigned rowa = (unsigned)(in.e2r[(long long)f.la * ntet + e] - 1), rowb = (unsigned)(in.e2r[(long long)f.lb * ntet + e] - 1);
                        H[0] = (unsigned)ebase | ((unsigned)(n + 1) << 16) | (1u << 24);
                        H[1 * 32] = (unsigned)pa[f.la] | ((unsigned)pa[f.lb] << 8) | ((unsigned)pa[i] << 16);
                        const unsigned o_ab = next_vimg(), o_a4 = next_vimg(), o_ba = next_vimg(), o_b4 = next_vimg();
                        H[2 * 32] = o_ab | (o_a4 << 16);
                        H[3 * 32] = o_ba | (o_b4 << 16);
                        H[4 * 32] = r;
                        lane_va[(size_t)sg * 32 + lane] = rowa;
                        lane_vb[(size_t)sg * 32 + lane] = rowb;
                    }
                    for (int st = 0; st <= n; ++st) {
                        unsigned* W = out.steps.data() + (size_t)(step0 + st) * RING_SW * 32 + lane;
                        const bool term = st == n;
                        int kk; Frame f;
                        unpack_vis(ringvis[a0 + (term ? n - 1 : st)], &kk, &f);
                        const long long A = a0 + kk;
                        const unsigned t = in.radj[A];
                        const long long e = t / 10u;
                        const unsigned char* pa = in.pos + (size_t)A * 10;
                        // the ring vertex whose group ends here: r of this tet, or (terminal step) s of the last tet.  A closed ring
                        // ends with the group of its first ring vertex (kept since step 0)
                        int lg = term ? f.ls : f.lr;
                        unsigned w0 = 0;
                        if (!term) {
                            int pc, hf;
                            w0 = eloc_of(e);
                            ring_pair_piece(f.la, f.lb, &pc, &hf); w0 |= (unsigned)pc << RW0_TAU_SHIFT; w0 |= (unsigned)hf << RW0_SWAP_SHIFT;
                            ring_pair_piece(f.la, f.lr, &pc, &hf); w0 |= (unsigned)pc << (RW0_TAU_SHIFT + 2); w0 |= (unsigned)hf << (RW0_SWAP_SHIFT + 1);
                            ring_pair_piece(f.la, f.ls, &pc, &hf); w0 |= (unsigned)pc << (RW0_TAU_SHIFT + 4); w0 |= (unsigned)hf << (RW0_SWAP_SHIFT + 2);
                            if (f.lb < f.lr && f.lb < f.ls) w0 |= RW0_FLAGA;
                            if (f.la < f.lr && f.la < f.ls) w0 |= RW0_FLAGB;
                        }
                        if (st == 0 && cyc) w0 |= RW0_HOLDF; else w0 |= RW0_EMITR;
                        if (term && cyc) w0 |= RW0_ADDF;
                        W[0] = w0;
                        W[1 * 32] = (unsigned)pa[lg] | ((unsigned)pa[ring_eidx(f.la, lg)] << 8) | ((unsigned)pa[ring_eidx(f.lb, lg)] << 16) |
                                    (term ? 0u : ((unsigned)pa[ring_eidx(f.lr, f.ls)] << 24));
                        W[2 * 32] = (st == 0 && cyc) ? 0u : next_vimg();
                    }
                    // steps beyond the ring (padding to the slice maximum) stay zero: no element, nothing emitted
                }
                step0 += nst;
            }
        }
    });
    if (fail.load() != F_OK) { out.why = fail_text(fail.load()); return 0; }
    out.sptr[nslices] = nsteps;

    lap("fill");
    // ---- vertex diagonals / loads: partial sums per vertex row in (slice, lane, endpoint) order
    out.vptr.assign(nvert + 1, 0);
    out.vrow.assign(nvert, 0);
    out.vdpos.assign(nvert, 0);
    for (long long r = 0; r < nrows; ++r)
        if (vert_index[r] != 0xffffffffu) {
            const unsigned vi = vert_index[r];
            out.vrow[vi] = (unsigned)r;
            const long long a0 = in.radj_ptr[r];
            out.vdpos[vi] = in.rowptr[r] + in.pos[(size_t)a0 * 10 + (in.radj[a0] % 10u)];
        }
    for (size_t k = 0; k < lane_va.size(); ++k) {
        if (lane_va[k] != 0xffffffffu) out.vptr[vert_index[lane_va[k]] + 1]++;
        if (lane_vb[k] != 0xffffffffu) out.vptr[vert_index[lane_vb[k]] + 1]++;
    }
    for (long long vtx = 0; vtx < nvert; ++vtx) out.vptr[vtx + 1] += out.vptr[vtx];
    out.vlist.assign((size_t)out.vptr[nvert], 0);
    {
        std::vector<long long> fill(out.vptr.begin(), out.vptr.end() - 1);
        for (size_t k = 0; k < lane_va.size(); ++k) {
            if (lane_va[k] != 0xffffffffu) out.vlist[(size_t)fill[vert_index[lane_va[k]]]++] = (unsigned)(k * 4);
            if (lane_vb[k] != 0xffffffffu) out.vlist[(size_t)fill[vert_index[lane_vb[k]]]++] = (unsigned)(k * 4 + 1);
        }
    }

    lap("vertex lists");
    out.ncl = ncl; out.nslices = nslices; out.nsteps = nsteps; out.nedges = nedges; out.nvert = nvert; out.nstaged = out.eptr[ncl];
    out.ok = true;
    return 0;
}

}  // namespace afb
