// CPU test of the ring plan (inmost-fem_b200/csrc/afb_ring_plan.cpp): builds P2 problems on tet meshes on the host, runs the
// plan builder and a scalar emulation of the k_rings kernel (same plan words, same order of operations) and compares the CSR
// values / load vector with the plain scatter of the same tensor-representation element matrices
//     val[row(e,i)][slot(e,i,j)] += sum_q TG[q][i][j] G_q(e)        (assembler.inl:397-481 on the local matrices of fem3Dtet).
// Meshes: Kuhn cubes with relabelled vertices and randomly permuted local vertex orders (every frame / closed and open rings).
// Test infrastructure only (no GPU, nothing here ships in the library).  Exit code 0 = pass.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

#include "../../inmost-fem_b200/csrc/afb_ring_plan.h"

namespace afb {   // host tables of the product (afb_tables.cpp)
void basis_values(int fem, int q, const double* XYL, double* phi);
void basis_ref_grads(int fem, int q, const double* XYL, double* G);
int tet_rule(int order, const double** p, const double** w);
}

using namespace afb;

static const int EV[10][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static const double SENT = -1.2345678901234567e+300;

struct Problem {
    long long ntet, nrows, nv;
    std::vector<int32_t> v[4];
    std::vector<int32_t> e2r;            // [10*ntet]
    std::vector<long long> rowptr, radj_ptr;
    std::vector<int32_t> colind;
    std::vector<unsigned> radj;
    std::vector<unsigned char> pos;
    std::vector<unsigned> old2new;
};

static void build_problem(int nx, int ny, int nz, unsigned seed, bool scramble, Problem& P, bool superset = false) {
    std::mt19937 rng(seed);
    const int NX = nx + 1, NY = ny + 1, NZ = nz + 1;
    P.nv = (long long)NX * NY * NZ;
    std::vector<int32_t> relabel(P.nv);
    for (long long k = 0; k < P.nv; ++k) relabel[k] = (int32_t)k;
    if (scramble) std::shuffle(relabel.begin(), relabel.end(), rng);
    auto node = [&](int i, int j, int k) { return relabel[((long long)i * NY + j) * NZ + k]; };
    for (int d = 0; d < 4; ++d) P.v[d].clear();
    int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k)
                for (int p = 0; p < 6; ++p) {
                    int c[3] = {i, j, k};
                    int32_t t[4];
                    t[0] = node(c[0], c[1], c[2]);
                    for (int s = 0; s < 3; ++s) { c[perms[p][s]] += 1; t[s + 1] = node(c[0], c[1], c[2]); }
                    if (scramble) std::shuffle(t, t + 4, rng);
                    for (int d = 0; d < 4; ++d) P.v[d].push_back(t[d]);
                }
    P.ntet = (long long)P.v[0].size();
    // P2 numbering: vertices, then edges in order of their sorted vertex pair
    std::map<std::pair<int32_t, int32_t>, int> eid;
    for (long long e = 0; e < P.ntet; ++e)
        for (int i = 4; i < 10; ++i) {
            int32_t a = P.v[EV[i][0]][e], b = P.v[EV[i][1]][e];
            if (a > b) std::swap(a, b);
            eid[{a, b}] = 0;
        }
    int ne = 0;
    for (auto& kv : eid) kv.second = ne++;
    P.nrows = P.nv + ne;
    P.e2r.assign((size_t)10 * P.ntet, 0);
    for (long long e = 0; e < P.ntet; ++e) {
        for (int i = 0; i < 4; ++i) P.e2r[(size_t)i * P.ntet + e] = P.v[i][e] + 1;
        for (int i = 4; i < 10; ++i) {
            int32_t a = P.v[EV[i][0]][e], b = P.v[EV[i][1]][e];
            if (a > b) std::swap(a, b);
            P.e2r[(size_t)i * P.ntet + e] = (int32_t)(P.nv + eid[{a, b}] + 1);
        }
    }
    // adjacency (ascending e*10+i per row), sorted pattern, slot table
    std::vector<std::vector<unsigned>> adj(P.nrows);
    for (long long e = 0; e < P.ntet; ++e)
        for (int i = 0; i < 10; ++i) adj[P.e2r[(size_t)i * P.ntet + e] - 1].push_back((unsigned)(e * 10 + i));
    P.radj_ptr.assign(P.nrows + 1, 0);
    P.rowptr.assign(P.nrows + 1, 0);
    P.radj.clear(); P.colind.clear();
    for (long long r = 0; r < P.nrows; ++r) {
        std::sort(adj[r].begin(), adj[r].end());
        std::vector<int32_t> cols;
        cols.push_back((int32_t)r);   // forced diagonal
        for (unsigned t : adj[r]) {
            const long long e = t / 10;
            for (int j = 0; j < 10; ++j) cols.push_back(P.e2r[(size_t)j * P.ntet + e] - 1);
            P.radj.push_back(t);
        }
        if (superset && r % 5 == 0) {   // columns no local element contributes to (what the union pattern of a partitioned mesh contains)
            cols.push_back((int32_t)((r * 7 + 3) % P.nrows)); cols.push_back((int32_t)((r * 13 + 1) % P.nrows)); cols.push_back((int32_t)(P.nrows - 1 - r % 3));
        }
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        P.colind.insert(P.colind.end(), cols.begin(), cols.end());
        P.radj_ptr[r + 1] = (long long)P.radj.size();
        P.rowptr[r + 1] = (long long)P.colind.size();
    }
    P.pos.assign(P.radj.size() * 10, 0);
    for (long long r = 0; r < P.nrows; ++r)
        for (long long a = P.radj_ptr[r]; a < P.radj_ptr[r + 1]; ++a) {
            const long long e = P.radj[a] / 10;
            for (int j = 0; j < 10; ++j) {
                const int32_t c = P.e2r[(size_t)j * P.ntet + e] - 1;
                const int32_t* b = P.colind.data() + P.rowptr[r];
                P.pos[(size_t)a * 10 + j] = (unsigned char)(std::lower_bound(b, (const int32_t*)(P.colind.data() + P.rowptr[r + 1]), c) - b);
            }
        }
    P.old2new.resize(P.ntet);
    for (long long e = 0; e < P.ntet; ++e) P.old2new[e] = (unsigned)e;
    if (!scramble) {   // Morton order of the hexes (the product sorts the Morton codes of the element centroids)
        auto spread = [](unsigned x) { unsigned r = 0; for (int b = 0; b < 10; ++b) r |= ((x >> b) & 1u) << (3 * b); return r; };
        std::vector<std::pair<unsigned long long, unsigned>> key(P.ntet);
        long long e = 0;
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < ny; ++j)
                for (int k = 0; k < nz; ++k)
                    for (int p = 0; p < 6; ++p, ++e) key[e] = {((unsigned long long)(spread(i) | (spread(j) << 1) | (spread(k) << 2)) << 3) | (unsigned)p, (unsigned)e};
        std::sort(key.begin(), key.end());
        for (long long m = 0; m < P.ntet; ++m) P.old2new[key[m].second] = (unsigned)m;
    }
    if (scramble) {   // any bijection works as "Morton" order; keep some locality: reverse blocks of 7
        for (long long e = 0; e + 7 <= P.ntet; e += 7) std::reverse(P.old2new.begin() + e, P.old2new.begin() + e + 7);
    }
}

// tables TM[c][i][j] like afb_tensor.cu::build_form_table (GRAD x GRAD, symmetric 6), mass table, load table
static void build_tables(int order, std::vector<double>& TG, std::vector<double>& Tm, std::vector<double>& Tf) {
    const double *pq, *wq;
    const int q = tet_rule(order, &pq, &wq);
    std::vector<double> phi((size_t)q * 10), G((size_t)q * 10 * 3);
    basis_values(3 /*P2*/, q, pq, phi.data());
    basis_ref_grads(3, q, pq, G.data());
    auto S = [&](int a, int b, int i, int j) { double s = 0; for (int n = 0; n < q; ++n) s += wq[n] * G[((size_t)n * 10 + i) * 3 + a] * G[((size_t)n * 10 + j) * 3 + b]; return s; };
    std::vector<double> TM(600);
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) {
            auto at = [&](int c) -> double& { return TM[((size_t)c * 10 + i) * 10 + j]; };
            at(0) = S(0, 0, i, j); at(1) = S(1, 1, i, j); at(2) = S(2, 2, i, j);
            at(3) = S(0, 1, i, j) + S(1, 0, i, j); at(4) = S(0, 2, i, j) + S(2, 0, i, j); at(5) = S(1, 2, i, j) + S(2, 1, i, j);
        }
    TG.assign(600, 0.0);
    ring_table_from_M(TM.data(), 10, TG.data());
    Tm.assign(100, 0.0); Tf.assign(10, 0.0);
    for (int i = 0; i < 10; ++i) {
        for (int j = 0; j < 10; ++j) { double s = 0; for (int n = 0; n < q; ++n) s += wq[n] * phi[(size_t)n * 10 + i] * phi[(size_t)n * 10 + j]; Tm[i * 10 + j] = s; }
        double s = 0; for (int n = 0; n < q; ++n) s += wq[n] * phi[(size_t)n * 10 + i]; Tf[i] = s;
    }
}

// scalar emulation of k_rings + k_ring_vertices on the plan (mirrors afb_rings.cu step by step)
static void emulate(const RingPlan& pl, const std::vector<double>& TG, const std::vector<double>& Tm, const std::vector<double>& Tf,
                    const std::vector<double>& gbuf /*[ntet*8] Morton order*/, std::vector<double>& val, std::vector<double>& rhs, bool accumulate, double drop) {
    auto T = [&](int q, int i, int j) { return TG[((size_t)q * 10 + i) * 10 + j]; };
    auto dropf = [&](double x) { return (std::fabs(x) <= drop) ? 0.0 : x; };
    std::vector<double> scratch((size_t)pl.nslices * 32 * 4, 0.0);
    if (!accumulate) for (long long z : pl.zlist) val[z] = 0.0;
    for (long long c = 0; c < pl.ncl; ++c) {
        std::vector<double> img(pl.vimg[2 * c + 1], SENT);
        if (pl.cinfo[c].pad & 1) for (int k = 0; k < pl.vimg[2 * c]; ++k) img[k] = 0.0;
        const unsigned* el = pl.elist.data() + pl.eptr[c];
        for (int s = pl.cs[c]; s < pl.cs[c + 1]; ++s) {
            const long long st0 = pl.sptr[s], st1 = pl.sptr[s + 1];
            for (int lane = 0; lane < 32; ++lane) {
                const unsigned* H = pl.hdr.data() + (size_t)s * RING_HW * 32 + lane;
                const unsigned h0 = H[0], h1 = H[32], h2 = H[64], h3 = H[96], h4 = H[128], h5 = H[160];
                if ((long long)(h5 & 0xffff) != st0 - pl.sptr[pl.cs[c]] || (long long)(h5 >> 16) != st1 - st0) { std::printf("slice header word 5 wrong\n"); std::exit(2); }
                if (pl.cinfo[c].sl0 != pl.cs[c] || pl.cinfo[c].e0 != pl.eptr[c] || pl.cinfo[c].stc0 != pl.sptr[pl.cs[c]] || pl.cinfo[c].x0 != pl.xptr[c] || pl.cinfo[c].vim0 != pl.vimg[2 * c]) { std::printf("cluster record wrong\n"); std::exit(2); }
                if (h4 == 0xffffffffu) continue;
                const int ebase = h0 & 0xffff;
                double Sa = 0, Sb = 0, Sab = 0, Vab = 0, Va4 = 0, Vba = 0, Vb4 = 0, Da = 0, Db = 0, Fab = 0, Fa = 0, Fb = 0;
                double C[4] = {0, 0, 0, 0}, Fg[4] = {0, 0, 0, 0};
                for (long long st = st0; st < st1; ++st) {
                    const unsigned* W = pl.steps.data() + (size_t)st * RING_SW * 32 + lane;
                    const unsigned w0 = W[0], w1 = W[32], w2 = W[64];
                    const unsigned eloc = w0 & ((1u << RW0_EL_BITS) - 1);
                    double rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    if (eloc) for (int k = 0; k < 8; ++k) rec[k] = gbuf[(size_t)el[eloc - 1] * 8 + k];
                    double G[6];
                    for (int p = 0; p < 3; ++p) {
                        const int tau = (w0 >> (RW0_TAU_SHIFT + 2 * p)) & 3, sw = (w0 >> (RW0_SWAP_SHIFT + p)) & 1;
                        G[2 * p] = rec[2 * tau + sw]; G[2 * p + 1] = rec[2 * tau + 1 - sw];
                    }
                    const double m = rec[6], f = rec[7];
                    auto E = [&](int i, int j) {
                        double x = 0;
                        for (int q = 0; q < 6; ++q) x = std::fma(T(q, i, j), G[q], x);
                        x = std::fma(Tm[i * 10 + j], m, x);
                        return dropf(x);
                    };
                    Sa += E(4, 0); Sb += E(4, 1); Sab += E(4, 4);
                    Vab += E(0, 1); Va4 += E(0, 4); Vba += E(1, 0); Vb4 += E(1, 4);
                    if (w0 & RW0_FLAGA) { Da += E(0, 0); Fa += Tf[0] * f; }
                    if (w0 & RW0_FLAGB) { Db += E(1, 1); Fb += Tf[1] * f; }
                    Fab += Tf[4] * f;
                    if (eloc) img[ebase + ((w1 >> 24) & 0xff)] = E(4, 9);
                    const double Rg[4] = {E(4, 2), E(4, 5), E(4, 7), E(2, 4)};
                    const double Sg[4] = {E(4, 3), E(4, 6), E(4, 8), E(3, 4)};
                    double o[4];
                    for (int k = 0; k < 4; ++k) o[k] = C[k] + Rg[k];
                    if (w0 & RW0_ADDF) for (int k = 0; k < 4; ++k) o[k] += Fg[k];
                    if (w0 & RW0_HOLDF) for (int k = 0; k < 4; ++k) Fg[k] = Rg[k];
                    if (w0 & RW0_EMITR) {
                        img[ebase + (w1 & 0xff)] = o[0]; img[ebase + ((w1 >> 8) & 0xff)] = o[1]; img[ebase + ((w1 >> 16) & 0xff)] = o[2];
                        img[w2 & 0xffff] = o[3];
                    }
                    for (int k = 0; k < 4; ++k) C[k] = Sg[k];
                }
                img[ebase + (h1 & 0xff)] = Sa; img[ebase + ((h1 >> 8) & 0xff)] = Sb; img[ebase + ((h1 >> 16) & 0xff)] = Sab;
                img[h2 & 0xffff] = Vab; img[h2 >> 16] = Va4; img[h3 & 0xffff] = Vba; img[h3 >> 16] = Vb4;
                double* sc = scratch.data() + ((size_t)s * 32 + lane) * 4;
                sc[0] = Da; sc[1] = Db; sc[2] = Fa; sc[3] = Fb;
                if (accumulate) rhs[h4] += Fab; else rhs[h4] = Fab;
            }
        }
        for (int d = pl.dptr[c]; d < pl.dptr[c + 1]; ++d) {
            const RingRowDesc& R = pl.desc[d];
            for (int k = 0; k < R.len; ++k) {
                const double x = img[R.off + k];
                if (x == SENT) { std::printf("edge-row image entry never written\n"); std::exit(2); }
                if (accumulate) val[R.p0 + k] += x; else val[R.p0 + k] = x;
            }
        }
        for (int k = pl.xptr[c]; k < pl.xptr[c] + (pl.vimg[2 * c + 1] - pl.vimg[2 * c]); ++k) {
            const double x = img[pl.vimg[2 * c] + (k - pl.xptr[c])];
            if (x == SENT) { std::printf("vertex-row entry never written\n"); std::exit(2); }
            const long long p = pl.xbase[c] + pl.xpos[k];
            if (accumulate) val[p] += x; else val[p] = x;
        }
    }
    for (long long vtx = 0; vtx < pl.nvert; ++vtx) {
        double d = 0, f = 0;
        for (long long k = pl.vptr[vtx]; k < pl.vptr[vtx + 1]; ++k) { d += scratch[pl.vlist[k]]; f += scratch[pl.vlist[k] + 2]; }
        if (accumulate) { val[pl.vdpos[vtx]] += d; rhs[pl.vrow[vtx]] += f; } else { val[pl.vdpos[vtx]] = d; rhs[pl.vrow[vtx]] = f; }
    }
}

static int run_case(int nx, int ny, int nz, unsigned seed, bool scramble, int EC, int nthreads, double drop, bool accumulate, bool superset = false) {
    Problem P;
    build_problem(nx, ny, nz, seed, scramble, P, superset);
    std::vector<double> TG, Tm, Tf;
    build_tables(2, TG, Tm, Tf);
    const double defect = ring_table_symmetry_defect(TG.data());
    if (!(defect < 1e-13)) { std::printf("FAILED: table not invariant under vertex relabelling (%.2e)\n", defect); return 1; }
    std::mt19937 rng(seed + 17);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> gnat((size_t)P.ntet * 8), gbuf((size_t)P.ntet * 8);
    for (auto& x : gnat) x = U(rng);
    if (drop > 0) for (size_t k = 0; k < gnat.size(); k += 5) gnat[k] *= 1e-9;   // some contributions below the drop threshold
    for (long long e = 0; e < P.ntet; ++e) std::memcpy(&gbuf[(size_t)P.old2new[e] * 8], &gnat[(size_t)e * 8], 64);
    // naive scatter
    const long long nnz = P.rowptr[P.nrows];
    std::vector<double> vref(nnz, accumulate ? 0.5 : 0.0), rref(P.nrows, accumulate ? 0.25 : 0.0);
    for (long long r = 0; r < P.nrows; ++r)
        for (long long a = P.radj_ptr[r]; a < P.radj_ptr[r + 1]; ++a) {
            const long long e = P.radj[a] / 10;
            const int i = (int)(P.radj[a] % 10);
            const double* g = &gnat[(size_t)e * 8];
            for (int j = 0; j < 10; ++j) {
                double x = 0;
                for (int q = 0; q < 6; ++q) x = std::fma(TG[((size_t)q * 10 + i) * 10 + j], g[q], x);
                x = std::fma(Tm[i * 10 + j], g[6], x);
                if (std::fabs(x) > drop) vref[P.rowptr[r] + P.pos[(size_t)a * 10 + j]] += x;
            }
            rref[r] += Tf[i] * g[7];
        }
    RingPlanIn in;
    in.ntet = P.ntet; in.nrows = P.nrows;
    for (int d = 0; d < 4; ++d) in.v[d] = P.v[d].data();
    in.e2r = P.e2r.data(); in.rowptr = P.rowptr.data(); in.radj_ptr = P.radj_ptr.data(); in.radj = P.radj.data(); in.pos = P.pos.data();
    in.old2new = P.old2new.data(); in.edges_per_cluster = EC; in.nthreads = nthreads;
    in.max_smem_bytes = 1 << 30;
    RingPlan pl;
    const auto t0 = std::chrono::steady_clock::now();
    ring_plan_build(in, pl);
    const double plan_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("RING_PLAN_VERBOSE")) std::printf("plan build %.1f ms on %d threads\n", plan_ms, nthreads);
    if (!pl.ok) { std::printf("FAILED: plan refused: %s\n", pl.why.c_str()); return 1; }
    std::vector<double> val(nnz, accumulate ? 0.5 : NAN), rhs(P.nrows, accumulate ? 0.25 : NAN);
    emulate(pl, TG, Tm, Tf, gbuf, val, rhs, accumulate, drop);
    double err = 0, errf = 0, scale = 0;
    for (long long k = 0; k < nnz; ++k) scale = std::max(scale, std::fabs(vref[k]));
    long long nanv = 0;
    for (long long k = 0; k < nnz; ++k) { if (val[k] != val[k]) ++nanv; else err = std::max(err, std::fabs(val[k] - vref[k])); }
    for (long long r = 0; r < P.nrows; ++r) { if (rhs[r] != rhs[r]) ++nanv; else errf = std::max(errf, std::fabs(rhs[r] - rref[r])); }
    long long tot_ring = 0;
    for (long long s = 0; s < pl.nslices; ++s) tot_ring += (pl.sptr[s + 1] - pl.sptr[s]) * 32;
    std::printf("cube %dx%dx%d%s EC %d: %lld tets, %lld rows, %lld edges in %lld clusters / %lld slices, %lld steps (%.0f%% lanes busy), staged x%.2f, "
                "image %d doubles, unwritten %lld, err %.2e rhs %.2e\n", nx, ny, nz, scramble ? " scrambled" : "", EC, P.ntet, P.nrows, pl.nedges, pl.ncl,
                pl.nslices, pl.nsteps, 100.0 * (7.0 * P.ntet /* 6 visits + ~1 terminal per edge share */) / std::max<long long>(1, tot_ring),
                (double)pl.nstaged / P.ntet, pl.imgcap, nanv, err / scale, errf);
    return (nanv == 0 && err <= 1e-13 * scale && errf <= 1e-13) ? 0 : 1;
}

int main(int argc, char** argv) {
    int fails = 0;
    if (argc >= 3) {   // statistics / timing of one cube: N, edges per cluster [, threads]
        const int n = std::atoi(argv[1]);
        return run_case(n, n, n, 7, false, std::atoi(argv[2]), argc > 3 ? std::atoi(argv[3]) : 8, 0.0, false);
    }
    fails += run_case(2, 2, 2, 1, false, 32, 1, 0.0, false);
    fails += run_case(3, 2, 2, 2, true, 32, 2, 0.0, false);
    fails += run_case(4, 3, 3, 3, true, 64, 3, 1e-6, false);
    fails += run_case(5, 4, 3, 4, true, 256, 4, 0.0, true);
    fails += run_case(1, 1, 1, 5, true, 32, 1, 0.0, false);
    fails += run_case(6, 6, 6, 6, false, 256, 4, 0.0, false);
    fails += run_case(5, 4, 4, 8, true, 64, 2, 0.0, false, true);    // superset pattern
    fails += run_case(4, 4, 3, 9, false, 128, 2, 0.0, true, true);   // superset pattern, accumulate
    std::printf(fails ? "test_ring_plan: %d FAILED\n" : "test_ring_plan: all passed\n", fails);
    return fails ? 1 : 0;
}
