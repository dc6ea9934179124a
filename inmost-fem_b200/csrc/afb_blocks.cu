// Block-decomposed fused assembly for vector-valued and mixed spaces (FemVec<3,P2> elasticity, Taylor-Hood Stokes, ...).
//
// Reference semantics: the band structure of vector operators (Operator<OP, FemVec<DIM,FEM>> as BandDenseMatrix<NPART>,
// fem/operators.h:127-155; DIV :320-353; composite spaces :189-259) makes the local matrix of a vector / mixed problem a grid
// of scalar blocks, and the NATURAL enumeration (inmost_interface/global_enumerator.cpp:702-777: VAR, DIM, ELEM_TYPE, ELEM_ID,
// DOF_ID) numbers every scalar field (variable, component) contiguously.  Hence
//   * a matrix row of field R is the concatenation, field by field, of the rows of the scalar "pair patterns"
//     (space of R) x (space of C), and
//   * every block (R, C) of every form is a scalar form of the tensor representation (afb_tensor.cu).
// So the vector / mixed assembly runs as one fused scalar assembly (k_geom + k_rows_cl) per field pair, on the gather plan of
// the pair of base spaces, writing straight into its sub-block of the caller's CSR rows.  The pair plans live in
// sub-contexts that borrow the mesh of the owning context.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "afb_internal.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}
#define GRID_STRIDE(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// scalar codes of one field: code c of local dof loff+i -> 1 + scalar id of (c - 1) inside the field's intervals (0 stays 0)
__global__ void k_extract_codes(long long ntet, int nloc, int loff, SegMap map, const int32_t* __restrict__ src, int32_t* dst, int* bad) {
    GRID_STRIDE(t, ntet * nloc) {
        const long long i = t / ntet, e = t - i * ntet;
        const int c = src[(long long)(loff + i) * ntet + e];
        int o = 0;
        if (c > 0) {
            const long long sid = map.to_scalar(c - 1);
            if (sid < 0) *bad = 1; else o = (int)sid + 1;
        } else if (c < 0) *bad = 1;
        dst[t] = o;
    }
}

__device__ __forceinline__ long long find_col(const int32_t* __restrict__ gcol, long long b0, long long b1, long long g) {
    long long lo = b0, hi = b1;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (gcol[mid] < g) lo = mid + 1; else hi = mid; }
    return (lo < b1 && gcol[lo] == g) ? lo : -1;
}

// Where the block (R, C) of every scalar row sits inside the caller's CSR row: p0 = first entry when the block is a
// contiguous run of the row (NATURAL numbering on one rank: the row is the concatenation of its field blocks), else
// p0 = row start and tcount = block length: the entries then go through an offset table (per-rank numbering of a partitioned
// mesh: the columns of a field are split by owner rank; columns contributed by other ranks only sit in between).
__global__ void k_block_scan(long long nrows_s, SegMap rowsR, SegMap colsC, const long long* __restrict__ srowptr, const int32_t* __restrict__ scol,
                             const long long* __restrict__ grow, const int32_t* __restrict__ gcol, long long* p0, int* tcount, int* bad,
                             unsigned long long* covered, int* rowcov) {
    unsigned long long mine = 0;
    GRID_STRIDE(r, nrows_s) {
        const long long R = rowsR.to_full(r);
        const long long a0 = srowptr[r], a1 = srowptr[r + 1];
        long long first = 0;
        int tc = 0;
        if (R < 0) { if (a1 > a0) *bad = 1; }
        else {
            const long long b0 = grow[R], b1 = grow[R + 1];
            first = b0;
            bool contig = true;
            long long prev = -1;
            for (long long a = a0; a < a1; ++a) {
                const long long pos = find_col(gcol, b0, b1, colsC.to_full(scol[a]));
                if (pos < 0) { *bad = 1; break; }
                if (a == a0) first = pos; else if (pos != prev + 1) contig = false;
                prev = pos;
            }
            if (!contig) { first = b0; tc = (int)(a1 - a0); if (b1 - b0 > 65535) *bad = 1; }
            mine += (unsigned long long)(a1 - a0);
            if (a1 > a0) atomicAdd(rowcov + R, (int)(a1 - a0));   // setup-time integer count
        }
        p0[r] = first;
        tcount[r] = tc;
    }
    if (mine) atomicAdd(covered, mine);   // setup-time integer count
}

__global__ void k_block_tab(long long nrows_s, SegMap rowsR, SegMap colsC, const long long* __restrict__ srowptr, const int32_t* __restrict__ scol,
                            const long long* __restrict__ grow, const int32_t* __restrict__ gcol, const int* __restrict__ tcount,
                            const int* __restrict__ toff, unsigned short* tab) {
    GRID_STRIDE(r, nrows_s) {
        if (!tcount[r]) continue;
        const long long R = rowsR.to_full(r);
        const long long a0 = srowptr[r], a1 = srowptr[r + 1], b0 = grow[R], b1 = grow[R + 1];
        for (long long a = a0; a < a1; ++a) tab[toff[r] + (a - a0)] = (unsigned short)(find_col(gcol, b0, b1, colsC.to_full(scol[a])) - b0);
    }
}

// rows with entries that belong to no block (columns only other ranks contribute to): flag for the compaction
__global__ void k_gap_flag(long long nrows, const long long* __restrict__ grow, const int* __restrict__ rowcov, int* flag) {
    GRID_STRIDE(r, nrows) flag[r] = rowcov[r] < (int)(grow[r + 1] - grow[r]) ? 1 : 0;
}
__global__ void k_gap_list(long long nrows, const int* __restrict__ flag, const int* __restrict__ off, int* list) {
    GRID_STRIDE(r, nrows) if (flag[r]) list[off[r]] = (int)r;
}
// one warp per listed row: clear its values
__global__ void k_zero_rows(int n, const int* __restrict__ list, const long long* __restrict__ grow, double* val) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = w; k < n; k += nw) {
        const long long r = list[k];
        for (long long a = grow[r] + lane; a < grow[r + 1]; a += 32) val[a] = 0.0;
    }
}

// per (slice, lane) of the pair plan: first entry of the block, table index (0 = contiguous)
__global__ void k_block_dst(long long n, const unsigned* __restrict__ srow, const long long* __restrict__ p0, const int* __restrict__ tcount,
                            const int* __restrict__ toff, long long* dst, int* tix) {
    GRID_STRIDE(t, n) {
        const unsigned r = srow[t];
        dst[t] = r != 0xffffffffu ? p0[r] : 0;
        tix[t] = (r != 0xffffffffu && tcount[r]) ? 1 + toff[r] : 0;
    }
}

// per (slice, lane): local row of the caller that receives the rhs entry of the scalar row
__global__ void k_block_rdst(long long n, const unsigned* __restrict__ srow, SegMap rowsR, int* rdst) {
    GRID_STRIDE(t, n) {
        const unsigned r = srow[t];
        rdst[t] = r != 0xffffffffu ? (int)rowsR.to_full(r) : -1;
    }
}

afb_ctx* find_pair(afb_ctx* ctx, int femR, int femC) {
    for (auto& p : ctx->pairs)
        if (p.femR == femR && p.femC == femC) return p.sub;
    return nullptr;
}

SForm blank_sform(int kind, int ng, int femA, int femB, int order, double alpha, const double* D, int layout, int dstride) {
    SForm s;
    std::memset(&s, 0, sizeof(s));
    s.kind = kind; s.ng = ng; s.layout = layout; s.dstride = dstride; s.alpha = alpha; s.D = D;
    for (int t = 0; t < 9; ++t) s.kidx[t] = -1;
    OpInfo a, b;
    resolve_op(AFB_IDEN, femA, 1, &a);
    resolve_op(AFB_IDEN, femB, 1, &b);
    s.femA = femA; s.femB = femB; s.nfa = a.nf_base; s.nfb = b.nf_base; s.quad_order = order;
    return s;
}

}  // namespace

namespace afb {

// drops the block destinations (they depend on the installed pattern); the pair plans depend on the dof map only and stay
void blocks_clear_dst(afb_ctx* ctx) {
    for (auto* v : {&ctx->block_dst, &ctx->block_tix, &ctx->block_tab, &ctx->block_rdst}) {
        for (auto& d : *v) d.release();
        v->clear();
    }
    ctx->block_gap.release();
    ctx->n_gap_rows = 0;
    ctx->blocks_ready = false;
    ctx->blocks_cover = true;
}

void blocks_clear(afb_ctx* ctx) {
    for (auto& p : ctx->pairs) {
        if (p.sub) afb_ctx_destroy(p.sub);
    }
    ctx->pairs.clear();
    blocks_clear_dst(ctx);
}

// Builds the pair plans and block destinations after the main pattern exists.  Never fails the caller: when something
// does not fit the scheme the block path is simply not offered (the generic staged path assembles the problem).
int blocks_build(afb_ctx* ctx) {
    blocks_clear_dst(ctx);
    const bool reuse_pairs = !ctx->pairs.empty();   // same dof map and fields, new pattern (afb_pattern_set after afb_pattern_build)
    if (ctx->is_sub || getenv("AFB_DISABLE_BLOCKS")) return 0;
    const int nf = (int)ctx->fields.size();
    if (nf < 2 || nf > 8 || ctx->has_signs || ctx->nrow_loc != ctx->ncol_loc || (ctx->has_diag && !ctx->fields_custom)) { blocks_clear(ctx); return 0; }
    std::vector<int> fems;
    for (const Field& f : ctx->fields) {
        if (f.fem != AFB_FEM_P1 && f.fem != AFB_FEM_P2) { blocks_clear(ctx); return 0; }
        if (std::find(fems.begin(), fems.end(), f.fem) == fems.end()) fems.push_back(f.fem);
    }
    if (fems.size() > 3) { blocks_clear(ctx); return 0; }
    cudaStream_t st = ctx->stream;
    const long long ntet = ctx->ntet;
    if (ctx->flag.reserve(64) != cudaSuccess) { blocks_clear(ctx); return 0; }
    cudaMemsetAsync(ctx->flag.p, 0, 64, st);
    int* bad = ctx->flag.as<int>();
    unsigned long long* covered = reinterpret_cast<unsigned long long*>(ctx->flag.as<char>() + 16);
    // first field of every base space (all fields of a space share the scalar numbering)
    auto first_field = [&](int fem) -> const Field& {
        for (const Field& f : ctx->fields)
            if (f.fem == fem) return f;
        return ctx->fields[0];
    };
    for (const Field& f : ctx->fields) {   // same interval structure for all fields of a space
        const Field& g = first_field(f.fem);
        bool same = f.rows.n == g.rows.n && f.cols.n == g.cols.n;
        for (int k = 0; same && k < f.rows.n; ++k) same = f.rows.count[k] == g.rows.count[k];
        for (int k = 0; same && k < f.cols.n; ++k) same = f.cols.count[k] == g.cols.count[k];
        if (!same) { blocks_clear(ctx); return 0; }
    }
    for (int femR : fems)
        for (int femC : fems) {
            if (reuse_pairs) break;
            const Field& fr = first_field(femR);
            const Field& fc = first_field(femC);
            afb_ctx* sub = new afb_ctx();
            sub->device = ctx->device; sub->stream = st; sub->own_stream = false; sub->is_sub = true;
            afb::DevBuf* src[7] = {&ctx->x, &ctx->y, &ctx->z, &ctx->v[0], &ctx->v[1], &ctx->v[2], &ctx->v[3]};
            afb::DevBuf* dst[7] = {&sub->x, &sub->y, &sub->z, &sub->v[0], &sub->v[1], &sub->v[2], &sub->v[3]};
            for (int k = 0; k < 7; ++k) { dst[k]->p = src[k]->p; dst[k]->cap = src[k]->cap; dst[k]->borrowed = true; }
            sub->nnode = ctx->nnode; sub->ntet = ntet;
            ctx->pairs.push_back({femR, femC, sub});
            sub->nrow_loc = fr.nloc; sub->ncol_loc = fc.nloc;
            if (sub->e2r.reserve((size_t)ntet * fr.nloc * 4) != cudaSuccess || sub->e2c.reserve((size_t)ntet * fc.nloc * 4) != cudaSuccess) { blocks_clear(ctx); return 0; }
            k_extract_codes<<<grid_for(ntet * fr.nloc), 256, 0, st>>>(ntet, fr.nloc, fr.loff, fr.rows, ctx->e2r.as<int32_t>(), sub->e2r.as<int32_t>(), bad);
            k_extract_codes<<<grid_for(ntet * fc.nloc), 256, 0, st>>>(ntet, fc.nloc, fc.loff, fc.cols, ctx->e2c.as<int32_t>(), sub->e2c.as<int32_t>(), bad);
            ctx->launches += 2;
            const long long nr = fr.rows.total();
            sub->row_begin = 0; sub->row_end = nr; sub->ncols_global = fc.cols.total();
            sub->has_dofmap = true; sub->has_signs = false;
            if (femR != femC || ctx->fields_custom) {
                // no forced diagonal in a rectangular block; under a segmented numbering the scalar row and column numberings
                // differ, and the diagonal entry is structural in the caller's pattern anyway
                if (sub->diag_col.reserve(std::max<long long>(1, nr) * 4) != cudaSuccess) { blocks_clear(ctx); return 0; }
                cudaMemsetAsync(sub->diag_col.p, 0xFF, nr * 4, st);
                sub->has_diag = true;
            }
            int64_t nnz = 0;
            const int rc = afb_pattern_build(sub, &nnz);
            ctx->launches += sub->launches; sub->launches = 0;
            if (rc != 0 || !sub->has_rows_plan) { blocks_clear(ctx); return 0; }
        }
    // ---- destinations of every field pair inside the caller's rows
    const long long* grow = ctx->rowptr.as<long long>();
    const int32_t* gcol = ctx->colind.as<int32_t>();
    ctx->block_dst.resize((size_t)nf * nf); ctx->block_tix.resize((size_t)nf * nf); ctx->block_tab.resize((size_t)nf * nf);
    ctx->block_rdst.resize((size_t)nf);
    afb::DevBuf p0, tcount, toff, cubtmp, rowcov, gflag, goff;
    const long long nrows_main = ctx->row_end - ctx->row_begin;
    if (rowcov.reserve(std::max<long long>(1, nrows_main) * 4) != cudaSuccess) { blocks_clear(ctx); return 0; }
    cudaMemsetAsync(rowcov.p, 0, std::max<long long>(1, nrows_main) * 4, st);
    auto fail = [&]() {
        if (getenv("AFB_VERBOSE")) fprintf(stderr, "[afb] block path not offered: block destinations could not be resolved (%s)\n", cudaGetErrorString(cudaGetLastError()));
        p0.release(); tcount.release(); toff.release(); cubtmp.release(); rowcov.release(); gflag.release(); goff.release(); cudaGetLastError(); blocks_clear(ctx); return 0;
    };
    for (int fR = 0; fR < nf; ++fR)
        for (int fC = 0; fC < nf; ++fC) {
            const Field& f = ctx->fields[fR];
            const Field& g = ctx->fields[fC];
            afb_ctx* sub = find_pair(ctx, f.fem, g.fem);
            const long long nrs = sub->row_end, n = sub->rp_nslices * 32;
            const size_t b = (size_t)fR * nf + fC;
            if (p0.reserve(std::max<long long>(1, nrs) * 8) != cudaSuccess || tcount.reserve((nrs + 1) * 4) != cudaSuccess || toff.reserve((nrs + 1) * 4) != cudaSuccess ||
                ctx->block_dst[b].reserve(std::max<long long>(1, n) * 8) != cudaSuccess || ctx->block_tix[b].reserve(std::max<long long>(1, n) * 4) != cudaSuccess)
                return fail();
            cudaMemsetAsync(tcount.p, 0, (nrs + 1) * 4, st);
            k_block_scan<<<grid_for(nrs), 256, 0, st>>>(nrs, f.rows, g.cols, sub->rowptr.as<long long>(), sub->colind.as<int32_t>(), grow, gcol, p0.as<long long>(),
                                                       tcount.as<int>(), bad, covered, rowcov.as<int>());
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, tcount.as<int>(), toff.as<int>(), nrs + 1, st);
            if (cubtmp.reserve(tb) != cudaSuccess) return fail();
            if (cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, tcount.as<int>(), toff.as<int>(), nrs + 1, st) != cudaSuccess) return fail();
            int ntab = 0, hb = 0;
            if (cudaMemcpyAsync(&ntab, toff.as<int>() + nrs, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess || hb || ntab < 0)
                return fail();
            if (ntab > 0) {
                if (ctx->block_tab[b].reserve((size_t)ntab * 2) != cudaSuccess) return fail();
                k_block_tab<<<grid_for(nrs), 256, 0, st>>>(nrs, f.rows, g.cols, sub->rowptr.as<long long>(), sub->colind.as<int32_t>(), grow, gcol, tcount.as<int>(),
                                                          toff.as<int>(), ctx->block_tab[b].as<unsigned short>());
            }
            k_block_dst<<<grid_for(n), 256, 0, st>>>(n, sub->rp_order.as<unsigned>(), p0.as<long long>(), tcount.as<int>(), toff.as<int>(),
                                                    ctx->block_dst[b].as<long long>(), ctx->block_tix[b].as<int>());
            if (ntab == 0) ctx->block_tix[b].release();   // all blocks contiguous: the gather skips the table lookups
            if (fC == fR) {
                if (ctx->block_rdst[fR].reserve(std::max<long long>(1, n) * 4) != cudaSuccess) return fail();
                k_block_rdst<<<grid_for(n), 256, 0, st>>>(n, sub->rp_order.as<unsigned>(), f.rows, ctx->block_rdst[fR].as<int>());
            }
            ctx->launches += 4;
        }
    int hb = 0;
    unsigned long long hcov = 0;
    if (cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(&hcov, covered, sizeof(hcov), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess || hb)
        return fail();
    ctx->blocks_cover = hcov == (unsigned long long)ctx->nnz;
    ctx->n_gap_rows = 0;
    if (!ctx->blocks_cover && nrows_main > 0) {
        // rows holding entries outside every block are cleared before a non-accumulating assembly (k_zero_rows)
        size_t tb = 0;
        if (gflag.reserve((nrows_main + 1) * 4) != cudaSuccess || goff.reserve((nrows_main + 1) * 4) != cudaSuccess) return fail();
        cudaMemsetAsync(gflag.p, 0, (nrows_main + 1) * 4, st);
        k_gap_flag<<<grid_for(nrows_main), 256, 0, st>>>(nrows_main, grow, rowcov.as<int>(), gflag.as<int>());
        cub::DeviceScan::ExclusiveSum(nullptr, tb, gflag.as<int>(), goff.as<int>(), nrows_main + 1, st);
        if (cubtmp.reserve(tb) != cudaSuccess) return fail();
        if (cub::DeviceScan::ExclusiveSum(cubtmp.p, tb, gflag.as<int>(), goff.as<int>(), nrows_main + 1, st) != cudaSuccess) return fail();
        int ngap = 0;
        if (cudaMemcpyAsync(&ngap, goff.as<int>() + nrows_main, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return fail();
        if (ctx->block_gap.reserve(std::max(1, ngap) * 4) != cudaSuccess) return fail();
        k_gap_list<<<grid_for(nrows_main), 256, 0, st>>>(nrows_main, gflag.as<int>(), goff.as<int>(), ctx->block_gap.as<int>());
        if (cudaStreamSynchronize(st) != cudaSuccess) return fail();
        ctx->n_gap_rows = ngap;
        ctx->launches += 2;
    }
    p0.release(); tcount.release(); toff.release(); cubtmp.release(); rowcov.release(); gflag.release(); goff.release();
    ctx->blocks_ready = true;
    if (getenv("AFB_VERBOSE"))
        fprintf(stderr, "[afb] block path: %d fields, %zu pair plans, blocks cover %llu of %lld entries (%d rows with uncovered entries)%s\n", nf,
                ctx->pairs.size(), hcov, (long long)ctx->nnz, ctx->n_gap_rows, ctx->fields_custom ? " (segmented numbering)" : "");
    return 0;
}

// Returns 2 when the problem was assembled block by block, 0 when the block path does not apply (nothing launched), < 0 on error.
int assemble_block_path(afb_ctx* ctx, int nfA, int nfF, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa,
                        const std::vector<OpInfo>& ob, const std::vector<const double*>& Dd, double* dval, double* drhs, int accumulate,
                        double drop_val, int* status_flag) {
    if (!ctx->blocks_ready || getenv("AFB_DISABLE_TENSOR_PATH")) return 0;
    const int nf = (int)ctx->fields.size();
    struct Group { std::vector<SForm> mat, rhs; };
    std::vector<Group> groups((size_t)nf * nf);
    auto field_at = [&](int loff, int fem) -> int {
        for (int g = 0; g < nf; ++g)
            if (ctx->fields[g].loff == loff && ctx->fields[g].fem == fem) return g;
        return -1;
    };
    for (int k = 0; k < nfA + nfF; ++k) {
        const afb_form& f = fm[k];
        const OpInfo &A = oa[k], &B = ob[k];
        const bool is_rhs = k >= nfA;
        if (f.coef_layout == AFB_COEF_PER_POINT) return 0;
        const int tt = f.tensor_type;
        const int dlen = form_dlen(f, A, B);
        const int nbB = B.nf_base, nbA = A.nf_base;
        if (is_rhs) {
            if (B.vec == 1) {
                const int fR = field_at(f.row_off, B.fem);
                SForm s;
                if (fR < 0 || !make_sform(f, A, B, Dd[k], &s)) return 0;
                s.row_off = s.col_off = 0;
                groups[(size_t)fR * nf + fR].rhs.push_back(s);
            } else if (B.op == AFB_IDEN && B.vec == 3 && tt >= AFB_TENSOR_SYMMETRIC && dlen == 3) {
                for (int a = 0; a < 3; ++a) {
                    const int fR = field_at(f.row_off + a * nbB, B.fem);
                    if (fR < 0) return 0;
                    SForm s = blank_sform(1, 1, AFB_FEM_P0, B.fem, f.quad_order, f.alpha, Dd[k], f.coef_layout, dlen);
                    s.kidx[0] = a;
                    groups[(size_t)fR * nf + fR].rhs.push_back(s);
                }
            } else return 0;
            continue;
        }
        if (A.vec == 1 && B.vec == 1) {
            const int fR = field_at(f.row_off, B.fem), fC = field_at(f.col_off, A.fem);
            SForm s;
            if (fR < 0 || fC < 0 || !make_sform(f, A, B, Dd[k], &s)) return 0;
            s.row_off = s.col_off = 0;
            groups[(size_t)fR * nf + fC].mat.push_back(s);
        } else if (A.vec == 3 && B.vec == 3 && A.op == B.op && (A.op == AFB_GRAD || A.op == AFB_IDEN)) {
            const bool grad = A.op == AFB_GRAD;
            const int dimc = grad ? 3 : 1;              // tensor rows/cols per component
            const int full_len = 9 * dimc * dimc;       // 81 or 9
            if (tt >= AFB_TENSOR_SYMMETRIC && dlen != full_len) return 0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    if (tt < AFB_TENSOR_SYMMETRIC && a != b) continue;
                    const int fR = field_at(f.row_off + a * nbB, B.fem), fC = field_at(f.col_off + b * nbA, A.fem);
                    if (fR < 0 || fC < 0) return 0;
                    SForm s = blank_sform(grad ? 0 : 1, grad ? (tt >= AFB_TENSOR_SYMMETRIC ? 9 : 6) : 1, A.fem, B.fem, f.quad_order, f.alpha, Dd[k],
                                          f.coef_layout, dlen);
                    if (tt < AFB_TENSOR_SYMMETRIC) s.kidx[0] = tt == AFB_TENSOR_SCALAR ? 0 : -2;   // c * identity
                    else if (grad) {
                        s.full = 1;
                        for (int kk = 0; kk < 3; ++kk)
                            for (int l = 0; l < 3; ++l) s.kidx[kk + 3 * l] = (3 * a + kk) + 9 * (3 * b + l);   // K(test 3a+k, trial 3b+l)
                    } else s.kidx[0] = a + 3 * b;
                    groups[(size_t)fR * nf + fC].mat.push_back(s);
                }
        } else if (A.op == AFB_DIV && A.vec == 3 && B.op == AFB_IDEN && B.vec == 1 && dlen <= 1) {
            // <c div u, q>: rows = q, columns = u_b: c * int phi^q_i d_b phi^u_j  ->  GRAD(A) x IDEN(B) with K(0,l) = c delta_lb
            const int fR = field_at(f.row_off, B.fem);
            for (int b = 0; b < 3; ++b) {
                const int fC = field_at(f.col_off + b * nbA, A.fem);
                if (fR < 0 || fC < 0) return 0;
                SForm s = blank_sform(2, 3, A.fem, B.fem, f.quad_order, f.alpha, Dd[k], f.coef_layout, dlen);
                s.kidx[b] = dlen ? 0 : -2;
                groups[(size_t)fR * nf + fC].mat.push_back(s);
            }
        } else if (A.op == AFB_IDEN && A.vec == 1 && B.op == AFB_DIV && B.vec == 3 && dlen <= 1) {
            // <c p, div v>: rows = v_a, columns = p: c * int d_a phi^v_i phi^p_j  ->  IDEN(A) x GRAD(B) with K(k,0) = c delta_ka
            const int fC = field_at(f.col_off, A.fem);
            for (int a = 0; a < 3; ++a) {
                const int fR = field_at(f.row_off + a * nbB, B.fem);
                if (fR < 0 || fC < 0) return 0;
                SForm s = blank_sform(3, 3, A.fem, B.fem, f.quad_order, f.alpha, Dd[k], f.coef_layout, dlen);
                s.kidx[a] = dlen ? 0 : -2;
                groups[(size_t)fR * nf + fC].mat.push_back(s);
            }
        } else return 0;
    }
    // ---- every group must be servable by the cluster gather of its pair plan
    bool empty_mat = false;
    std::vector<char> row_has_rhs(nf, 0);
    for (int fR = 0; fR < nf; ++fR)
        for (int fC = 0; fC < nf; ++fC) {
            Group& g = groups[(size_t)fR * nf + fC];
            if (g.mat.empty()) empty_mat = true;
            if (!g.rhs.empty()) row_has_rhs[fR] = 1;
            if (g.mat.empty() && g.rhs.empty()) continue;
            int nga = 0, ngf = 0;
            for (auto& s : g.mat) nga += s.ng;
            for (auto& s : g.rhs) ngf += s.ng;
            afb_ctx* sub = find_pair(ctx, ctx->fields[fR].fem, ctx->fields[fC].fem);
            if (!dval) { g.mat.clear(); nga = 0; }
            if (!drhs) { g.rhs.clear(); ngf = 0; }
            if (nga + ngf == 0) continue;
            if (!rows_supports(sub, nga, ngf)) return 0;
        }
    cudaStream_t st = ctx->stream;
    const long long nrows = ctx->row_end - ctx->row_begin;
    cudaEventRecord(ctx->ev[1], st);
    cudaEventRecord(ctx->ev[2], st);
    if (!accumulate) {
        // blocks no form touches are structural zeros of the template (assembler.inl:642-684)
        if (dval && empty_mat && ctx->nnz) AFB_CUDA(ctx, cudaMemsetAsync(dval, 0, ctx->nnz * sizeof(double), st));
        else if (dval && ctx->n_gap_rows > 0) {
            k_zero_rows<<<grid_for((long long)ctx->n_gap_rows * 32), 256, 0, st>>>(ctx->n_gap_rows, ctx->block_gap.as<int>(), ctx->rowptr.as<long long>(), dval);
            ctx->launches++;
        }
        if (drhs && std::find(row_has_rhs.begin(), row_has_rhs.end(), 0) != row_has_rhs.end() && nrows)
            AFB_CUDA(ctx, cudaMemsetAsync(drhs, 0, nrows * sizeof(double), st));
    }
    for (int fR = 0; fR < nf; ++fR)
        for (int fC = 0; fC < nf; ++fC) {
            Group& g = groups[(size_t)fR * nf + fC];
            if (g.mat.empty() && g.rhs.empty()) continue;
            afb_ctx* sub = find_pair(ctx, ctx->fields[fR].fem, ctx->fields[fC].fem);
            const size_t b = (size_t)fR * nf + fC;
            // one interval per field (NATURAL numbering on one rank): the load entries of field fR are a contiguous piece of the
            // rhs and the plain kernel serves the block unless it needs offset tables; else the rows go through the row map
            const bool mapped = ctx->fields[fR].rows.n != 1;
            double* rdst_base = g.rhs.empty() ? nullptr : (mapped ? drhs : drhs + ctx->fields[fR].rows.start[0]);
            const int rc = fused_group(ctx, sub, g.mat, g.rhs, g.mat.empty() ? nullptr : dval, rdst_base,
                                       ctx->block_dst[b].as<long long>(), accumulate, drop_val, status_flag, false, 0,
                                       ctx->block_tix[b].as<int>(), ctx->block_tab[b].as<unsigned short>(),
                                       (g.rhs.empty() || !mapped) ? nullptr : ctx->block_rdst[fR].as<int>());
            if (rc < 0) return rc;
            if (rc != 2) { set_error(ctx, "internal: block path lost a group after the support check"); return -4; }
        }
    cudaEventRecord(ctx->ev[3], st);
    return 2;
}

}  // namespace afb
