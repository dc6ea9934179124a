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
#include <algorithm>
#include <cstring>

#include "afb_internal.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}
#define GRID_STRIDE(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// scalar codes of one field: dst[i*ntet + e] = src[(loff+i)*ntet + e] - goff   (codes stay id+1)
__global__ void k_extract_codes(long long ntet, int nloc, int loff, int goff, const int32_t* __restrict__ src, int32_t* dst) {
    GRID_STRIDE(t, ntet * nloc) {
        const long long i = t / ntet, e = t - i * ntet;
        dst[t] = src[(long long)(loff + i) * ntet + e] - goff;
    }
}

struct LenSrc { const long long* rowptr[3]; int mult[3]; };

// every row of field R must be the concatenation of its pair-pattern rows (mult = number of fields per base space)
__global__ void k_check_rows(long long nrows, long long goff, const long long* __restrict__ grow, LenSrc ls, int* bad) {
    GRID_STRIDE(r, nrows) {
        long long len = 0;
        for (int k = 0; k < 3; ++k)
            if (ls.rowptr[k]) len += ls.mult[k] * (ls.rowptr[k][r + 1] - ls.rowptr[k][r]);
        if (len != grow[goff + r + 1] - grow[goff + r]) *bad = 1;
    }
}

// first CSR entry of the block (R, C) of every (slice, lane) row of the pair plan: row start + the blocks of the fields before C
__global__ void k_block_dst(long long n, const unsigned* __restrict__ srow, long long goff, const long long* __restrict__ grow, LenSrc before,
                            long long* dst) {
    GRID_STRIDE(t, n) {
        const unsigned r = srow[t];
        long long p0 = 0;
        if (r != 0xffffffffu) {
            p0 = grow[goff + r];
            for (int k = 0; k < 3; ++k)
                if (before.rowptr[k]) p0 += before.mult[k] * (before.rowptr[k][r + 1] - before.rowptr[k][r]);
        }
        dst[t] = p0;
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

void blocks_clear(afb_ctx* ctx) {
    for (auto& p : ctx->pairs) {
        if (p.sub) afb_ctx_destroy(p.sub);
    }
    ctx->pairs.clear();
    for (auto& d : ctx->block_dst) d.release();
    ctx->block_dst.clear();
    ctx->blocks_ready = false;
}

// Builds the pair plans and block destinations after the main pattern exists.  Never fails the caller: when something
// does not fit the scheme the block path is simply not offered (the generic staged path assembles the problem).
int blocks_build(afb_ctx* ctx) {
    blocks_clear(ctx);
    if (ctx->is_sub || getenv("AFB_DISABLE_BLOCKS")) return 0;
    const int nf = (int)ctx->fields.size();
    if (nf < 2 || nf > 8 || ctx->has_signs || ctx->has_diag || ctx->nrow_loc != ctx->ncol_loc) return 0;
    std::vector<int> fems;
    for (const Field& f : ctx->fields) {
        if (f.fem != AFB_FEM_P1 && f.fem != AFB_FEM_P2) return 0;
        if (std::find(fems.begin(), fems.end(), f.fem) == fems.end()) fems.push_back(f.fem);
    }
    cudaStream_t st = ctx->stream;
    const long long ntet = ctx->ntet;
    // first field of every base space (all fields of a space share the scalar numbering)
    auto first_field = [&](int fem) -> const Field& {
        for (const Field& f : ctx->fields)
            if (f.fem == fem) return f;
        return ctx->fields[0];
    };
    for (int femR : fems)
        for (int femC : fems) {
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
            k_extract_codes<<<grid_for(ntet * fr.nloc), 256, 0, st>>>(ntet, fr.nloc, fr.loff, (int)fr.goff, ctx->e2r.as<int32_t>(), sub->e2r.as<int32_t>());
            k_extract_codes<<<grid_for(ntet * fc.nloc), 256, 0, st>>>(ntet, fc.nloc, fc.loff, (int)fc.goff, ctx->e2c.as<int32_t>(), sub->e2c.as<int32_t>());
            ctx->launches += 2;
            sub->row_begin = 0; sub->row_end = fr.count; sub->ncols_global = fc.count;
            sub->has_dofmap = true; sub->has_signs = false;
            if (femR != femC) {
                // no forced diagonal in a rectangular block
                if (sub->diag_col.reserve(std::max<long long>(1, fr.count) * 4) != cudaSuccess) { blocks_clear(ctx); return 0; }
                cudaMemsetAsync(sub->diag_col.p, 0xFF, fr.count * 4, st);
                sub->has_diag = true;
            }
            int64_t nnz = 0;
            const int rc = afb_pattern_build(sub, &nnz);
            ctx->launches += sub->launches; sub->launches = 0;
            if (rc != 0 || !sub->has_rows_plan) { blocks_clear(ctx); return 0; }
        }
    // ---- consistency of the row structure + destinations of every field pair
    if (ctx->flag.reserve(64) != cudaSuccess) { blocks_clear(ctx); return 0; }
    cudaMemsetAsync(ctx->flag.p, 0, 64, st);
    const long long* grow = ctx->rowptr.as<long long>();
    auto len_src = [&](int femR, int upto_field) {
        LenSrc ls;
        for (int k = 0; k < 3; ++k) { ls.rowptr[k] = nullptr; ls.mult[k] = 0; }
        for (size_t k = 0; k < fems.size() && k < 3; ++k) {
            int m = 0;
            for (int g = 0; g < upto_field; ++g) m += ctx->fields[g].fem == fems[k];
            if (m) { ls.rowptr[k] = find_pair(ctx, femR, fems[k])->rowptr.as<long long>(); ls.mult[k] = m; }
        }
        return ls;
    };
    if (fems.size() > 3) { blocks_clear(ctx); return 0; }
    for (int fR = 0; fR < nf; ++fR) {
        const Field& f = ctx->fields[fR];
        k_check_rows<<<grid_for(f.count), 256, 0, st>>>(f.count, f.goff, grow, len_src(f.fem, nf), ctx->flag.as<int>());
    }
    ctx->block_dst.resize((size_t)nf * nf);
    for (int fR = 0; fR < nf; ++fR)
        for (int fC = 0; fC < nf; ++fC) {
            const Field& f = ctx->fields[fR];
            afb_ctx* sub = find_pair(ctx, f.fem, ctx->fields[fC].fem);
            const long long n = sub->rp_nslices * 32;
            afb::DevBuf& d = ctx->block_dst[(size_t)fR * nf + fC];
            if (d.reserve(std::max<long long>(1, n) * 8) != cudaSuccess) { blocks_clear(ctx); return 0; }
            k_block_dst<<<grid_for(n), 256, 0, st>>>(n, sub->rp_order.as<unsigned>(), f.goff, grow, len_src(f.fem, fC), d.as<long long>());
        }
    ctx->launches += nf + nf * nf;
    int bad = 0;
    if (cudaMemcpyAsync(&bad, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess || bad) {
        cudaGetLastError();
        blocks_clear(ctx);
        return 0;
    }
    ctx->blocks_ready = true;
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
        if (drhs && std::find(row_has_rhs.begin(), row_has_rhs.end(), 0) != row_has_rhs.end() && nrows)
            AFB_CUDA(ctx, cudaMemsetAsync(drhs, 0, nrows * sizeof(double), st));
    }
    for (int fR = 0; fR < nf; ++fR)
        for (int fC = 0; fC < nf; ++fC) {
            Group& g = groups[(size_t)fR * nf + fC];
            if (g.mat.empty() && g.rhs.empty()) continue;
            afb_ctx* sub = find_pair(ctx, ctx->fields[fR].fem, ctx->fields[fC].fem);
            const int rc = fused_group(ctx, sub, g.mat, g.rhs, g.mat.empty() ? nullptr : dval,
                                       g.rhs.empty() ? nullptr : drhs + ctx->fields[fR].goff,
                                       ctx->block_dst[(size_t)fR * nf + fC].as<long long>(), accumulate, drop_val, status_flag, false);
            if (rc < 0) return rc;
            if (rc != 2) { set_error(ctx, "internal: block path lost a group after the support check"); return -4; }
        }
    cudaEventRecord(ctx->ev[3], st);
    return 2;
}

}  // namespace afb
