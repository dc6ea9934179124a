// Surface terms of the global assembly: Neumann / Robin data on boundary faces.
//
// Reference semantics: the local assemblers of the examples call fem3Dface<OpA,OpB>(XYZ, face, D, A_face, ...) for every face of
// the cell that carries a boundary label and add A_face / F_face to the cell matrix before the essential conditions and the
// scatter (examples/Fem/Ani/diffusion.cpp:215-245, examples/Fem/Ani/lin_elast.cpp:198-217; fem/operations/int_face.inl:160-199).
// Here the caller hands over the list of such faces once (afb_boundary_set); afb_assemble_faces then
//   1. evaluates the face matrices of all listed faces in one batched launch per form (k_element_generic in face mode), staged
//      as [face][local row][local column];
//   2. adds them into the CSR values / rhs that afb_assemble produced: one thread per touched CSR row walks the row's
//      (face, local row) entries in ascending face order -- no atomics, deterministic -- and uses the slot table of the volume
//      plan (the face's cell is in the row's adjacency, afb_pattern.cu), the drop_val rule of assembler.inl:416 and, when
//      essential conditions are set (afb_dirichlet_set), the free-row part of applyDir (dc_on_dof.h:27-45): Dirichlet rows are
//      left alone (they are deg * identity rows), Dirichlet columns go to the rhs.
// The sum order differs from the reference's (cell matrix first, then scatter) by rounding only.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "afb_internal.h"

using namespace afb;

namespace {

inline unsigned grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}
#define GRID_STRIDE(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// item t = b*nrl + i  ->  key = local row (0xFFFFFFFF: skipped row)
__global__ void k_bf_items(long long nbf, int nrl, long long ntet, const int32_t* __restrict__ bf_tet, const int32_t* __restrict__ bf_face,
                           const int32_t* __restrict__ e2r, unsigned* key, unsigned* item, int* bad) {
    GRID_STRIDE(t, nbf * nrl) {
        const long long b = t / nrl;
        const int i = (int)(t - b * nrl);
        const int e = bf_tet[b];
        unsigned k = 0xffffffffu;
        if (e < 0 || e >= ntet || bf_face[b] < 0 || bf_face[b] > 3) *bad = 1;
        else {
            const int c = e2r[(long long)i * ntet + e];
            if (c > 0) k = (unsigned)(c - 1);
            else if (c < 0) *bad = 2;
        }
        key[t] = k;
        item[t] = (unsigned)t;
    }
}

// index of (cell, local row) inside the adjacency of its row (ascending e*nrl + i), i.e. the row of the slot table
__global__ void k_bf_adj(long long n, int nrl, const unsigned* __restrict__ key, const unsigned* __restrict__ item, const int32_t* __restrict__ bf_tet,
                         const long long* __restrict__ radj_ptr, const unsigned* __restrict__ radj, long long* aidx, int* bad) {
    GRID_STRIDE(t, n) {
        const unsigned r = key[t];
        const long long b = item[t] / (unsigned)nrl;
        const unsigned want = (unsigned)bf_tet[b] * (unsigned)nrl + item[t] % (unsigned)nrl;
        long long lo = radj_ptr[r], hi = radj_ptr[r + 1];
        const long long end = hi;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (radj[mid] < want) lo = mid + 1; else hi = mid; }
        if (lo >= end || radj[lo] != want) { *bad = 3; lo = 0; }
        aidx[t] = lo;
    }
}

template <typename PosT>
__global__ void k_bf_gather(int nu, const unsigned* __restrict__ urow, const int* __restrict__ uoff, const unsigned* __restrict__ item,
                            const long long* __restrict__ aidx, int nrl, int ncl, long long ntet, const int32_t* __restrict__ bf_tet,
                            const double* __restrict__ sA, const double* __restrict__ sF, const PosT* __restrict__ pos,
                            const long long* __restrict__ rowptr, const int32_t* __restrict__ e2c, long long row_begin,
                            const int32_t* __restrict__ diag_col, const int32_t* __restrict__ row_gid, const unsigned char* __restrict__ isdir,
                            const double* __restrict__ bc,
                            double* val, double* rhs, double drop, int* status) {
    GRID_STRIDE(u, nu) {
        const long long r = urow[u];
        if (isdir) {
            const long long gid = row_gid ? row_gid[r] : ((diag_col && diag_col[r] >= 0) ? diag_col[r] : row_begin + r);
            if (isdir[gid]) continue;   // Dirichlet row: stays deg * identity (applyDir zeroes the row of every cell matrix)
        }
        const long long p0 = rowptr[r];
        double fr = 0.0;
        bool bad = false;
        for (int k = uoff[u]; k < uoff[u + 1]; ++k) {
            const unsigned it = item[k];
            const long long b = it / (unsigned)nrl;
            const int i = (int)(it % (unsigned)nrl);
            if (sF) {
                const double f = sF[b * nrl + i];
                bad |= ((unsigned)__double2hiint(f) & 0x7ff00000u) == 0x7ff00000u;
                fr += f;
            }
            if (sA) {
                const double* a = sA + (b * nrl + i) * ncl;
                const PosT* ps = pos + aidx[k] * ncl;
                const long long e = bf_tet[b];
                for (int j = 0; j < ncl; ++j) {
                    const double v = a[j];
                    bad |= ((unsigned)__double2hiint(v) & 0x7ff00000u) == 0x7ff00000u;
                    if (fabs(v) <= drop) continue;   // |A| > drop_val (assembler.inl:416); NaN passes and poisons the entry
                    if (isdir) {
                        const int cc = e2c[(long long)j * ntet + e];
                        const long long c = (cc > 0 ? cc : -cc) - 1;
                        if (cc != 0 && isdir[c]) { fr -= v * bc[c]; continue; }   // F(i) -= A(i,k) bc; column k zeroed
                    }
                    val[p0 + ps[j]] += v;
                }
            }
        }
        if (rhs) rhs[r] += fr;
        if (bad) *status = 1;
    }
}

}  // namespace

extern "C" {

int afb_boundary_set(afb_ctx* ctx, int64_t nbf, const int32_t* face_tet, const int32_t* face_num, int mem_space) {
    if (!ctx) return -7;
    ctx->bf_n = 0; ctx->bf_plan_valid = false;
    if (nbf == 0) return 0;
    if (nbf < 0 || !face_tet || !face_num || nbf > 2147483000LL) { set_error(ctx, "afb_boundary_set: bad arguments"); return -7; }
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    cudaSetDevice(ctx->device);
    const cudaMemcpyKind kin = mem_space == AFB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    AFB_CUDA(ctx, ctx->bf_tet.reserve(nbf * sizeof(int32_t)));
    AFB_CUDA(ctx, ctx->bf_face.reserve(nbf * sizeof(int32_t)));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->bf_tet.p, face_tet, nbf * sizeof(int32_t), kin, ctx->stream));
    AFB_CUDA(ctx, cudaMemcpyAsync(ctx->bf_face.p, face_num, nbf * sizeof(int32_t), kin, ctx->stream));
    AFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bf_n = nbf;
    return 0;
}

}  // extern "C"

namespace {

// row -> (face, local row) lists for the current dof map / pattern
int faces_plan(afb_ctx* ctx) {
    const long long nbf = ctx->bf_n, ntet = ctx->ntet;
    const int nrl = ctx->nrow_loc;
    const long long n = nbf * nrl;
    if (n > 2147483000LL) { set_error(ctx, "afb_assemble_faces: too many boundary faces"); return -7; }
    cudaStream_t st = ctx->stream;
    afb::DevBuf key, key2, item, cubtmp, cnt, nrun;
    auto cleanup = [&]() { key.release(); key2.release(); item.release(); cubtmp.release(); cnt.release(); nrun.release(); };
#define F_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return afb::cuda_fail(ctx, _e, #call); } } while (0)
    F_CUDA(key.reserve(n * 4)); F_CUDA(key2.reserve(n * 4)); F_CUDA(item.reserve(n * 4)); F_CUDA(ctx->bf_item.reserve(n * 4));
    F_CUDA(ctx->bf_urow.reserve(n * 4)); F_CUDA(cnt.reserve((n + 1) * 4)); F_CUDA(ctx->bf_uoff.reserve((n + 1) * 4)); F_CUDA(nrun.reserve(8));
    F_CUDA(ctx->flag.reserve(64));
    F_CUDA(cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    int* bad = ctx->flag.as<int>() + 8;
    k_bf_items<<<grid_for(n), 256, 0, st>>>(nbf, nrl, ntet, ctx->bf_tet.as<int32_t>(), ctx->bf_face.as<int32_t>(), ctx->e2r.as<int32_t>(), key.as<unsigned>(),
                                           item.as<unsigned>(), bad);
    size_t tb = 0, tb2 = 0, tb3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key.as<unsigned>(), key2.as<unsigned>(), item.as<unsigned>(), ctx->bf_item.as<unsigned>(), n, 0, 32, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, tb2, key2.as<unsigned>(), ctx->bf_urow.as<unsigned>(), cnt.as<int>(), nrun.as<int>(), n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, cnt.as<int>(), ctx->bf_uoff.as<int>(), n + 1, st);
    F_CUDA(cubtmp.reserve(std::max(tb, std::max(tb2, tb3))));
    F_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, key.as<unsigned>(), key2.as<unsigned>(), item.as<unsigned>(), ctx->bf_item.as<unsigned>(), n, 0, 32, st));
    F_CUDA(cudaMemsetAsync(cnt.p, 0, (n + 1) * 4, st));
    F_CUDA(cub::DeviceRunLengthEncode::Encode(cubtmp.p, tb2, key2.as<unsigned>(), ctx->bf_urow.as<unsigned>(), cnt.as<int>(), nrun.as<int>(), n, st));
    int nu = 0, hb = 0;
    F_CUDA(cudaMemcpyAsync(&nu, nrun.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    F_CUDA(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    F_CUDA(cudaStreamSynchronize(st));
    if (hb == 1) { cleanup(); set_error(ctx, "afb_boundary_set: face outside the mesh or face number outside 0..3"); return -7; }
    if (hb == 2) { cleanup(); set_error(ctx, "afb_assemble_faces: dof maps with orientation signs are not supported"); return -3; }
    // counts beyond the runs are zero (memset), so the scan over nu + 1 entries ends with the total
    F_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp.p, tb3, cnt.as<int>(), ctx->bf_uoff.as<int>(), (long long)nu + 1, st));
    // the last run holds the skipped rows (key 0xFFFFFFFF) when there are any
    unsigned last = 0;
    if (nu > 0) F_CUDA(cudaMemcpyAsync(&last, ctx->bf_urow.as<unsigned>() + (nu - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    int total = 0;
    F_CUDA(cudaStreamSynchronize(st));
    if (nu > 0 && last == 0xffffffffu) --nu;
    F_CUDA(cudaMemcpyAsync(&total, ctx->bf_uoff.as<int>() + nu, sizeof(int), cudaMemcpyDeviceToHost, st));
    F_CUDA(cudaStreamSynchronize(st));
    F_CUDA(ctx->bf_aidx.reserve(std::max(1, total) * sizeof(long long)));
    if (total > 0)
        k_bf_adj<<<grid_for(total), 256, 0, st>>>(total, nrl, key2.as<unsigned>(), ctx->bf_item.as<unsigned>(), ctx->bf_tet.as<int32_t>(),
                                                 ctx->radj_ptr.as<long long>(), ctx->radj.as<unsigned>(), ctx->bf_aidx.as<long long>(), bad);
    F_CUDA(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    F_CUDA(cudaStreamSynchronize(st));
    ctx->launches += 4;
    cleanup();
#undef F_CUDA
    if (hb) { set_error(ctx, "afb_assemble_faces: a boundary face's cell is missing from the gather plan"); return -4; }
    ctx->bf_nu = nu;
    ctx->bf_plan_valid = true;
    return 0;
}

}  // namespace

extern "C" int afb_assemble_faces(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs, const afb_form* rhs_forms, double* csr_val,
                                  double* rhs, double drop_val, int mem_space) {
    if (!ctx) return -7;
    if (ctx->ntet <= 0) { set_error(ctx, "Mesh was not specified"); return -6; }
    if (!ctx->has_pattern) { set_error(ctx, "pattern was not built: call afb_pattern_build first"); return -6; }
    if ((nforms > 0 && !forms) || (nrhs > 0 && !rhs_forms) || nforms < 0 || nrhs < 0) { set_error(ctx, "afb_assemble_faces: bad form arrays"); return -7; }
    // matrix forms need csr_val (with essential conditions they also move the Dirichlet columns into rhs when it is given)
    const bool doA = csr_val != nullptr && nforms > 0, doF = rhs != nullptr && (nrhs > 0 || (ctx->has_dirichlet && doA));
    if (ctx->bf_n == 0 || (!doA && !doF)) return 0;
    if (ctx->has_signs) { set_error(ctx, "afb_assemble_faces: dof maps with orientation signs are not supported"); return -3; }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    if (!ctx->bf_plan_valid) { const int rc = faces_plan(ctx); if (rc) return rc; }
    const long long nbf = ctx->bf_n, nrows = ctx->row_end - ctx->row_begin;
    const int nrl = ctx->nrow_loc, ncl = ctx->ncol_loc;
    const int nfA = doA ? nforms : 0, nfF = rhs ? nrhs : 0;
    std::vector<OpInfo> oa(nfA + nfF), ob(nfA + nfF);
    std::vector<afb_form> fm(nfA + nfF);
    std::vector<const double*> Dd(nfA + nfF, nullptr);
    size_t total = 0;
    std::vector<size_t> offs(nfA + nfF, 0), sizes(nfA + nfF, 0);
    for (int k = 0; k < nfA + nfF; ++k) {
        fm[k] = k < nfA ? forms[k] : rhs_forms[k - nfA];
        if (resolve_op(fm[k].opA, fm[k].femA, fm[k].vecA, &oa[k]) || resolve_op(fm[k].opB, fm[k].femB, fm[k].vecB, &ob[k])) {
            set_error(ctx, "unsupported operator/space in form");
            return -3;
        }
        if (k >= nfA) {
            if (!(fm[k].opA == AFB_IDEN && fm[k].femA == AFB_FEM_P0 && fm[k].vecA == 1)) { set_error(ctx, "rhs form must use OpA = IDEN(P0)"); return -7; }
            fm[k].col_off = 0;
            if (fm[k].row_off < 0 || fm[k].row_off + ob[k].nfa > nrl) { set_error(ctx, "rhs form block outside the element vector"); return -7; }
        } else if (fm[k].row_off < 0 || fm[k].col_off < 0 || fm[k].row_off + ob[k].nfa > nrl || fm[k].col_off + oa[k].nfa > ncl) {
            set_error(ctx, "form block outside the element matrix");
            return -7;
        }

        const int dlen = form_dlen(fm[k], oa[k], ob[k]);
        const int q = afb_tri_quadrature(fm[k].quad_order, nullptr, nullptr, 0);
        if (q < 0) { set_error(ctx, "quadrature order must be in 0..20"); return -7; }
        const size_t n = fm[k].coef_layout == AFB_COEF_CONST ? 1 : (fm[k].coef_layout == AFB_COEF_PER_TET ? (size_t)nbf : (size_t)nbf * q);
        sizes[k] = dlen ? n * dlen : 0;
        if (sizes[k] && !fm[k].D) { set_error(ctx, "tensor data missing"); return -7; }
        if (sizes[k] && fm[k].coef_space == AFB_HOST) { offs[k] = total; total += (sizes[k] + 1) & ~(size_t)1; }
    }
    if (total) AFB_CUDA(ctx, ctx->coef.reserve(total * sizeof(double)));
    for (int k = 0; k < nfA + nfF; ++k) {
        if (!sizes[k]) continue;
        if (fm[k].coef_space == AFB_HOST) {
            AFB_CUDA(ctx, cudaMemcpyAsync(ctx->coef.as<double>() + offs[k], fm[k].D, sizes[k] * sizeof(double), cudaMemcpyHostToDevice, st));
            Dd[k] = ctx->coef.as<double>() + offs[k];
        } else Dd[k] = fm[k].D;
    }
    // caller's arrays: this call ADDS, so host arrays make a round trip
    double* dval = csr_val;
    double* drhs = rhs;
    if (mem_space == AFB_HOST) {
        if (csr_val) {
            AFB_CUDA(ctx, ctx->io_val.reserve(std::max<long long>(1, ctx->nnz) * sizeof(double)));
            dval = ctx->io_val.as<double>();
            AFB_CUDA(ctx, cudaMemcpyAsync(dval, csr_val, ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (rhs) {
            AFB_CUDA(ctx, ctx->io_rhs.reserve(std::max<long long>(1, nrows) * sizeof(double)));
            drhs = ctx->io_rhs.as<double>();
            AFB_CUDA(ctx, cudaMemcpyAsync(drhs, rhs, nrows * sizeof(double), cudaMemcpyHostToDevice, st));
        }
    }
    // ---- face matrices, staged [face][local row][local column]
    double *sA = nullptr, *sF = nullptr;
    if (nfA) {
        AFB_CUDA(ctx, ctx->stageA.reserve((size_t)nbf * nrl * ncl * sizeof(double)));
        sA = ctx->stageA.as<double>();
        AFB_CUDA(ctx, cudaMemsetAsync(sA, 0, (size_t)nbf * nrl * ncl * sizeof(double), st));
    }
    if (nfF) {
        AFB_CUDA(ctx, ctx->stageF.reserve((size_t)nbf * nrl * sizeof(double)));
        sF = ctx->stageF.as<double>();
        AFB_CUDA(ctx, cudaMemsetAsync(sF, 0, (size_t)nbf * nrl * sizeof(double), st));
    }
    for (int k = 0; k < nfA + nfF; ++k) {
        const bool matrix = k < nfA;
        const int rc = launch_form(ctx, fm[k], oa[k], ob[k], nbf, ctx->x.as<double>(), ctx->y.as<double>(), ctx->z.as<double>(), ctx->v[0].as<int32_t>(),
                                   ctx->v[1].as<int32_t>(), ctx->v[2].as<int32_t>(), ctx->v[3].as<int32_t>(), nullptr, matrix ? sA : sF,
                                   matrix ? (long long)nrl * ncl : nrl, matrix ? ncl : 1, matrix ? 1 : 0, 1, Dd[k], ctx->bf_face.as<int32_t>(),
                                   ctx->bf_tet.as<int32_t>());
        if (rc) return rc;
    }
    // ---- add into the CSR rows / rhs
    AFB_CUDA(ctx, ctx->flag.reserve(64));
    AFB_CUDA(ctx, cudaMemsetAsync(ctx->flag.p, 0, 64, st));
    const bool dir = ctx->has_dirichlet;
    const unsigned char* isdir = dir ? ctx->dir_flag.as<unsigned char>() : nullptr;
    const double* bc = dir ? ctx->dir_val.as<double>() : nullptr;
    const int32_t* dcol = ctx->has_diag ? ctx->diag_col.as<int32_t>() : nullptr;
    if (dir) { const int rc = dirichlet_prepare(ctx); if (rc) return rc; }
    const int32_t* rgid = (dir && ctx->row_gid_valid) ? ctx->row_gid.as<int32_t>() : nullptr;
    if (ctx->bf_nu > 0) {
        if (ctx->pos_bytes == 1)
            k_bf_gather<unsigned char><<<grid_for(ctx->bf_nu), 128, 0, st>>>(ctx->bf_nu, ctx->bf_urow.as<unsigned>(), ctx->bf_uoff.as<int>(), ctx->bf_item.as<unsigned>(),
                ctx->bf_aidx.as<long long>(), nrl, ncl, ctx->ntet, ctx->bf_tet.as<int32_t>(), sA, sF, ctx->pos.as<unsigned char>(),
                ctx->rowptr.as<long long>(), ctx->e2c.as<int32_t>(), ctx->row_begin, dcol, rgid, isdir, bc, dval, drhs, drop_val, ctx->flag.as<int>());
        else
            k_bf_gather<unsigned short><<<grid_for(ctx->bf_nu), 128, 0, st>>>(ctx->bf_nu, ctx->bf_urow.as<unsigned>(), ctx->bf_uoff.as<int>(), ctx->bf_item.as<unsigned>(),
                ctx->bf_aidx.as<long long>(), nrl, ncl, ctx->ntet, ctx->bf_tet.as<int32_t>(), sA, sF, ctx->pos.as<unsigned short>(),
                ctx->rowptr.as<long long>(), ctx->e2c.as<int32_t>(), ctx->row_begin, dcol, rgid, isdir, bc, dval, drhs, drop_val, ctx->flag.as<int>());
        ctx->launches++;
        AFB_CUDA(ctx, cudaGetLastError());
    }
    if (mem_space == AFB_HOST) {
        if (csr_val && ctx->nnz) AFB_CUDA(ctx, cudaMemcpyAsync(csr_val, dval, ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (rhs && nrows) AFB_CUDA(ctx, cudaMemcpyAsync(rhs, drhs, nrows * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int bad = 0;
    AFB_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(ctx, cudaStreamSynchronize(st));
    if (bad) { set_error(ctx, "not a number in local matrix or rhs"); return -1; }
    return 0;
}
