// Internal declarations shared by the .cu files of libanifem_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/anifem_b200.h"

#define AFB_MAX_BASE_NF 20

namespace afb {

void set_error(afb_ctx* ctx, const std::string& msg);
int cuda_fail(afb_ctx* ctx, cudaError_t e, const char* what);

#define AFB_CUDA(ctx, call)                                             \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return afb::cuda_fail(ctx, _e, #call);   \
    } while (0)

// growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool borrowed = false;  // view of a buffer owned by another context (pair sub-contexts share the mesh)
    cudaError_t reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// description of Operator<op, FemFix|FemVec> resolved on the host
struct OpInfo {
    int op, fem, vec;
    int nf_base;   // basis functions of the scalar base space
    int nfa;       // Nfa of the operator
    int dim;       // Dim of the operator
    int dim_base;  // rows of the per-component base table: IDEN 1, GRAD/DIV 3
};
int resolve_op(int op, int fem, int vec, OpInfo* out);

// host-side basis tables at the points of a rule (the product's own evaluation of the P0..P3
// Lagrange bases; formulas documented in afb_tables.cpp)
void basis_values(int fem, int q, const double* XYL, double* phi /*[n*nf + i]*/);
void basis_ref_grads(int fem, int q, const double* XYL, double* G /*[(n*nf + i)*3 + d]*/);
int tet_rule(int order, const double** p, const double** w);
int tri_rule(int order, const double** p /*3 per point*/, const double** w);

}  // namespace afb

// device view of one volume form, passed by value to the element kernels
struct FormDev {
    int opA, vecA, nfbA, nfa, idim, dbA;   // dbA = dim_base of A
    int opB, vecB, nfbB, nfb, jdim, dbB;
    int same;            // OpA == OpB (V = U)
    int q;               // quadrature points
    int qc;              // points per shared-memory chunk
    const double* W;     // [q]
    const double* phiA;  // [q*nfbA]
    const double* grdA;  // [q*nfbA*3]
    const double* phiB;
    const double* grdB;
    int ttype, layout, dlen;
    const double* D;
    double alpha;
    // output addressing: out[e*s_e + (row_off+ib)*s_ib + (col_off+ia)*s_ia]
    long long s_e, s_ib, s_ia;
    int row_off, col_off;
    int add;             // 0: store, 1: add into out
    // surface integrals (fem3Dface): face[e] in 0..3 selects the face {k, k+1, k+2 mod 4} of item e; the tables then hold the four
    // lifted triangle rules one after the other (phi[face][q*nfb], grd[face][q*nfb*3]) and the measure is the face area
    const int32_t* face;      // NULL = volume integral
    const int32_t* item_tet;  // NULL, or the mesh element behind item e (coefficients and output stay indexed by the item)
};

struct TableEntry {
    int fem, order;
    double *W, *phi, *grd;  // device
};

// intervals of a scalar field inside the row (local ids) or column (global ids) space; the scalar numbering of the field is
// the concatenation of the intervals.  One interval under a single-rank NATURAL numbering, one per owner rank under the
// per-rank numbering of a partitioned mesh (afb_fields_set).
#define AFB_MAX_SEG 8
struct SegMap {
    int n;
    long long start[AFB_MAX_SEG], count[AFB_MAX_SEG];
    __host__ __device__ long long total() const { long long t = 0; for (int k = 0; k < n; ++k) t += count[k]; return t; }
    __host__ __device__ long long to_scalar(long long id) const {   // -1: not in the field
        long long pre = 0;
        for (int k = 0; k < n; ++k) {
            if (id >= start[k] && id < start[k] + count[k]) return pre + (id - start[k]);
            pre += count[k];
        }
        return -1;
    }
    __host__ __device__ long long to_full(long long s) const {
        for (int k = 0; k < n; ++k) {
            if (s < count[k]) return start[k] + s;
            s -= count[k];
        }
        return -1;
    }
};
// scalar field (variable, component) of a dof map: base space, local offset, row / column intervals
struct Field { int fem, nloc, loff; long long goff, count; SegMap rows, cols; };
struct PairPlan { int femR, femC; afb_ctx* sub; };

struct afb_ctx {
    int device = 0;
    bool is_sub = false;          // pair sub-context of afb_blocks.cu
    std::vector<Field> fields;    // set by afb_dofmap_natural
    std::vector<PairPlan> pairs;  // gather plans of (row space, column space) pairs
    std::vector<afb::DevBuf> block_dst;   // [fR*nf + fC]: first CSR entry of block (fR,fC) per (slice, lane) of its pair plan
    std::vector<afb::DevBuf> block_tix;   // [fR*nf + fC]: int32 per (slice, lane): 0 = the block is contiguous in the row, else 1 + first entry of its offset table
    std::vector<afb::DevBuf> block_tab;   // [fR*nf + fC]: uint16 offsets (relative to block_dst) of the entries of non-contiguous blocks
    std::vector<afb::DevBuf> block_rdst;  // [fR]: int32 per (slice, lane) of the pair plan (space of fR, space of fR): local row that receives the rhs
    afb::DevBuf block_gap;        // int32[n_gap_rows]: rows holding entries that belong to no block
    int n_gap_rows = 0;
    bool blocks_ready = false;
    bool blocks_cover = true;     // every entry of the pattern belongs to some block (else the values are cleared before a non-accumulating assembly)
    bool fields_custom = false;   // fields came from afb_fields_set (segmented numbering)
    std::vector<TableEntry> table_cache;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;

    // mesh (SoA)
    int64_t nnode = 0, ntet = 0;
    afb::DevBuf x, y, z;          // double[nnode]
    afb::DevBuf v[4];             // int32[ntet]

    // dof map: codes sign*(local+1), 0 = skipped; element-fastest SoA [i*ntet + e]
    int nrow_loc = 0, ncol_loc = 0;
    int64_t row_begin = 0, row_end = 0, ncols_global = 0;
    afb::DevBuf e2r;              // int32[nrow_loc*ntet]  code of (row - row_begin)
    afb::DevBuf e2c;              // int32[ncol_loc*ntet]  code of global column
    bool has_dofmap = false;
    bool has_signs = false;
    bool has_diag = false;        // explicit forced-diagonal columns (afb_dofmap_set_diag)
    afb::DevBuf diag_col;         // int32[nrows], -1 = none

    // pattern + plan
    bool has_pattern = false;
    int64_t nnz = 0;
    int max_row_len = 0;
    afb::DevBuf rowptr;           // int64[nrows+1]
    afb::DevBuf colind;           // int32[nnz]
    afb::DevBuf radj_ptr;         // int64[nrows+1]
    afb::DevBuf radj;             // uint32[n_adj]: e*nrow_loc + i, ascending per row
    afb::DevBuf pos;              // uint8|uint16[n_adj*ncol_loc]: slot of column j of adjacency entry a inside its row
    int pos_bytes = 2;
    int64_t n_adj = 0;
    bool pos_has_dup = false;     // some element maps two local columns to the same global column

    // cluster plan of the thread-per-row gather (afb_rows.cu)
    bool has_rows_plan = false;
    int rp_nloc = 0, rp_ncol = 0, rp_gcap = 0, rp_maxlen = 0;
    long long rp_steps = 0, rp_ncl = 0, rp_nslices = 0, rp_long_steps[2] = {0, 0};
    afb::DevBuf rp_new2old, rp_old2new;  // uint32[ntet]: Morton order of the elements
    afb::DevBuf rp_cs;            // int32[ncl+1]: slices of a cluster
    afb::DevBuf rp_eptr;          // int32[ncl+1]: element list of a cluster
    afb::DevBuf rp_elist;         // uint32: Morton ids of the elements a cluster touches
    afb::DevBuf rp_order;         // uint32[nslices*32]: rows of a slice (0xFFFFFFFF = empty lane)
    afb::DevBuf rp_p0, rp_len, rp_smax;  // int64 / uint16 [nslices*32]: first CSR entry and length of the row; uint16[nslices] longest row
    afb::DevBuf rp_cnt;           // uint16[nslices*nloc]: visit-steps of class i in slice s
    afb::DevBuf rp_sptr;          // int64[nslices+1]: first visit-step of a slice
    afb::DevBuf rp_ell;           // uint32[steps*NW*32]: slot bytes + local element per (step, lane)
    afb::DevBuf rp_clist;         // int32[ncl]: clusters holding priority rows first, then the others (phased assembly)
    bool rp_prio_valid = false, phase_done = false;
    int phase_status = 0;
    long long rp_nprio = 0, priority_row = -1;

    // ring plan of square P2 problems (afb_rings.cu, afb_ring_plan.h)
    bool has_ring_plan = false, rg_prio_valid = false, ring_plan_tried = false;
    double ring_plan_ms = 0.0;   // host + transfer time of the last ring plan build
    long long rg_ncl = 0, rg_nslices = 0, rg_nsteps = 0, rg_nvert = 0, rg_nz = 0, rg_nprio = 0, rg_vsplit = 0;
    int rg_gcap = 0, rg_imgcap = 0, rg_stepcap = 0, rg_xcap = 0, rg_edges = 0;
    afb::DevBuf rg_cinfo, rg_elist, rg_hdr, rg_steps, rg_desc, rg_xpos;
    afb::DevBuf rg_vptr, rg_vlist, rg_vdpos, rg_vrow, rg_zlist, rg_scratch, rg_clist;

    // multi-GPU interface exchange (afb_comm.cu): NCCL communicator (void* = ncclComm_t), halo plan
    void* comm = nullptr;
    bool own_comm = false;
    int comm_rank = 0, comm_size = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t comm_ev[2] = {nullptr, nullptr};   // [0] context stream -> exchange, [1] exchange -> context stream
    bool has_halo_plan = false, halo_in_flight = false;
    int64_t halo_n_own = 0, halo_nnz_own = 0;
    std::vector<int64_t> halo_send_val, halo_send_rhs, halo_recv_val, halo_recv_rhs;   // [comm_size] counts per peer
    afb::DevBuf halo_val_slots, halo_rhs_slots, halo_val_recv, halo_rhs_recv;
    std::vector<unsigned char> rg_cinfo_host;   // host copy of the cluster records (permuted for phased launches)
    std::vector<unsigned> rg_maxrow, rg_vrow_host;   // largest row a cluster writes to; rows of the vertex list (ascending)

    // essential boundary conditions (afb_dirichlet.cu): per global dof flag + value, list of affected rows
    bool has_dirichlet = false, dir_rows_valid = false;
    afb::DevBuf dir_flag, dir_val, dir_rows;
    afb::DevBuf row_gid;          // int32[nrows]: global dof of every local row (explicit dof maps with diag tables), built with dir_rows
    bool row_gid_valid = false;
    long long n_dir_rows = 0;

    // boundary faces carrying surface terms (afb_faces.cu): face bf_face[b] of element bf_tet[b]; row -> (face, local row) lists
    long long bf_n = 0;
    int bf_nu = 0;                // rows touched by the listed faces
    bool bf_plan_valid = false;
    afb::DevBuf bf_tet, bf_face;  // int32[bf_n]
    afb::DevBuf bf_item;          // uint32: b*nrow_loc + i, sorted by row, ascending inside a row
    afb::DevBuf bf_urow, bf_uoff; // uint32[bf_nu] touched rows, int32[bf_nu+1] their item ranges
    afb::DevBuf bf_aidx;          // int64 per item: row of the slot table (index into the row adjacency)

    // work buffers
    afb::DevBuf stageA, stageF, tables, coef, io_val, io_rhs, flag, tmp1, tmp2, tmp3, xy;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double times[4] = {0, 0, 0, 0};
    const char* k_elem = "";     // kernels of the generic staged path that ran last (afb_last_kernels)
    const char* k_gath = "";
};

// one scalar block form of the tensor representation (afb_tensor.cu): A_e(i,j) = sum_c T[c][i][j] g_e[c]
struct SForm {
    int kind;             // 0 GRADxGRAD, 1 IDENxIDEN, 2 GRAD(A)xIDEN(B), 3 IDEN(A)xGRAD(B)
    int ng;               // components of g_e it uses
    int full;             // kind 0: a 3x3 tensor is given (else c * identity)
    int layout, dstride;  // AFB_COEF_CONST / PER_TET, doubles per element record of D
    int kidx[9];          // where the tensor values sit in the record; -1 = 0, -2 = 1
    double alpha;
    const double* D;      // device
    int femA, femB, nfa, nfb;   // base spaces (trial, test) and their sizes
    int quad_order;
    int row_off, col_off;       // inside the local matrix of the plan it is assembled with
};

namespace afb {
bool make_sform(const afb_form& f, const OpInfo& oa, const OpInfo& ob, const double* Ddev, SForm* out);
int fused_group(afb_ctx* ctx, afb_ctx* plan, const std::vector<SForm>& mat, const std::vector<SForm>& rhsf, double* dval, double* drhs,
                const long long* p0_override, int accumulate, double drop_val, int* status_flag, bool record_events, int phase = 0,
                const int* tix = nullptr, const unsigned short* rtab = nullptr, const int* rdst = nullptr);
// element kernels (afb_element.cu)
int launch_form(afb_ctx* ctx, const afb_form& form, const OpInfo& A, const OpInfo& B, int64_t f,
                const double* x, const double* y, const double* z,                 // SoA nodes (or NULL)
                const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                const double* XY /* 4 arrays 3 x f, AoS variant, or NULL */,
                double* out, long long s_e, long long s_ib, long long s_ia, int add, const double* Ddev,
                const int32_t* face = nullptr, const int32_t* item_tet = nullptr);
int form_dlen(const afb_form& form, const OpInfo& A, const OpInfo& B);
int launch_forms_sq(afb_ctx* ctx, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa, const std::vector<const double*>& Dd,
                    const std::vector<int>& sel, int64_t e_lo, int64_t nel, double* out, long long s_e, const double* XY = nullptr,
                    int colmajor = 0);
// device tables W[q], phi[q*nf], G^[q*nf*3] of (space, rule); uploaded on first use (afb_ctx.cu)
int get_tables(afb_ctx* ctx, int fem, int order, const double** W, const double** phi, const double** grd, bool faces = false);
// block-decomposed fused path of vector / mixed spaces (afb_blocks.cu)
void blocks_clear(afb_ctx* ctx);
void blocks_clear_dst(afb_ctx* ctx);
int blocks_build(afb_ctx* ctx);
int assemble_block_path(afb_ctx* ctx, int nfA, int nfF, const std::vector<afb_form>& fm, const std::vector<OpInfo>& oa,
                        const std::vector<OpInfo>& ob, const std::vector<const double*>& Dd, double* dval, double* drhs, int accumulate,
                        double drop_val, int* status_flag);
// essential boundary conditions (afb_dirichlet.cu)
int dirichlet_apply(afb_ctx* ctx, double* val, double* rhs);
int dirichlet_prepare(afb_ctx* ctx);   // row list + row -> global dof table for the current pattern
// thread-per-row gather + its plan (afb_rows.cu)
int build_rows_plan(afb_ctx* ctx);
bool rows_supports(const afb_ctx* ctx, int nga, int ngf);
int launch_rows(afb_ctx* ctx, int nga, int ngf, const double* TA, const double* TF, const double* gbuf, double* val, double* rhs,
                int accumulate, double drop_val, int* status, const long long* p0_override = nullptr, int phase = 0,
                const int* tix = nullptr, const unsigned short* rtab = nullptr, const int* rdst = nullptr);
int rows_priority_build(afb_ctx* ctx, long long first_priority_row);
// ring-traversal assembly of square P2 problems + its plan (afb_rings.cu)
int build_ring_plan(afb_ctx* ctx);
int ensure_ring_plan(afb_ctx* ctx);   // builds the plan (and its priority split) on first use; 0 ok
int rings_priority_build(afb_ctx* ctx, long long first_priority_row);
bool rings_supports(const afb_ctx* ctx, int nstiff, int nmass, int nload);
int launch_rings(afb_ctx* ctx, const double* TG, const double* Tm, const double* Tf, const double* gbuf, double* val, double* rhs,
                 int accumulate, double drop_val, int* status, int phase);
// gather (afb_gather.cu)
int launch_gather(afb_ctx* ctx, const double* stageA, const double* stageF, double* val, double* rhs,
                  int accumulate, double drop_val, int* status_flag, long long e_lo, long long e_hi);
// afb_comm.cu: destroys an owned NCCL communicator, the exchange stream / events and the halo plan buffers
void comm_release(afb_ctx* ctx);
}  // namespace afb
