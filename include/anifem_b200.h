/* anifem_b200.h -- C ABI of the B200-native element-matrix + global-assembly path.
 *
 * This is the drop-in boundary: plain C, pointers + sizes, no C++/torch
 * types.  Every entry point names the reference interface it replaces (paths relative to the
 * AniFem++ tree, INMOST-DEV/INMOST-FEM).  The C++ mirror of the reference API that sits on top
 * of this file (Ani::fem3Dtet<...>, Ani::Assembler) lives in inmost-fem_b200/include/anifem_b200/.
 *
 * Conventions
 *   - FP64 values, column-major dense matrices, exactly as the reference (fem/fem_memory.h:55-64).
 *   - `mem_space` says where the caller's buffers live: AFB_HOST (copied through the context's
 *     stream) or AFB_DEVICE (used in place / copied device-to-device).  The context owns device
 *     copies of mesh, dof tables, pattern and plan; callers own everything they pass in.
 *   - One context per GPU, single owner, all work on the context's CUDA stream.
 *   - Return codes: 0 ok; -1 a local matrix/rhs value is not finite (assembler.inl:419-424,
 *     475-479); -2 bad mesh element (assembler.inl:309-312); <= -3 usage errors (the reference
 *     throws std::runtime_error there): -3 unsupported operator/space, -4 CUDA/runtime failure,
 *     -5 identity/scalar tensor with incompatible operator dimensions (diff_tensor.h:315-317),
 *     -6 call order / missing state (assembler.inl:195-196,316-317), -7 bad argument.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with -4.
 */
#ifndef ANIFEM_B200_H
#define ANIFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct afb_ctx afb_ctx;

enum { AFB_HOST = 0, AFB_DEVICE = 1 };

/* values = Ani::OperatorType (fem/operators.h:36-44) */
enum { AFB_IDEN = 1, AFB_GRAD = 2, AFB_DIV = 3 };
/* values = Ani::FiniteElement (fem/operators.h:24-34) */
enum { AFB_FEM_P0 = 1, AFB_FEM_P1 = 2, AFB_FEM_P2 = 3, AFB_FEM_P3 = 4 };
/* values = Ani::TensorType (fem/diff_tensor.h:17-22) */
enum { AFB_TENSOR_NULL = 1, AFB_TENSOR_SCALAR = 2, AFB_TENSOR_SYMMETRIC = 3, AFB_TENSOR_GENERAL = 4 };
/* where the coefficient varies: replaces DfuncTraits<..., isConstant> + the per-point callback
 * (fem/diff_tensor.h:36-53): CONST = one tensor, PER_TET = one per tetrahedron (callback ignoring x),
 * PER_POINT = one per quadrature point, the FusiveTensor layout D[len*(n + q*r)] (diff_tensor.h:112-124) */
enum { AFB_COEF_CONST = 0, AFB_COEF_PER_TET = 1, AFB_COEF_PER_POINT = 2 };

/* One volume form  int_T (D OpA(u)) . OpB(v) dx  = one fem3Dtet<OpA,OpB,Traits> call
 * (fem/operations/int_tet.h:17-29).  Operator<op, FemFix<fem>> when vec == 1,
 * Operator<op, FemVec<3,fem>> when vec == 3 (fem/operators.h:50-67,127-155,320-353).
 * D is in the layout a reference user callback writes: a col-major (jdim x idim) matrix
 * K(k,j) at D[k + jdim*j], jdim = Dim(OpB), idim = Dim(OpA) (fem/operations/core.h:27-40);
 * 1 value for TENSOR_SCALAR, unused for TENSOR_NULL. */
typedef struct afb_form {
    int opA, femA, vecA;   /* trial side: columns of the element matrix */
    int opB, femB, vecB;   /* test side: rows of the element matrix */
    int quad_order;        /* 0..20, tetrahedron_quadrature_formulas(order) (fem/quadrature_formulas.h:214) */
    int tensor_type;       /* AFB_TENSOR_* */
    int coef_layout;       /* AFB_COEF_* */
    int coef_space;        /* AFB_HOST / AFB_DEVICE: where D lives */
    const double* D;
    double alpha;          /* the block is scaled by alpha before it is added; used as given by EVERY entry point
                            * (0 contributes nothing): set it to 1.0 explicitly, a zero-initialised struct is a zero form */
    int row_off, col_off;  /* offset of the block inside the (nrow_loc x ncol_loc) element matrix */
} afb_form;

/* ---- context --------------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for a private one. */
int afb_ctx_create(int device, void* cuda_stream, afb_ctx** out);
void afb_ctx_destroy(afb_ctx* ctx);
const char* afb_last_error(const afb_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
int afb_sync(afb_ctx* ctx);
/* number of kernels this library launched on the context since the last reset (bench evidence) */
int64_t afb_launch_count(afb_ctx* ctx, int reset);
/* the cudaStream_t every kernel / copy of this context is issued on (the one given to afb_ctx_create or the private one):
 * work issued by the caller around afb_assemble_phase (NCCL exchange, afb_halo_add inputs) must be ordered against it. */
void* afb_stream_get(afb_ctx* ctx);

/* ---- element level: replaces Ani::fem3Dtet (fem/operations/int_tet.inl:3-57) ------------------ */
/* Batched element matrices: XYk are 3 x f col-major (fem/geometry.h:108-122), A is nfB x (nfA*f)
 * col-major, block r at column offset nfA*r (core.h:41).  The tets need not be oriented. */
int afb_fem3dtet_batched(afb_ctx* ctx, const afb_form* form, int64_t f,
                         const double* XY0, const double* XY1, const double* XY2, const double* XY3,
                         double* A, int mem_space);
/* Ani::fem3Dface<OpA,OpB,FuncTraits> (fem/operations/int_face.h:15-159, int_face.inl:160-199): element matrices of the
 * surface integral int_f (D OpA(u)) . OpB(v) over face face_num[r] of tet r; face k = vertices {k, k+1, k+2 mod 4}
 * (int_face.h:49-55).  Operators, tensor kinds, layouts and the layout of A as afb_fem3dtet_batched; the rule is the
 * reference's triangle rule of form->quad_order (PER_POINT coefficients follow its points), the measure the face area.
 * Errors: -7 "Wrong face index" (int_face.inl:28), the others as afb_fem3dtet_batched. */
int afb_fem3dface_batched(afb_ctx* ctx, const afb_form* form, int64_t f, const int32_t* face_num /*[f]*/,
                          const double* XY0, const double* XY1, const double* XY2, const double* XY3, double* A, int mem_space);
/* triangle_quadrature_formulas(order) (fem/quadrature_formulas.cpp:109-516): p[3*q] barycentric, w[q], sum w = 1; returns q */
int afb_tri_quadrature(int order, double* p, double* w, int capacity);
/* Nfa / Dim of an operator (Operator<>::Nfa, ::Dim) */
int afb_op_dims(int op, int fem, int vec, int* nfa, int* dim);
/* quadrature rule (fem/quadrature_formulas.cpp:526,1493-1502): returns q, fills p[4q], w[q] if non-NULL */
int afb_tet_quadrature(int order, double* p, double* w, int capacity);
/* physical quadrature points XYG[3*(n + q*r)] (core.inl:249-269), for building PER_POINT coefficients */
int afb_quad_points(afb_ctx* ctx, int order, int64_t f, const double* XY0, const double* XY1,
                    const double* XY2, const double* XY3, double* XYG, int mem_space);

/* FE-function evaluation, fem3DapplyL (fem/operations/eval.h:13-120, core.inl:369-404): Op(u_h) at q points given by barycentric
 * coordinates XYL[4q] (the same on every tet): opU[k + dim*(n + q*r)] = sum_i Op(phi_i)(x_n)[k] * dofs[i + nfa*r]; dofs is nfa x f,
 * opU is (dim*q) x f, both col-major. */
int afb_fem3dapply_batched(afb_ctx* ctx, int op, int fem, int vec, int q, const double* XYL, int64_t f, const double* XY0,
                           const double* XY1, const double* XY2, const double* XY3, const double* dofs, double* opU, int mem_space);
/* The same on the context's mesh at the points of the rule `order`, the dofs of the variable occupying the local slots
 * [col_off, col_off + Nfa) gathered from the global vector u[ncols_global] through the dof map (what InitValueSetFromTag +
 * fem3DapplyL do inside a nonlinear local assembler, assembler.h:13-48).  out[dim*(n + q*e)] has the layout of a PER_POINT
 * coefficient, so it can be passed to afb_assemble as form.D.  Returns q (>= 0) or an error code. */
int afb_eval_quadrature(afb_ctx* ctx, int op, int fem, int vec, int col_off, int order, const double* u, double* out, int mem_space);

/* ---- mesh: the per-cell inputs of AssemblerT::Assemble (assembler.inl:353-364) ---------------- */
/* SoA coordinates + connectivity.  Copies into the context. */
int afb_mesh_set(afb_ctx* ctx, int64_t nnode, const double* x, const double* y, const double* z,
                 int64_t ntet, const int32_t* v0, const int32_t* v1, const int32_t* v2, const int32_t* v3,
                 int mem_space);
/* GenerateParallelepiped on [0,size]^3 (utils/mesh_utils.cpp:20-48,110-145): the (bx,by,bz)+(lx,ly,lz)
 * block of hexes of an nx*ny*nz cube, 6 tets per hex, nodes in (i,j,k) order; built on the device. */
int afb_mesh_cube(afb_ctx* ctx, int nx, int ny, int nz, double size,
                  int bx, int by, int bz, int lx, int ly, int lz);
/* reorderNodesOnTetrahedron (inmost_interface/ordering.inl:8-26): swap nodes 2,3 where det < 0 */
int afb_mesh_orient(afb_ctx* ctx);
int afb_mesh_get(afb_ctx* ctx, int64_t* nnode, int64_t* ntet, double* xyz_soa /*3*nnode*/, int32_t* v_soa /*4*ntet*/, int mem_space);

/* ---- dof map: m_indexesR / m_indexesC of fill_assemble_templates (assembler.inl:139-184) ------ */
/* Explicit tables: elem2row[i + nrow_loc*e], elem2col[j + ncol_loc*e] hold assemble_index_encode
 * codes sign*(id+1), 0 = row skipped (ghost) (assembler.inl:49-55).  Rows live in
 * [row_begin,row_end) = [getBegInd(),getEndInd()), columns in [0,ncols_global). */
int afb_dofmap_set(afb_ctx* ctx, int nrow_loc, int ncol_loc, const int64_t* elem2row, const int64_t* elem2col,
                   int64_t row_begin, int64_t row_end, int64_t ncols_global, int mem_space);
/* GlobEnumeration NATURAL on one rank (global_enumerator.cpp:562-605,702-777): variables
 * (fem[v], vec[v]) numbered (VAR, DIM, ELEM_TYPE, ELEM_ID, DOF_ID); edges/faces get ids in
 * lexicographic order of their sorted node pairs/triples (our stand-in for INMOST GlobalIDs).
 * Same trial and test space. Builds the tables on the device. */
int afb_dofmap_natural(afb_ctx* ctx, int nvars, const int* fem, const int* vec);
/* Optional: the column of the forced diagonal entry of every row (set_elements_on_matrix_diagonal,
 * assembler.inl:114-136) when rows are not numbered like columns (multi-GPU: owned rows followed by the
 * interface rows of other ranks); -1 = no forced entry.  Default: row_begin + r. */
int afb_dofmap_set_diag(afb_ctx* ctx, const int64_t* diag_col /*nrows*/, int mem_space);
/* Optional, for vector-valued / mixed problems with an explicit dof map (afb_dofmap_set): the scalar fields (variable,
 * component) behind the local dofs, so that the assembly can run block by block on scalar gather plans like it does after
 * afb_dofmap_natural (BandDenseMatrix structure of FemVec / FemCom operators, fem/operators.h:127-155,189-259).  Field f
 * lives on base space fem[f] and owns the local dofs loff[f] .. loff[f]+nf(fem[f])-1.  Under the per-rank NATURAL
 * enumeration (global_enumerator.cpp:594-604,702-777) a field is contiguous inside every rank's interval only: its rows
 * (ids relative to row_begin) are nseg_row intervals row_seg[(f*nseg_row + k)*2 + {0,1}] = {first, count} and its global
 * columns nseg_col intervals (count 0 = unused entry); the scalar numbering of a field is the concatenation of its
 * intervals and must enumerate the entities of all fields of one space in the same order.  Call after afb_dofmap_set
 * [+ afb_dofmap_set_diag] and before afb_pattern_build / afb_pattern_set.  nfields = 0 clears. */
int afb_fields_set(afb_ctx* ctx, int nfields, const int* fem, const int* loff, int nseg_row, const int64_t* row_seg,
                   int nseg_col, const int64_t* col_seg);
int afb_dofmap_get(afb_ctx* ctx, int* nrow_loc, int* ncol_loc, int64_t* row_begin, int64_t* row_end,
                   int64_t* ncols_global, int64_t* elem2row, int64_t* elem2col, int mem_space);

/* ---- pattern: AssembleTemplate (assembler.inl:589-695) + forced diagonal (:114-136) ----------- */
/* Structural CSR over the owned rows, columns ascending, built on the device from the dof map;
 * also builds the gather plan (row -> contributing element rows, slot table) used by afb_assemble. */
int afb_pattern_build(afb_ctx* ctx, int64_t* nnz);
int afb_pattern_get(afb_ctx* ctx, int64_t* rowptr /*nrows+1*/, int32_t* colind /*nnz*/, int mem_space);
/* Use a caller-supplied sorted pattern instead (it must contain every structural entry of the dof map): this is
 * "Assemble into a matrix that already includes the template" (opts.is_mtx_include_template, assembler.h:214-220),
 * e.g. the union with the columns that other ranks contribute to interface rows. Rebuilds the gather plan. */
int afb_pattern_set(afb_ctx* ctx, const int64_t* rowptr /*nrows+1*/, const int32_t* colind /*nnz*/, int64_t nnz, int mem_space);

/* ---- assembly: AssemblerT::Assemble / AssembleMatrix / AssembleRHS (assembler.inl:313-488,
 * 497-580, 704-865) with is_mtx_include_template = use_ordered_insert = true ------------------ */
/* csr_val[nnz] and rhs[nrows] (either may be NULL).  accumulate != 0 adds to the existing contents
 * like the reference (assembler.inl:305-306), otherwise they are overwritten.  Entries with
 * |A_e(i,j)| <= drop_val are not added (assembler.h:212, assembler.inl:416).  rhs forms use the
 * reference's RHS trick OpA = IDEN(P0) (tests/fem/operations/int_tet_test.cpp:448-501): opA/femA/vecA
 * of an rhs form must be (AFB_IDEN, AFB_FEM_P0, 1).  Deterministic: fixed summation order
 * (ascending element index per row), no atomics. */
int afb_assemble(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs_forms, const afb_form* rhs_forms,
                 double* csr_val, double* rhs, int accumulate, double drop_val, int mem_space);

/* The element-evaluator plug-in point of the reference (MatFuncWrap, inmost_interface/func_wrap.h:96-187, installed with
 * AssemblerT::SetMatFunc / SetRHSFunc / SetMatRHSFunc, assembler.h:326-328): the caller evaluates the local matrices itself (a
 * host lambda per cell, e.g. examples/tutorials/ex1.cpp:83-106) and hands them over for the cells [e_lo, e_lo + nel) of the
 * context's mesh: A_elem = nel matrices in the reference's layout m_A[j*nRows + i] (column-major nrow_loc x ncol_loc,
 * assembler.inl:417), F_elem = nel local right-hand sides; either may be NULL; elem_space says where they live.  The
 * contributions are ADDED into csr_val / rhs (mem_space) with the scatter rule of assembler.inl:397-481 (|A| > drop_val, signs
 * of the index codes; essential conditions are whatever the lambda did with applyDir).  Returns 0 / -1 like afb_assemble. */
int afb_assemble_elemental(afb_ctx* ctx, int64_t e_lo, int64_t nel, const double* A_elem, const double* F_elem, int elem_space,
                           double* csr_val, double* rhs, double drop_val, int mem_space);

/* Essential (Dirichlet) boundary conditions = applyDir(A, F, k, bc) on every Dirichlet dof of every cell, as the reference's local
 * assemblers do (fem/operations/dc_on_dof.h:27-45; examples/tutorials/ex1.cpp:96-105).  is_dirichlet[ncols_global] (0/1) and
 * value[ncols_global] are indexed by the global dof; NULL clears.  Every following afb_assemble then returns constrained rows:
 * a Dirichlet row r is deg(r) on the diagonal (deg = cells containing the dof) and deg(r)*bc_r in the rhs, free rows lose their
 * Dirichlet columns to the rhs (b_r -= A_rc bc_c).  The setting is dropped when the dof map changes. */
int afb_dirichlet_set(afb_ctx* ctx, const unsigned char* is_dirichlet, const double* value, int mem_space);

/* Phased assembly for overlapping the interface exchange with the bulk of the work (multi-GPU, device buffers, values are
 * overwritten like accumulate = 0).  afb_priority_rows_set marks the rows >= first_priority_row (the interface rows of other
 * ranks, which come last in the local row space) as priority; -1 clears.  afb_assemble_phase(.., phase = 1) enqueues the
 * coefficient kernel and the part of the gather that produces the priority rows and returns WITHOUT synchronising: the caller
 * starts its exchange on those rows; phase = 2 (same forms, same buffers) enqueues the rest, synchronises and returns the
 * status.  When the phased cluster gather does not apply, phase 1 does the whole assembly and phase 2 only returns its status. */
int afb_priority_rows_set(afb_ctx* ctx, int64_t first_priority_row);
int afb_assemble_phase(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs_forms, const afb_form* rhs_forms, double* csr_val,
                       double* rhs, double drop_val, int phase);

/* ---- surface terms: fem3Dface inside the local assembler (examples/Fem/Ani/diffusion.cpp:215-245, lin_elast.cpp:198-217) ---- */
/* The boundary faces that carry Neumann / Robin data: face face_num[b] (0..3, int_face.h:49-55) of mesh element face_tet[b].
 * Copies the list; the row lists are built at the first afb_assemble_faces after a pattern change. nbf = 0 clears. */
int afb_boundary_set(afb_ctx* ctx, int64_t nbf, const int32_t* face_tet, const int32_t* face_num, int mem_space);
/* Adds the surface forms int_f (D OpA(u)) . OpB(v) over the listed faces INTO csr_val / rhs (call after afb_assemble, which
 * produces the volume part; the reference's local assembler adds the face matrices to the cell matrix before the scatter).
 * Forms as in afb_assemble with the triangle rule of quad_order; coefficient layouts: CONST, PER_TET = one record per listed
 * face b, PER_POINT = D[len*(n + q*b)] over the points of the triangle rule.  drop_val as in afb_assemble.  Matrix forms need
 * csr_val.  With essential conditions set (afb_dirichlet_set) the faces get the free-row part of applyDir: Dirichlet rows stay
 * untouched, Dirichlet columns move to rhs.  Deterministic (every row sums its faces in ascending b).
 * Returns 0 / -1 (non-finite value). */
int afb_assemble_faces(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs_forms, const afb_form* rhs_forms,
                       double* csr_val, double* rhs, double drop_val, int mem_space);

/* Multi-GPU interface rows: dst[slot[k]] += contrib[k] for the n contributions received from ONE peer (device
 * pointers; the slots of one call are distinct, peers are applied in rank order => deterministic, no atomics).
 * Replaces the value exchange the reference avoids by recomputing ghost cells (assembler.inl:162-183). */
int afb_halo_add(afb_ctx* ctx, int64_t n, const int64_t* slot, const double* contrib, double* dst);

/* ---- multi-GPU: communicator and interface exchange inside the library ("afb_ctx_create(device, nccl_comm, ...)",
 * "afb_halo_exchange(ctx)").  One process (MPI rank) per GPU.  The reference exchanges only the numbering under MPI
 * (global_enumerator.cpp:594-604, :698, :754) and recomputes ghost cells (assembler.inl:162-183); here every rank assembles its own
 * elements into (owned rows ++ interface rows of other ranks) and the interface contributions travel to their owners over NCCL.
 * NCCL is loaded at run time (dlopen "libnccl.so.2", override with AFB_NCCL_LIB); errors of NCCL return -8. */
#define AFB_COMM_ID_BYTES 128
/* ncclGetUniqueId: rank 0 calls it, the application hands the 128 bytes to every rank (MPI_Bcast in an MPI code) */
int afb_comm_unique_id(void* id128);
/* ncclCommInitRank on the context's device (collective over the nranks contexts); the context owns the communicator */
int afb_comm_init(afb_ctx* ctx, const void* id128, int rank, int nranks);
/* adopt an ncclComm_t owned by the application instead */
int afb_comm_set(afb_ctx* ctx, void* nccl_comm, int rank, int nranks);
/* Exchange plan of the extended row space [owned rows 0..n_own) ++ [interface rows of other ranks, sorted by owner]: the values of
 * the interface rows start at entry nnz_own of the extended CSR arrays and are sent to their owners as they lie (send_val[p] values,
 * send_rhs[p] rhs entries for rank p); recv_val[p] / recv_rhs[p] entries arrive from rank p and are added at val_slots / rhs_slots
 * (positions in the extended arrays, concatenated over the peers in rank order; mem_space says where the two slot arrays live). */
int afb_halo_plan_set(afb_ctx* ctx, int nranks, int64_t n_own, int64_t nnz_own, const int64_t* send_val, const int64_t* send_rhs,
                      const int64_t* recv_val, const int64_t* recv_rhs, const int64_t* val_slots, const int64_t* rhs_slots, int mem_space);
/* grouped ncclSend / ncclRecv on the context's communication stream, ordered after the work already issued on the context's
 * stream; _finish orders the context's stream after the transfers and adds the received contributions peer by peer in rank order
 * (deterministic, no atomics).  val_ext / rhs_ext: device pointers, either may be NULL.  afb_halo_exchange = start + finish. */
int afb_halo_exchange_start(afb_ctx* ctx, double* val_ext, double* rhs_ext);
int afb_halo_exchange_finish(afb_ctx* ctx, double* val_ext, double* rhs_ext);
int afb_halo_exchange(afb_ctx* ctx, double* val_ext, double* rhs_ext);
/* One assembly of a partitioned problem: interface rows first (afb_priority_rows_set(ctx, n_own)), their exchange overlapped with
 * the remaining clusters, additions in rank order.  Device pointers; returns 0 / -1 like afb_assemble, results complete in stream
 * order of the context's stream. */
int afb_assemble_distributed(afb_ctx* ctx, int nforms, const afb_form* forms, int nrhs_forms, const afb_form* rhs_forms,
                             double* val_ext, double* rhs_ext, double drop_val);

/* phase times of the last afb_assemble in ms (CUDA events on the context stream):
 * [0] element kernels (k_element_generic / k_geom), [1] gather/scatter (k_gather / k_gather_tensor / k_rows_cl),
 * [2] coefficient copies, [3] = path: 0 generic staged (k_element_generic + k_gather), 1 fused tensor representation with the
 * lane-group gather (k_geom + k_gather_tensor), 2 fused with the cluster-tiled thread-per-row gather (k_geom + k_rows_cl);
 * mirrors the GetTimeEvalLocFunc / GetTimeFillMapTemplate style getters (assembler.inl:949-964) */
int afb_last_times(afb_ctx* ctx, double* ms4);
/* "element kernel|gather kernel" of the generic staged path that ran last (k_element_generic / k_element_sq / k_element_mma,
 * k_gather / k_gather_cols (+ k_gather) / k_gather_flat): evidence for the bench lines */
int afb_last_kernels(afb_ctx* ctx, char* buf, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* ANIFEM_B200_H */
