// anifem_b200/assembler.hpp -- C++ mirror of the AniFem++ global assembler on top of the C ABI.
//
// Mirrors inmost_interface/assembler.h:210-456 (AssemblerT / AssmOpts) for the path in scope.  What the reference pulls
// out of an INMOST mesh per cell is passed in as flat arrays; what the reference's user lambda computes with
// fem3Dtet calls is described as a list of volume forms (one entry = one fem3Dtet<OpA,OpB>(..) call whose result
// the lambda adds into the block (row_off, col_off) of the local matrix, e.g. examples/Fem/Ani/stokes.cpp:138-185).
//
//   reference                                    here
//   -------------------------------------------  --------------------------------------------------------------
//   SetMesh(INMOST::Mesh*)                       SetMesh(nnode, x, y, z, ntet, v0..v3)  |  SetCubeMesh(nx, ny, nz)
//   SetProbDescr(FemExprDescr) + m_enum(NATURAL) SetProbDescr({{FEM_P2, 1}, ...})  or  SetDofMap(explicit index codes)
//   SetMatRHSFunc(GenerateElemMatRhs(lambda))    AddMatForm<OpA,OpB>(...), AddRhsForm<OpB>(...)
//   PrepareProblem()                             PrepareProblem()   (numbering + pattern + gather plan on the device)
//   AssembleTemplate(Matrix&)                    AssembleTemplate(CsrMatrix&)
//   Assemble(Matrix&, Vector&, AssmOpts)         Assemble(CsrMatrix&, std::vector<double>&, AssmOpts)  -> 0 / -1
//   AssembleMatrix / AssembleRHS                 AssembleMatrix / AssembleRHS
//   getBegInd() / getEndInd()                    getBegInd() / getEndInd()
//   applyDir(A, F, k, bc) inside the lambda      SetDirichlet(flag per dof, value per dof)
// The matrix always has the structural pattern of AssembleTemplate with rows sorted ascending, i.e. the reference's
// opts.is_mtx_include_template = opts.use_ordered_insert = true mode (assembler.inl:428-438); Assemble ADDS into it.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <vector>

#include <thread>

#include "enumerator.hpp"
#include "fem.hpp"
#include "func_wrap.hpp"

namespace Ani {

/// assembler.h:210-231
struct AssmOpts {
    void* user_data = nullptr;       ///< handed to the data gatherer / local evaluator of the MatFuncWrap path
    double drop_val = 1e-100;
    bool is_mtx_sorted = true;
    bool is_mtx_include_template = true;
    bool use_ordered_insert = true;
    AssmOpts& SetDropVal(double v) { drop_val = v; return *this; }
    AssmOpts& SetUserData(void* p) { user_data = p; return *this; }
};

/// stand-in for INMOST::Sparse::Matrix over [BegInd, EndInd): sorted CSR
struct CsrMatrix {
    int64_t row_begin = 0, row_end = 0;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> colind;
    std::vector<double> val;
    bool empty() const { return rowptr.empty(); }
};

struct FemVarDescr { int fem; int vec; };

class Assembler {
public:
    /// assembler.h:297-302
    struct AssembleMode {
        bool reorder_nodes = true;
        bool prepare_edges = true;
        bool prepare_faces = true;
        int num_threads = -1;   ///< host threads evaluating the local assembler of the MatFuncWrap path; < 0 = all cores
    };
    AssembleMode m_assm_traits;

    explicit Assembler(int device = 0) {
        if (afb_ctx_create(device, nullptr, &m_ctx) != 0) throw std::runtime_error(std::string("anifem_b200: ") + afb_last_error(nullptr));
    }
    ~Assembler() { afb_ctx_destroy(m_ctx); }
    Assembler(const Assembler&) = delete;
    Assembler& operator=(const Assembler&) = delete;

    Assembler& SetMesh(int64_t nnode, const double* x, const double* y, const double* z, int64_t ntet, const int32_t* v0, const int32_t* v1,
                       const int32_t* v2, const int32_t* v3, bool reorder_nodes = true) {
        ck(afb_mesh_set(m_ctx, nnode, x, y, z, ntet, v0, v1, v2, v3, AFB_HOST));
        if (reorder_nodes) ck(afb_mesh_orient(m_ctx));  // AssembleMode::reorder_nodes (assembler.h:297-302)
        m_has_mesh = true; m_prepared = false;
        return *this;
    }
    /// utils/mesh_utils.h:11-17 GenerateCube
    Assembler& SetCubeMesh(int nx, int ny, int nz, double size = 1.0) {
        ck(afb_mesh_cube(m_ctx, nx, ny, nz, size, 0, 0, 0, nx, ny, nz));
        m_has_mesh = true; m_prepared = false;
        return *this;
    }
    Assembler& SetProbDescr(std::vector<FemVarDescr> vars) { m_vars = std::move(vars); m_explicit = false; m_prepared = false; return *this; }
    /// SetEnumerator(ASSEMBLING_TYPE) (assembler.h:333-336, global_enumerator.h:393-401): NATURAL (default) is numbered on the
    /// device; the other types are numbered on the host (enumerator.hpp, as the reference does) and go in through afb_dofmap_set
    Assembler& SetEnumerator(ASSEMBLING_TYPE t) { m_enum_type = t; m_prepared = false; return *this; }
    /// explicit elem -> global index codes (m_indexesR / m_indexesC, assembler.inl:49-55,139-184)
    Assembler& SetDofMap(int nrow_loc, int ncol_loc, const int64_t* rowcode, const int64_t* colcode, int64_t row_begin, int64_t row_end, int64_t ncols) {
        if (!m_has_mesh) throw std::runtime_error("Mesh was not specified");
        ck(afb_dofmap_set(m_ctx, nrow_loc, ncol_loc, rowcode, colcode, row_begin, row_end, ncols, AFB_HOST));
        m_explicit = true; m_prepared = false;
        return *this;
    }
    /// local offset of variable v inside the element vector (variable-major order, fem/tetdofmap.cpp:606-643)
    int VarOffset(int v) const {
        int o = 0;
        for (int k = 0; k < v; ++k) o += b200_detail::base_nf(m_vars[k].fem) * m_vars[k].vec;
        return o;
    }
    /// The element-evaluator plug-in point of the reference (assembler.h:326-328, func_wrap.h:310-347): a host callback per cell.
    /// Compatibility path: the callback runs on host threads, the scatter on the GPU (afb_assemble_elemental).
    Assembler& SetMatFunc(MatFuncWrapDynamic<> f) { mat_func = std::make_shared<MatFuncWrapDynamic<>>(std::move(f)); return *this; }
    Assembler& SetRHSFunc(MatFuncWrapDynamic<> f) { rhs_func = std::make_shared<MatFuncWrapDynamic<>>(std::move(f)); return *this; }
    Assembler& SetMatRHSFunc(MatFuncWrapDynamic<> f) { mat_rhs_func = std::make_shared<MatFuncWrapDynamic<>>(std::move(f)); return *this; }
    /// assembler.h:324: the handler that gathers per-cell data and calls p.compute(args, user_data); default = ex1.cpp:108-118
    /// without the label lookup: args = the four vertices, user_data = AssmOpts::user_data
    Assembler& SetDataGatherer(std::function<void(ElementalAssembler&)> h) { m_prob_handler = std::move(h); return *this; }
    std::shared_ptr<MatFuncWrap<>> mat_func, rhs_func, mat_rhs_func;
    std::function<void(ElementalAssembler&)> m_prob_handler;

    /// one fem3Dtet<OpA,OpB> term of the local matrix; D in the user-callback layout (col-major Dim(OpB) x Dim(OpA))
    template <typename OpA, typename OpB>
    Assembler& AddMatForm(int trial_var, int test_var, int order, TensorType ttype, int coef_layout, const double* D, double alpha = 1.0,
                          int coef_space = AFB_HOST) {
        m_forms.push_back(afb_form{OpA::op, OpA::fem, OpA::vec, OpB::op, OpB::fem, OpB::vec, order, ttype, coef_layout, coef_space, D, alpha,
                                   VarOffset(test_var), VarOffset(trial_var)});
        return *this;
    }
    /// rhs term int f . OpB(v): the reference's fem3Dtet<Operator<IDEN,FemFix<FEM_P0>>, OpB> idiom (ex1.cpp:92-95)
    template <typename OpB>
    Assembler& AddRhsForm(int test_var, int order, TensorType ttype, int coef_layout, const double* D, double alpha = 1.0, int coef_space = AFB_HOST) {
        m_rhs.push_back(afb_form{IDEN, FEM_P0, 1, OpB::op, OpB::fem, OpB::vec, order, ttype, coef_layout, coef_space, D, alpha, VarOffset(test_var), 0});
        return *this;
    }
    /// Surface terms: the fem3Dface<OpA,OpB>(XYZ, face, ...) calls a reference local assembler makes on the faces with a boundary
    /// label (examples/Fem/Ani/diffusion.cpp:215-245).  The labelled faces are given once as (cell, face number 0..3); the
    /// coefficient layouts are CONST, AFB_COEF_PER_TET = one record per listed face, PER_POINT = per point of the triangle rule.
    Assembler& SetBoundaryFaces(const std::vector<int32_t>& face_cell, const std::vector<int32_t>& face_num) {
        if (face_cell.size() != face_num.size()) throw std::runtime_error("SetBoundaryFaces: arrays differ in size");
        ck(afb_boundary_set(m_ctx, (int64_t)face_cell.size(), face_cell.data(), face_num.data(), AFB_HOST));
        return *this;
    }
    template <typename OpA, typename OpB>
    Assembler& AddFaceMatForm(int trial_var, int test_var, int order, TensorType ttype, int coef_layout, const double* D, double alpha = 1.0,
                              int coef_space = AFB_HOST) {
        m_fforms.push_back(afb_form{OpA::op, OpA::fem, OpA::vec, OpB::op, OpB::fem, OpB::vec, order, ttype, coef_layout, coef_space, D, alpha,
                                    VarOffset(test_var), VarOffset(trial_var)});
        return *this;
    }
    template <typename OpB>
    Assembler& AddFaceRhsForm(int test_var, int order, TensorType ttype, int coef_layout, const double* D, double alpha = 1.0, int coef_space = AFB_HOST) {
        m_frhs.push_back(afb_form{IDEN, FEM_P0, 1, OpB::op, OpB::fem, OpB::vec, order, ttype, coef_layout, coef_space, D, alpha, VarOffset(test_var), 0});
        return *this;
    }
    void ClearForms() { m_forms.clear(); m_rhs.clear(); m_fforms.clear(); m_frhs.clear(); }

    /// Essential boundary conditions: what the reference's local assemblers do with applyDir(A, F, k, bc) on every Dirichlet
    /// dof of every cell (fem/operations/dc_on_dof.h:27-45, examples/tutorials/ex1.cpp:96-105).  is_dirichlet / value are
    /// indexed by the global dof; call after PrepareProblem (the numbering must exist).  Empty vectors clear the setting.
    Assembler& SetDirichlet(const std::vector<unsigned char>& is_dirichlet, const std::vector<double>& value) {
        if (!m_prepared) throw std::runtime_error("SetDirichlet: call PrepareProblem first");
        if (is_dirichlet.empty()) { ck(afb_dirichlet_set(m_ctx, nullptr, nullptr, AFB_HOST)); return *this; }
        if (is_dirichlet.size() != value.size()) throw std::runtime_error("SetDirichlet: flag and value arrays differ in size");
        {   // the library copies one entry per global dof: the arrays must cover the whole column space
            int64_t ng = 0;
            ck(afb_dofmap_get(m_ctx, nullptr, nullptr, nullptr, nullptr, &ng, nullptr, nullptr, AFB_HOST));
            if ((int64_t)is_dirichlet.size() != ng) throw std::runtime_error("SetDirichlet: arrays must have one entry per global dof");
        }
        ck(afb_dirichlet_set(m_ctx, is_dirichlet.data(), value.data(), AFB_HOST));
        return *this;
    }

    /// assembler.inl:193-275
    void PrepareProblem() {
        if (!m_has_mesh) throw std::runtime_error("Mesh was not specified");
        if (!m_explicit) {
            if (m_vars.empty()) throw std::runtime_error("Description of fem expression is empty, try SetProbDescr(...)");
            std::vector<int> fem, vec;
            for (auto& v : m_vars) { fem.push_back(v.fem); vec.push_back(v.vec); }
            if (m_enum_type == NATURAL) ck(afb_dofmap_natural(m_ctx, static_cast<int>(fem.size()), fem.data(), vec.data()));
            else {
                int64_t nnode = 0, ntet = 0;
                ck(afb_mesh_get(m_ctx, &nnode, &ntet, nullptr, nullptr, AFB_HOST));
                std::vector<int32_t> v(static_cast<std::size_t>(4) * ntet);
                ck(afb_mesh_get(m_ctx, &nnode, &ntet, nullptr, v.data(), AFB_HOST));
                std::vector<EnumVar> ev;
                for (auto& d : m_vars) ev.push_back(EnumVar{d.fem, d.vec});
                DofEnumeration en = enumerate_dofs(m_enum_type, nnode, ntet, v.data(), v.data() + ntet, v.data() + 2 * ntet, v.data() + 3 * ntet, ev);
                for (auto& c : en.elem2dof) c += 1;   // assemble_index_encode: sign * (id + 1) (assembler.inl:49-55)
                ck(afb_dofmap_set(m_ctx, en.nloc, en.nloc, en.elem2dof.data(), en.elem2dof.data(), 0, en.nrows, en.nrows, AFB_HOST));
            }
        }
        ck(afb_pattern_build(m_ctx, &m_nnz));
        int nr, nc; int64_t ng;
        ck(afb_dofmap_get(m_ctx, &nr, &nc, &m_beg, &m_end, &ng, nullptr, nullptr, AFB_HOST));
        m_prepared = true;
    }
    int64_t getBegInd() const { return m_beg; }
    int64_t getEndInd() const { return m_end; }
    int64_t getNnz() const { return m_nnz; }

    /// assembler.inl:589-695: adds the zero matrix of structural non-zeros (here: defines the pattern)
    int AssembleTemplate(CsrMatrix& matrix) {
        need_prepared();
        matrix.row_begin = m_beg; matrix.row_end = m_end;
        matrix.rowptr.assign(m_end - m_beg + 1, 0);
        matrix.colind.assign(m_nnz, 0);
        ck(afb_pattern_get(m_ctx, matrix.rowptr.data(), matrix.colind.data(), AFB_HOST));
        matrix.val.assign(m_nnz, 0.0);
        return 0;
    }
    /// assembler.inl:313-488. matrix <- matrix + assembled, rhs <- rhs + assembled. Returns 0, or -1 on NaN/Inf.
    int Assemble(CsrMatrix& matrix, std::vector<double>& rhs, const AssmOpts& opts = AssmOpts()) {
        if (mat_rhs_func || (mat_func && rhs_func)) { prepare_outputs(&matrix, &rhs); return assemble_elemental(&matrix, &rhs, opts); }
        if (m_forms.empty() && m_rhs.empty()) throw std::runtime_error("System local evaluator is not specified");
        prepare_outputs(&matrix, &rhs);
        const int st = ck(afb_assemble(m_ctx, (int)m_forms.size(), m_forms.data(), (int)m_rhs.size(), m_rhs.data(), matrix.val.data(), rhs.data(), 1,
                                       opts.drop_val, AFB_HOST));
        return std::min(st, faces(matrix.val.data(), rhs.data(), opts));
    }
    int AssembleMatrix(CsrMatrix& matrix, const AssmOpts& opts = AssmOpts()) {
        if (mat_func || mat_rhs_func) { prepare_outputs(&matrix, nullptr); return assemble_elemental(&matrix, nullptr, opts); }
        if (m_forms.empty()) throw std::runtime_error("Matrix local evaluator is not specified");
        prepare_outputs(&matrix, nullptr);
        const int st = ck(afb_assemble(m_ctx, (int)m_forms.size(), m_forms.data(), 0, nullptr, matrix.val.data(), nullptr, 1, opts.drop_val, AFB_HOST));
        return std::min(st, faces(matrix.val.data(), nullptr, opts));
    }
    int AssembleRHS(std::vector<double>& rhs, const AssmOpts& opts = AssmOpts()) {
        if (rhs_func || mat_rhs_func) { prepare_outputs(nullptr, &rhs); return assemble_elemental(nullptr, &rhs, opts); }
        if (m_rhs.empty()) throw std::runtime_error("Right-hand side local evaluator is not specified");
        prepare_outputs(nullptr, &rhs);
        const int st = ck(afb_assemble(m_ctx, 0, nullptr, (int)m_rhs.size(), m_rhs.data(), nullptr, rhs.data(), 1, opts.drop_val, AFB_HOST));
        return std::min(st, faces(nullptr, rhs.data(), opts));
    }
    /// GetTimeEvalLocFunc-style getters (assembler.inl:949-964): ms of the last Assemble
    double GetTimeEvalLocFunc() const { double t[4]; afb_last_times(m_ctx, t); return t[0]; }
    double GetTimeFillGlobalStructs() const { double t[4]; afb_last_times(m_ctx, t); return t[1]; }
    afb_ctx* context() { return m_ctx; }

private:
    int ck(int rc) {
        if (rc < 0 && rc != -1) throw std::runtime_error(afb_last_error(m_ctx));
        return rc;
    }
    int faces(double* val, double* rhs, const AssmOpts& opts) {
        if (m_fforms.empty() && m_frhs.empty()) return 0;
        return ck(afb_assemble_faces(m_ctx, (int)m_fforms.size(), m_fforms.data(), (int)m_frhs.size(), m_frhs.data(), val, rhs, opts.drop_val, AFB_HOST));
    }
    void need_prepared() {
        if (!m_prepared) PrepareProblem();
    }
    /// The cell loop of AssemblerT::Assemble (assembler.inl:349-483) with the user's evaluator: cell ranges over std::threads
    /// (ThreadPar::parallel_for<STD>, fem/mutex_type.h:110-131), local matrices staged per chunk, scatter on the device.
    int assemble_elemental(CsrMatrix* matrix, std::vector<double>* rhs, const AssmOpts& opts) {
        int nrl = 0, ncl = 0; int64_t rb, re, ng;
        ck(afb_dofmap_get(m_ctx, &nrl, &ncl, &rb, &re, &ng, nullptr, nullptr, AFB_HOST));
        int64_t nnode = 0, ntet = 0;
        ck(afb_mesh_get(m_ctx, &nnode, &ntet, nullptr, nullptr, AFB_HOST));
        std::vector<double> xyz(static_cast<std::size_t>(3) * nnode);
        std::vector<int32_t> v(static_cast<std::size_t>(4) * ntet);
        ck(afb_mesh_get(m_ctx, &nnode, &ntet, xyz.data(), v.data(), AFB_HOST));
        const bool doA = matrix != nullptr, doF = rhs != nullptr;
        // which evaluator fills what: the combined one when present (generate_mat_rhs_func, assembler.inl), else the single ones
        const MatFuncWrap<>* fA = doA ? (mat_rhs_func && (doF || !mat_func) ? mat_rhs_func.get() : mat_func.get()) : nullptr;
        const MatFuncWrap<>* fF = doF ? (mat_rhs_func && (doA || !rhs_func) ? mat_rhs_func.get() : rhs_func.get()) : nullptr;
        const bool combined = (doA && fA == mat_rhs_func.get()) || (doF && fF == mat_rhs_func.get());
        const std::size_t per = static_cast<std::size_t>(nrl) * ncl + nrl;
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ntet, static_cast<int64_t>((std::size_t(512) << 20) / (per * sizeof(double)))));
        int nth = m_assm_traits.num_threads < 0 ? static_cast<int>(std::thread::hardware_concurrency()) : m_assm_traits.num_threads;
        nth = std::max(1, nth);
        std::vector<double> Abuf(static_cast<std::size_t>(chunk) * nrl * ncl), Fbuf(static_cast<std::size_t>(chunk) * nrl);
        int status = 0;
        for (int64_t e_lo = 0; e_lo < ntet; e_lo += chunk) {
            const int64_t nel = std::min<int64_t>(chunk, ntet - e_lo);
            std::vector<std::string> errs(nth);
            auto work = [&](int t) {
                try {
                    std::size_t sa, sr, sw, siw;
                    const MatFuncWrap<>* any = combined ? mat_rhs_func.get() : (fA ? fA : fF);
                    any->working_sizes(sa, sr, sw, siw);
                    std::vector<double> w(sw + 1), Atmp(static_cast<std::size_t>(nrl) * ncl), Ftmp(nrl);
                    std::vector<long> iw(siw + 1);
                    ElementalAssembler p;
                    p.nRows = nrl; p.nCols = ncl; p.m_w = w.data(); p.m_iw = iw.data(); p.m_thread = t; p.user_data = opts.user_data;
                    for (int64_t k = nel * t / nth; k < nel * (t + 1) / nth; ++k) {
                        const int64_t e = e_lo + k;
                        p.cell_id = e;
                        for (int n = 0; n < 4; ++n) {
                            p.node_ids[n] = v[static_cast<std::size_t>(n) * ntet + e];
                            for (int d = 0; d < 3; ++d) p.m_nn_p[3 * n + d] = xyz[static_cast<std::size_t>(d) * nnode + p.node_ids[n]];
                        }
                        double* Ae = Abuf.data() + static_cast<std::size_t>(k) * nrl * ncl;
                        double* Fe = Fbuf.data() + static_cast<std::size_t>(k) * nrl;
                        auto run = [&](const MatFuncWrap<>* f, double* A, double* F) {
                            // ElementalAssembler::update (elemental_assembler.cpp:105-117): local A, F start from zero
                            if (A) std::fill(A, A + static_cast<std::size_t>(nrl) * ncl, 0.0);
                            if (F) std::fill(F, F + nrl, 0.0);
                            p.m_func = f; p.m_A = A; p.m_F = F;
                            if (m_prob_handler) m_prob_handler(p);
                            else {
                                const double* args[4] = {p.m_nn_p.data(), p.m_nn_p.data() + 3, p.m_nn_p.data() + 6, p.m_nn_p.data() + 9};
                                p.compute(args, opts.user_data);
                            }
                        };
                        if (combined) run(mat_rhs_func.get(), doA ? Ae : Atmp.data(), doF ? Fe : Ftmp.data());
                        else {
                            if (fA) run(fA, Ae, nullptr);
                            if (fF) run(fF, nullptr, Fe);
                        }
                    }
                } catch (std::exception& ex) { errs[t] = ex.what(); if (errs[t].empty()) errs[t] = "error in the local evaluator"; }
            };
            if (nth == 1) work(0);
            else {
                std::vector<std::thread> ths;
                for (int t = 0; t < nth; ++t) ths.emplace_back(work, t);
                for (auto& th : ths) th.join();
            }
            for (auto& er : errs) if (!er.empty()) throw std::runtime_error(er);
            const int st = ck(afb_assemble_elemental(m_ctx, e_lo, nel, doA ? Abuf.data() : nullptr, doF ? Fbuf.data() : nullptr, AFB_HOST,
                                                     doA ? matrix->val.data() : nullptr, doF ? rhs->data() : nullptr, opts.drop_val, AFB_HOST));
            status = std::min(status, st);
        }
        return status;
    }
    void prepare_outputs(CsrMatrix* m, std::vector<double>* rhs) {
        need_prepared();
        if (m && (m->empty() || m->row_begin != m_beg || m->row_end != m_end || (int64_t)m->colind.size() != m_nnz)) AssembleTemplate(*m);
        if (rhs && (int64_t)rhs->size() != m_end - m_beg) rhs->assign(m_end - m_beg, 0.0);  // rhs.SetInterval (assembler.inl:323)
    }
    afb_ctx* m_ctx = nullptr;
    std::vector<FemVarDescr> m_vars;
    std::vector<afb_form> m_forms, m_rhs, m_fforms, m_frhs;
    bool m_has_mesh = false, m_explicit = false, m_prepared = false;
    ASSEMBLING_TYPE m_enum_type = NATURAL;
    int64_t m_nnz = 0, m_beg = 0, m_end = 0;
};

}  // namespace Ani
