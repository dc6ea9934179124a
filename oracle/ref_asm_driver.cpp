// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// ref_asm_driver.cpp: runs the reference's OWN, UNMODIFIED global assembler
//     anifem++/inmost_interface/assembler.h + assembler.inl       (AssemblerT: PrepareProblem, fill_assemble_templates, Assemble,
//                                                                   AssembleTemplate)
//     anifem++/inmost_interface/global_enumerator.h + .cpp        (GlobEnumeration: the six ASSEMBLING_TYPEs)
//     anifem++/inmost_interface/ordering.h + ordering.inl         (collectConnectivityInfo, createOrderPermutation, ...)
//     anifem++/inmost_interface/elemental_assembler.h + .cpp, func_wrap.h (ElementalAssembler, GenerateElemMatRhs)
// compiled in place from /root/reference on top of oracle/mock_inmost/inmost.h (a serial stand-in for the un-vendored INMOST).
// Nothing from the reference is copied here: this file builds a mesh, sets up an Ani::Assembler exactly like the reference's
// examples do (examples/tutorials/ex1.cpp:83-141), runs it, and hands the results back as flat arrays.  The local assembler is
// the reference's fem3Dtet evaluated form by form (CellRunner of ref_driver.cpp).  It pins oracle/asm_oracle.py at the assembler
// level (tests/test_oracle_golden.py::test_asm_oracle_vs_reference_assembler).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <atomic>

#include "inmost.h"   // oracle/mock_inmost

// AssemblerT::fill_assemble_templates (the reference's own element -> global index computation, assembler.inl:139-184) is a
// private member; the driver calls it directly to read the index tables instead of re-deriving them
#define private public
#include "anifem++/inmost_interface/fem.h"
#undef private

#include "ref_driver.cpp"   // CellRunner / make_runner / RefForm: the reference's fem3Dtet per cell

using namespace Ani;

namespace {

DofT::DofMap helper_of(int fem, int vec) {
#define H(F)                                                      \
    case F:                                                       \
        if (vec == 1) return GenerateHelper<FemFix<F>>();         \
        if (vec == 3) return GenerateHelper<FemVec<3, F>>();      \
        break;
    switch (fem) {
        H(FEM_P0) H(FEM_P1) H(FEM_P2) H(FEM_P3)
    }
#undef H
    throw std::runtime_error("unsupported variable");
}

struct AsmState {
    std::unique_ptr<INMOST::Mesh> mesh;
    std::unique_ptr<Assembler> discr;
    int nloc = 0;
    std::vector<long> rowptr;
    std::vector<int> colind;
    std::vector<double> val, rhs;
    long beg = 0, end = 0;
};
AsmState g_asm;

void matrix_to_csr(const INMOST::Sparse::Matrix& A, long beg, long end, std::vector<long>& rowptr, std::vector<int>& colind, std::vector<double>& val) {
    rowptr.assign(end - beg + 1, 0);
    colind.clear(); val.clear();
    for (long r = beg; r < end; ++r) {
        std::vector<std::pair<unsigned, double>> row;
        for (auto it = A[r].Begin(); it != A[r].End(); ++it) row.push_back({it->first, it->second});
        std::sort(row.begin(), row.end());
        for (auto& e : row) { colind.push_back((int)e.first); val.push_back(e.second); }
        rowptr[r - beg + 1] = (long)colind.size();
    }
}

}  // namespace

extern "C" {

// Sets up the reference assembler on a tetrahedral mesh.  enum_type: GlobEnumeration::ASSEMBLING_TYPE (0 ANITYPE .. 5
// ETDIMBLOCKS).  Writes the number of dofs and, per cell, the signed global index codes (sign * (id + 1), 0 = ghost row) of
// fill_assemble_templates into codesC / codesR [ntet x nloc].  Returns nloc (> 0) or a negative error code.
int refasm_setup(int enum_type, int nvars, const int* fem, const int* vec, long nnode, const double* xyz, long ntet, const long* tets,
                 long* nrows_out, long* codesC, long* codesR) {
    try {
        g_asm.discr.reset();   // the assembler releases its index tags on the mesh: it goes first
        g_asm.mesh.reset();
        g_asm = AsmState();
        g_asm.mesh.reset(new INMOST::Mesh());
        g_asm.mesh->BuildTets(nnode, xyz, ntet, tets);
        g_asm.discr.reset(new Assembler(g_asm.mesh.get()));
        Assembler& discr = *g_asm.discr;
        FemExprDescr fed;
        for (int v = 0; v < nvars; ++v) {
            fed.PushTrialFunc(helper_of(fem[v], vec[v]), "u" + std::to_string(v));
            fed.PushTestFunc(helper_of(fem[v], vec[v]), "phi_u" + std::to_string(v));
        }
        discr.SetProbDescr(std::move(fed));
        discr.m_enum.setAssemblingType(static_cast<GlobEnumeration::ASSEMBLING_TYPE>(enum_type));
        const int nloc = discr.m_info.TrialFuncs().NumDofOnTet();
        g_asm.nloc = nloc;
        // an evaluator must exist for PrepareProblem; the real one is installed by refasm_assemble
        std::function<void(const double**, double*, double*, void*)> dummy = [](const double**, double*, double*, void*) {};
        discr.SetMatRHSFunc(GenerateElemMatRhs(dummy, nloc, nloc));
        discr.SetDataGatherer([](ElementalAssembler& p) {
            double* nn_p = p.get_nodes();
            const double* args[] = {nn_p, nn_p + 3, nn_p + 6, nn_p + 9};
            p.compute(args);
        });
        discr.PrepareProblem();
        g_asm.beg = discr.getBegInd(); g_asm.end = discr.getEndInd();
        if (nrows_out) *nrows_out = discr.m_enum.getMatrixSize();
        if (codesC || codesR) {
            INMOST::ElementArray<INMOST::Node> nodes(g_asm.mesh.get(), 4);
            INMOST::ElementArray<INMOST::Edge> edges(g_asm.mesh.get(), 6);
            INMOST::ElementArray<INMOST::Face> faces(g_asm.mesh.get(), 4);
            std::vector<long> iC, iR;
            const bool comp_node_perm = !discr.m_enum.areVarsTriviallySymmetric() ||
                                        (discr.m_info.TestFuncs().GetGeomMask() & (DofT::EDGE_ORIENT | DofT::FACE_ORIENT));
            for (long e = 0; e < ntet; ++e) {
                INMOST::Cell cell = g_asm.mesh->CellByLocalID((int)e);
                collectConnectivityInfo(cell, nodes, edges, faces, discr.m_assm_traits.reorder_nodes, true);
                std::array<unsigned char, 4> cni{0, 1, 2, 3};
                if (comp_node_perm) {   // exactly the lines of the cell loop (assembler.inl:357-364)
                    std::array<long, 4> gni;
                    for (int i = 0; i < 4; ++i) gni[i] = discr.m_enum.GNodeIndex(nodes[i]);
                    cni = createOrderPermutation(gni.data());
                }
                discr.fill_assemble_templates(nodes, edges, faces, cell, iC, iR, cni.data());
                for (int i = 0; i < nloc; ++i) {
                    if (codesC) codesC[e * nloc + i] = iC[i];
                    if (codesR) codesR[e * nloc + i] = iR[i];
                }
            }
        }
        return nloc;
    } catch (std::exception& ex) { g_err = ex.what(); return -4; }
}

// AssembleTemplate (assembler.inl:589-695): structural pattern.  Returns nnz, fills rowptr [nrows+1]; colind via refasm_get.
long refasm_template() {
    try {
        INMOST::Sparse::Matrix A("A");
        const int st = g_asm.discr->AssembleTemplate(A);
        if (st < 0) return st;
        matrix_to_csr(A, g_asm.beg, g_asm.end, g_asm.rowptr, g_asm.colind, g_asm.val);
        g_asm.rhs.clear();
        return (long)g_asm.colind.size();
    } catch (std::exception& ex) { g_err = ex.what(); return -4; }
}

// Assemble (assembler.inl:313-488) with the forms evaluated by the reference's fem3Dtet.  mode bit 0: start from the template
// pattern (is_mtx_include_template), bit 1: use_ordered_insert, bit 2: is_mtx_sorted.  Returns the reference's status (0, -1, ...);
// *nnz_out = entries of the resulting matrix (sorted CSR via refasm_get).
int refasm_assemble(int nforms, const RefForm* forms, double drop_val, int mode, const long* cell_order /*NULL or permutation for D*/,
                    long* nnz_out) {
    try {
        Assembler& discr = *g_asm.discr;
        const int nloc = g_asm.nloc;
        std::vector<std::unique_ptr<CellRunner>> runners;
        for (int k = 0; k < nforms; ++k) {
            const RefForm& f = forms[k];
            runners.push_back(make_runner(f.opA, f.femA, f.vecA, f.opB, f.femB, f.vecB, f.order, f.ttype, f.layout, f.D));
            if (!runners.back()) { g_err = "unsupported operator/space"; return -3; }
        }
        std::vector<double> blk((size_t)nloc * nloc);
        struct Ud { long cell; };
        // local assembler with the signature of the reference's examples: (XY[4], A, F, user_data)
        std::function<void(const double**, double*, double*, void*)> local_assembler =
            [&](const double** XY, double* Adat, double* Fdat, void* user_data) {
                const long e = static_cast<Ud*>(user_data)->cell;
                std::fill(Adat, Adat + (size_t)nloc * nloc, 0.0);
                std::fill(Fdat, Fdat + nloc, 0.0);
                for (int k = 0; k < nforms; ++k) {
                    const RefForm& f = forms[k];
                    CellRunner& r = *runners[k];
                    r.run(XY[0], XY[1], XY[2], XY[3], e, blk.data());
                    if (f.is_rhs) for (long ib = 0; ib < r.nfb; ++ib) Fdat[f.row_off + ib] += f.alpha * blk[ib];
                    else
                        for (long ia = 0; ia < r.nfa; ++ia)
                            for (long ib = 0; ib < r.nfb; ++ib) Adat[(f.row_off + ib) + (size_t)nloc * (f.col_off + ia)] += f.alpha * blk[ib + r.nfb * ia];
                }
            };
        (void)cell_order;
        discr.SetMatRHSFunc(GenerateElemMatRhs(local_assembler, nloc, nloc));
        discr.SetDataGatherer([](ElementalAssembler& p) {
            double* nn_p = p.get_nodes();
            const double* args[] = {nn_p, nn_p + 3, nn_p + 6, nn_p + 9};
            Ud ud{p.cell->LocalID()};
            p.compute(args, &ud);
        });
        discr.PrepareProblem();
        INMOST::Sparse::Matrix A("A");
        INMOST::Sparse::Vector b("b");
        AssmOpts opts;
        opts.SetDropVal(drop_val);
        if (mode & 1) { discr.AssembleTemplate(A); opts.SetIsMtxIncludeTemplate(true).SetIsMtxSorted(true); }
        if (mode & 2) opts.SetUseOrderedInsert(true);
        if (mode & 4) opts.SetIsMtxSorted(true);
        const int st = discr.Assemble(A, b, opts);
        matrix_to_csr(A, g_asm.beg, g_asm.end, g_asm.rowptr, g_asm.colind, g_asm.val);
        g_asm.rhs.assign(g_asm.end - g_asm.beg, 0.0);
        for (long r = g_asm.beg; r < g_asm.end; ++r) g_asm.rhs[r - g_asm.beg] = b[r];
        if (nnz_out) *nnz_out = (long)g_asm.colind.size();
        return st;
    } catch (std::exception& ex) { g_err = ex.what(); return -4; }
}

int refasm_get(long* rowptr, int* colind, double* val, double* rhs) {
    if (rowptr) std::copy(g_asm.rowptr.begin(), g_asm.rowptr.end(), rowptr);
    if (colind) std::copy(g_asm.colind.begin(), g_asm.colind.end(), colind);
    if (val) std::copy(g_asm.val.begin(), g_asm.val.end(), val);
    if (rhs) std::copy(g_asm.rhs.begin(), g_asm.rhs.end(), rhs);
    return 0;
}

void refasm_clear() { g_asm.discr.reset(); g_asm.mesh.reset(); g_asm = AsmState(); }

}  // extern "C"
