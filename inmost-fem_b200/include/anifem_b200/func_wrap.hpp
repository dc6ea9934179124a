// anifem_b200/func_wrap.hpp -- the element-evaluator plug-in point of the reference on top of the C ABI.
//
// Mirrors inmost_interface/func_wrap.h:96-187 (MatFuncWrap), :236-309 (MatFuncWrapDynamic), :310-347 (GenerateElemMat[Rhs] /
// GenerateElemRhs) and the part of inmost_interface/elemental_assembler.h:118-253 (ElementalAssembler) a data gatherer touches:
// a user keeps the local assembler lambda of the reference's examples unchanged (examples/tutorials/ex1.cpp:83-106)
//     std::function<void(const double** XY, double* A, double* F, void* user_data)>
// and installs it with Assembler::SetMatRHSFunc(GenerateElemMatRhs(lambda, nRow, nCol)) (assembler.h:326-328).  The Assembler
// (assembler.hpp) then evaluates it per cell on std::threads (the reference's ThreadPar::parallel_for<STD>, fem/mutex_type.h:110-131)
// into a staged buffer and scatters the local matrices on the GPU (afb_assemble_elemental).  This is the compatibility path:
// the host callback is the bottleneck by construction; problems described with AddMatForm run the fused GPU kernels instead.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

namespace Ani {

/// func_wrap.h:16-88: the local matrices of this path are dense; the view only reports sizes
template <typename Int = long>
struct MatSparsityView {
    enum Type { DENSE = 0, SPARSE_CSC = 1 };
    Type m_type = DENSE;
    Int m_nnz = 0, m_sz1 = 0, m_sz2 = 0;
    MatSparsityView() = default;
    MatSparsityView(Int sz1, Int sz2) : m_nnz(sz1 * sz2), m_sz1(sz1), m_sz2(sz2) {}
};

/// func_wrap.h:96-187
template <typename TReal = double, typename TInt = long>
struct MatFuncWrap {
    using Real = TReal;
    using Int = TInt;
    struct Memory {
        Int* m_iw = nullptr;
        Real* m_w = nullptr;
        const Real** m_args = nullptr;
        Real** m_res = nullptr;
        void* user_data = nullptr;
        Int mem_id = -1;
    };
    virtual ~MatFuncWrap() = default;
    /// args: first n_in() pointers are the inputs (4 pointers to the xyz of the vertices, func_wrap.inl:129); res: first n_out()
    /// pointers receive the result matrices (res[0] = A column-major nRow x nCol, res[1] = F); returns 0 on success
    virtual int operator()(const Real** args, Real** res, Real* w = nullptr, Int* iw = nullptr, void* user_data = nullptr, Int mem_id = -1) const = 0;
    int operator()(Memory mem) const { return operator()(mem.m_args, mem.m_res, mem.m_w, mem.m_iw, mem.user_data, mem.mem_id); }
    virtual void working_sizes(std::size_t& sz_args, std::size_t& sz_res, std::size_t& sz_w, std::size_t& sz_iw) const {
        sz_args = n_in(); sz_res = n_out(); sz_w = sz_iw = 0;
    }
    virtual bool is_user_data_required() const { return false; }
    virtual std::size_t n_in() const { return 0; }
    virtual std::size_t n_out() const = 0;
    virtual MatSparsityView<Int> out_sparsity(Int res_id) const = 0;
    virtual Int out_nnz(Int res_id) const { return out_sparsity(res_id).m_nnz; }
    virtual Int out_size1(Int res_id) const { return out_sparsity(res_id).m_sz1; }
    virtual Int out_size2(Int res_id) const { return out_sparsity(res_id).m_sz2; }
    virtual bool isValid() const { return false; }
    operator bool() const { return isValid(); }
};

/// func_wrap.h:236-309: evaluator built from a std::function
template <typename TReal = double, typename TInt = long>
struct MatFuncWrapDynamic : public MatFuncWrap<TReal, TInt> {
    using Real = TReal;
    using Int = TInt;
    using Functor = std::function<int(const Real** args, Real** res, Real* w, Int* iw, void* user_data)>;
    int operator()(const Real** args, Real** res, Real* w = nullptr, Int* iw = nullptr, void* user_data = nullptr, Int mem_id = -1) const override {
        (void)mem_id;
        return m_f(args, res, w, iw, user_data);
    }
    void working_sizes(std::size_t& sz_args, std::size_t& sz_res, std::size_t& sz_w, std::size_t& sz_iw) const override {
        sz_args = m_n_in; sz_res = m_out.size(); sz_w = m_sz_w; sz_iw = m_sz_iw;
    }
    bool is_user_data_required() const override { return m_user_data_required; }
    std::size_t n_in() const override { return m_n_in; }
    std::size_t n_out() const override { return m_out.size(); }
    MatSparsityView<Int> out_sparsity(Int res_id) const override { return m_out.at(res_id); }
    bool isValid() const override { return static_cast<bool>(m_f); }
    MatFuncWrapDynamic() = default;
    MatFuncWrapDynamic(Functor f, std::size_t n_in, std::vector<MatSparsityView<Int>> out, std::size_t sz_w = 0, std::size_t sz_iw = 0, bool with_user_data = true)
        : m_f(std::move(f)), m_n_in(n_in), m_out(std::move(out)), m_sz_w(sz_w), m_sz_iw(sz_iw), m_user_data_required(with_user_data) {}
    Functor m_f;
    std::size_t m_n_in = 4;
    std::vector<MatSparsityView<Int>> m_out;
    std::size_t m_sz_w = 0, m_sz_iw = 0;
    bool m_user_data_required = true;
};

/// func_wrap.h:310-321: matrix + rhs evaluators
inline MatFuncWrapDynamic<> GenerateElemMatRhs(std::function<void(const double** XY, double* A, double* F, double* w, long* iw, void* user_data)> f,
                                               std::size_t nRow, std::size_t nCol, std::size_t nw, std::size_t niw) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double* w, long* iw, void* ud) { f(args, res[0], res[1], w, iw, ud); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, (long)nCol), MatSparsityView<long>((long)nRow, 1)}, nw, niw, true);
}
inline MatFuncWrapDynamic<> GenerateElemMatRhs(std::function<void(const double** XY, double* A, double* F, void* user_data)> f, std::size_t nRow, std::size_t nCol) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void* ud) { f(args, res[0], res[1], ud); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, (long)nCol), MatSparsityView<long>((long)nRow, 1)}, 0, 0, true);
}
inline MatFuncWrapDynamic<> GenerateElemMatRhs(std::function<void(const double** XY, double* A, double* F)> f, std::size_t nRow, std::size_t nCol) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void*) { f(args, res[0], res[1]); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, (long)nCol), MatSparsityView<long>((long)nRow, 1)}, 0, 0, false);
}
/// func_wrap.h:323-334: matrix only
inline MatFuncWrapDynamic<> GenerateElemMat(std::function<void(const double** XY, double* A, void* user_data)> f, std::size_t nRow, std::size_t nCol) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void* ud) { f(args, res[0], ud); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, (long)nCol)}, 0, 0, true);
}
inline MatFuncWrapDynamic<> GenerateElemMat(std::function<void(const double** XY, double* A)> f, std::size_t nRow, std::size_t nCol) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void*) { f(args, res[0]); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, (long)nCol)}, 0, 0, false);
}
/// func_wrap.h:336-347: rhs only
inline MatFuncWrapDynamic<> GenerateElemRhs(std::function<void(const double** XY, double* F, void* user_data)> f, std::size_t nRow) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void* ud) { f(args, res[0], ud); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, 1)}, 0, 0, true);
}
inline MatFuncWrapDynamic<> GenerateElemRhs(std::function<void(const double** XY, double* F)> f, std::size_t nRow) {
    return MatFuncWrapDynamic<>([f](const double** args, double** res, double*, long*, void*) { f(args, res[0]); return 0; }, 4,
                                {MatSparsityView<long>((long)nRow, 1)}, 0, 0, false);
}

/// What a data gatherer sees of one cell (elemental_assembler.h:118-253): the vertex coordinates of the positively oriented
/// tetrahedron (update(), elemental_assembler.cpp:105-117), the ids of the cell and its nodes (the INMOST handles of the reference:
/// `(*p.nodes)[i]` there, `p.node_ids[i]` here), and compute() which runs the installed evaluator into the local A / F.
class ElementalAssembler {
public:
    double* get_nodes() { return m_nn_p.data(); }
    const double* get_nodes() const { return m_nn_p.data(); }
    /// elemental_assembler.cpp:93-103: args[k] -> xyz of vertex k; user_data is handed to the evaluator
    void compute(const double** args, void* user_data = nullptr) {
        double* res[2] = {nullptr, nullptr};
        std::size_t k = 0;
        if (m_A) res[k++] = m_A;
        if (m_F) res[k++] = m_F;
        if ((*m_func)(args, res, m_w, m_iw, user_data, m_thread) != 0) throw std::runtime_error("local evaluator returned an error");
    }
    void compute(const double** args, void* user_data, std::nullptr_t) { compute(args, user_data); }
    int64_t cell_id = 0;                 ///< cell (local id in the mesh given to SetMesh)
    std::array<int64_t, 4> node_ids{};   ///< its nodes, positively oriented (ordering.inl:8-26)
    int nRows = 0, nCols = 0;            ///< local matrix sizes
    void* user_data = nullptr;           ///< AssmOpts::user_data of the running Assemble (assembler.h:211)
    // -- set by the Assembler
    const MatFuncWrap<>* m_func = nullptr;
    double *m_A = nullptr, *m_F = nullptr, *m_w = nullptr;
    long* m_iw = nullptr;
    long m_thread = 0;
    std::array<double, 12> m_nn_p{};
};

}  // namespace Ani
