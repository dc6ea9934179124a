// anifem_b200/memory.hpp -- the memory views and planners of the reference API (fem/fem_memory.h:262-520), host side.
//
// In the reference every fem3Dtet call carves its scratch (XYG, PSI, U, V, DU ...) out of memory the CALLER provides: a
// PlainMemory / PlainMemoryX view of one raw block sized by fem3Dtet_memory_requirements, or parts drawn from a DynMem pool.
// Here the scratch of the element kernels lives on the device, so the requirement of the fem3D* calls is zero -- but reference
// call sites size, allocate and hand over these objects themselves (`req.enoughRawSize()`, `mem.allocateFromRaw(buf, n)`,
// `wmem.alloc(...)`), so the classes are real: same members, same meaning, usable for the caller's own scratch as well.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace Ani {

template <typename ScalarType>
struct DenseMatrix;

namespace mem_detail {
// first address >= p aligned for T, or nullptr when n objects of T do not fit before `end`
template <typename T>
inline T* carve(char*& p, char* end, std::size_t n) {
    const std::uintptr_t a = reinterpret_cast<std::uintptr_t>(p), al = alignof(T);
    char* q = p + ((al - a % al) % al);
    if (q > end || static_cast<std::size_t>(end - q) < n * sizeof(T)) return nullptr;
    p = q + n * sizeof(T);
    return reinterpret_cast<T*>(q);
}
template <typename T>
inline std::size_t worst_bytes(std::size_t n) { return n ? (alignof(T) - 1) + n * sizeof(T) : 0; }
}  // namespace mem_detail

/// fem_memory.h:262-307: one contiguous block seen as dSize scalars followed by iSize indices
template <typename ScalarType = double, typename IndexType = int>
struct PlainMemory {
    ScalarType* ddata = nullptr;
    IndexType* idata = nullptr;
    std::size_t dSize = 0, iSize = 0;
    PlainMemory() = default;
    PlainMemory(ScalarType* d, IndexType* i, std::size_t ds, std::size_t is) : ddata(d), idata(i), dSize(ds), iSize(is) {}
    /// places the two arrays inside [mem_in, mem_in + mem_sz); returns the first free byte behind them, nullptr if they do not fit
    void* allocateFromRaw(void* mem_in, std::size_t mem_sz, std::size_t dsize, std::size_t isize) {
        char *p = static_cast<char*>(mem_in), *end = p + mem_sz;
        ScalarType* d = nullptr;
        IndexType* i = nullptr;
        if (dsize && !(d = mem_detail::carve<ScalarType>(p, end, dsize))) return nullptr;
        if (isize && !(i = mem_detail::carve<IndexType>(p, end, isize))) return nullptr;
        if (dsize) { ddata = d; dSize = dsize; }
        if (isize) { idata = i; iSize = isize; }
        return p;
    }
    void* allocateFromRaw(void* mem_in, std::size_t mem_sz) { return allocateFromRaw(mem_in, mem_sz, dSize, iSize); }
    /// bytes that are enough for allocateFromRaw whatever the alignment of the raw block
    std::size_t enoughRawSize() const { return mem_detail::worst_bytes<ScalarType>(dSize) + mem_detail::worst_bytes<IndexType>(iSize); }
    bool ge(const PlainMemory& o) const { return dSize >= o.dSize && iSize >= o.iSize; }
    void extend_size(const PlainMemory& o) { dSize = std::max(dSize, o.dSize); iSize = std::max(iSize, o.iSize); }
    void append_size(const PlainMemory& o) { dSize += o.dSize; iSize += o.iSize; }
    /// round-1 spelling kept for existing callers
    void allocateFromPlainMemory(ScalarType* d, IndexType* i) { ddata = d; idata = i; }
};

/// fem_memory.h:309-402: the same with an additional array of mSize matrix views
template <typename ScalarType = double, typename IndexType = int>
struct PlainMemoryX {
    ScalarType* ddata = nullptr;
    IndexType* idata = nullptr;
    DenseMatrix<ScalarType>* mdata = nullptr;
    std::size_t dSize = 0, iSize = 0, mSize = 0;
    PlainMemoryX() = default;
    PlainMemoryX(const PlainMemory<ScalarType, IndexType>& m) : ddata(m.ddata), idata(m.idata), dSize(m.dSize), iSize(m.iSize) {}
    PlainMemory<ScalarType, IndexType> getPlainMemory() const { return PlainMemory<ScalarType, IndexType>(ddata, idata, dSize, iSize); }
    void* allocateFromRaw(void* mem_in, std::size_t mem_sz, std::size_t dsize, std::size_t isize, std::size_t msize);
    void* allocateFromRaw(void* mem_in, std::size_t mem_sz) { return allocateFromRaw(mem_in, mem_sz, dSize, iSize, mSize); }
    std::size_t enoughRawSize() const;
    bool ge(const PlainMemoryX& o) const { return dSize >= o.dSize && iSize >= o.iSize && mSize >= o.mSize; }
    template <typename S1, typename I1>
    void extend_size(const PlainMemoryX<S1, I1>& o) { dSize = std::max(dSize, o.dSize); iSize = std::max(iSize, o.iSize); mSize = std::max(mSize, o.mSize); }
    template <typename S1, typename I1>
    void append_size(const PlainMemoryX<S1, I1>& o) { dSize += o.dSize; iSize += o.iSize; mSize += o.mSize; }
};

/// fem_memory.h:404-520: growing pool; alloc() hands out a part that returns its memory when it goes out of scope.  Blocks are
/// bump-allocated: a block whose parts have all been released is reused from its start; defragment() merges the idle blocks into
/// one block of their total capacity (so that a later, larger request fits without a new allocation).
template <typename ScalarType = double, typename IndexType = int>
struct DynMem {
    struct Block {
        std::vector<ScalarType> d;
        std::vector<IndexType> i;
        std::vector<DenseMatrix<ScalarType>> m;
        std::size_t dused = 0, iused = 0, mused = 0, live = 0;
    };
    struct MemPart {
        PlainMemoryX<ScalarType, IndexType> m_mem;
        MemPart() = default;
        MemPart(const MemPart&) = delete;
        MemPart& operator=(const MemPart&) = delete;
        MemPart(MemPart&& o) noexcept : m_mem(o.m_mem), pool(o.pool), block(o.block) { o.pool = nullptr; o.m_mem = PlainMemoryX<ScalarType, IndexType>(); }
        MemPart& operator=(MemPart&& o) noexcept {
            if (this != &o) { clear(); m_mem = o.m_mem; pool = o.pool; block = o.block; o.pool = nullptr; o.m_mem = PlainMemoryX<ScalarType, IndexType>(); }
            return *this;
        }
        ~MemPart() { clear(); }
        void clear() {
            if (pool) pool->release(block);
            pool = nullptr;
            m_mem = PlainMemoryX<ScalarType, IndexType>();
        }
        PlainMemory<ScalarType, IndexType> getPlainMemory() { return m_mem.getPlainMemory(); }
        PlainMemoryX<ScalarType, IndexType> getPlainMemoryX() { return m_mem; }

    private:
        friend struct DynMem;
        DynMem* pool = nullptr;
        std::size_t block = 0;
    };
    virtual ~DynMem() = default;
    virtual MemPart alloc(std::size_t dsize, std::size_t isize, std::size_t msize) {
        std::size_t b = 0;
        for (; b < blocks.size(); ++b) {
            Block& k = *blocks[b];
            if (k.d.size() - k.dused >= dsize && k.i.size() - k.iused >= isize && k.m.size() - k.mused >= msize) break;
        }
        if (b == blocks.size()) {
            // idle block that can grow without moving anything a live part points to, else a new block
            for (b = 0; b < blocks.size() && blocks[b]->live; ++b) {}
            if (b == blocks.size()) blocks.emplace_back(new Block());
            Block& k = *blocks[b];
            if (k.d.size() < dsize) k.d.resize(dsize);
            if (k.i.size() < isize) k.i.resize(isize);
            if (k.m.size() < msize) k.m.resize(msize);
        }
        Block& k = *blocks[b];
        MemPart r;
        r.pool = this; r.block = b;
        r.m_mem.ddata = dsize ? k.d.data() + k.dused : nullptr;
        r.m_mem.idata = isize ? k.i.data() + k.iused : nullptr;
        r.m_mem.mdata = msize ? k.m.data() + k.mused : nullptr;
        r.m_mem.dSize = dsize; r.m_mem.iSize = isize; r.m_mem.mSize = msize;
        k.dused += dsize; k.iused += isize; k.mused += msize; ++k.live;
        return r;
    }
    MemPart alloc(const PlainMemoryX<ScalarType, IndexType>& req) { return alloc(req.dSize, req.iSize, req.mSize); }
    MemPart alloc(const PlainMemory<ScalarType, IndexType>& req) { return alloc(req.dSize, req.iSize, 0); }
    /// merges the blocks without live parts into one
    virtual void defragment() {
        std::size_t nd = 0, ni = 0, nm = 0, idle = 0;
        for (auto& k : blocks)
            if (!k->live) { nd += k->d.size(); ni += k->i.size(); nm += k->m.size(); ++idle; }
        if (idle < 2) return;
        // live parts address their block by index: idle blocks are emptied in place, the first one receives the total capacity
        bool first = true;
        for (auto& k : blocks) {
            if (k->live) continue;
            if (first) { k->d.assign(nd, ScalarType()); k->i.assign(ni, IndexType()); k->m.assign(nm, DenseMatrix<ScalarType>()); first = false; }
            else { std::vector<ScalarType>().swap(k->d); std::vector<IndexType>().swap(k->i); std::vector<DenseMatrix<ScalarType>>().swap(k->m); }
        }
    }
    /// drops every block (no part may be alive)
    virtual void clear() { blocks.clear(); }
    std::size_t nBlocks() const { return blocks.size(); }
    std::size_t capacityScalars() const { std::size_t n = 0; for (auto& k : blocks) n += k->d.size(); return n; }
    std::size_t liveParts() const { std::size_t n = 0; for (auto& k : blocks) n += k->live; return n; }

private:
    void release(std::size_t b) {
        Block& k = *blocks[b];
        if (k.live && --k.live == 0) k.dused = k.iused = k.mused = 0;
    }
    std::vector<std::unique_ptr<Block>> blocks;   // stable addresses: parts keep pointers into the vectors of their block
};

}  // namespace Ani
