// Host-side API twins that need no GPU: memory views / planners (fem/fem_memory.h:262-520).  Built and run by the CPU test suite.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "anifem_b200/fem.hpp"

using namespace Ani;
static int fails = 0;
#define EXPECT(c)                                                                  \
    do {                                                                           \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } \
    } while (0)

int main() {
    {   // PlainMemory: the usage pattern of the reference's examples (size -> raw buffer -> views), with a misaligned raw block
        PlainMemory<double, int> req;
        req.dSize = 37; req.iSize = 11;
        PlainMemory<double, int> other;
        other.dSize = 5; other.iSize = 40;
        EXPECT(!req.ge(other));
        req.extend_size(other);
        EXPECT(req.dSize == 37 && req.iSize == 40 && req.ge(other));
        PlainMemory<double, int> sum = req;
        sum.append_size(other);
        EXPECT(sum.dSize == 42 && sum.iSize == 80);
        const std::size_t bytes = req.enoughRawSize();
        EXPECT(bytes >= 37 * sizeof(double) + 40 * sizeof(int) && bytes <= 37 * sizeof(double) + 40 * sizeof(int) + 16);
        std::vector<char> raw(bytes + 1);
        PlainMemory<double, int> mem = req;
        void* end = mem.allocateFromRaw(raw.data() + 1, bytes);   // odd address: the worst case of the estimate
        EXPECT(end != nullptr && mem.ddata && mem.idata);
        EXPECT(reinterpret_cast<std::uintptr_t>(mem.ddata) % alignof(double) == 0 && reinterpret_cast<std::uintptr_t>(mem.idata) % alignof(int) == 0);
        EXPECT(reinterpret_cast<char*>(mem.ddata + 37) <= reinterpret_cast<char*>(mem.idata));
        EXPECT(static_cast<char*>(end) <= raw.data() + 1 + bytes && static_cast<char*>(end) == reinterpret_cast<char*>(mem.idata + 40));
        for (int k = 0; k < 37; ++k) mem.ddata[k] = k;
        for (int k = 0; k < 40; ++k) mem.idata[k] = -k;
        EXPECT(mem.ddata[36] == 36.0 && mem.idata[39] == -39);
        PlainMemory<double, int> small = req;
        EXPECT(small.allocateFromRaw(raw.data(), 64) == nullptr && small.ddata == nullptr);   // does not fit: nothing is touched
        EXPECT(PlainMemory<>().enoughRawSize() == 0);
    }
    {   // PlainMemoryX: scalars, indices and matrix views
        PlainMemoryX<double, int> req;
        req.dSize = 9; req.iSize = 3; req.mSize = 4;
        std::vector<char> raw(req.enoughRawSize());
        PlainMemoryX<double, int> mem = req;
        EXPECT(mem.allocateFromRaw(raw.data(), raw.size()) != nullptr && mem.mdata != nullptr);
        mem.mdata[3] = DenseMatrix<double>(mem.ddata, 3, 3);
        mem.mdata[3](2, 1) = 7.0;
        EXPECT(mem.ddata[2 + 3 * 1] == 7.0 && mem.mdata[0].data == nullptr);
        PlainMemory<double, int> plain = mem.getPlainMemory();
        EXPECT(plain.ddata == mem.ddata && plain.iSize == 3);
    }
    {   // DynMem: parts return their memory at scope exit, idle blocks are reused / merged
        DynMem<double, int> pool;
        double* first = nullptr;
        {
            auto a = pool.alloc(100, 10, 0);
            auto b = pool.alloc(50, 0, 2);
            EXPECT(pool.liveParts() == 2 && a.m_mem.ddata && b.m_mem.ddata && b.m_mem.mdata);
            EXPECT(a.m_mem.ddata + 100 <= b.m_mem.ddata || b.m_mem.ddata + 50 <= a.m_mem.ddata);   // disjoint
            a.m_mem.ddata[99] = 1.0; b.m_mem.ddata[49] = 2.0;
            first = a.m_mem.ddata;
            auto c = std::move(a);                      // ownership moves, nothing is released
            EXPECT(pool.liveParts() == 2 && c.m_mem.ddata == first && a.m_mem.ddata == nullptr);
            PlainMemory<double, int> pm = c.getPlainMemory();
            EXPECT(pm.dSize == 100 && pm.iSize == 10);
        }
        EXPECT(pool.liveParts() == 0);
        const std::size_t nb = pool.nBlocks();
        {
            auto a = pool.alloc(100, 10, 0);            // served from the released capacity: no new block
            EXPECT(pool.nBlocks() == nb && a.m_mem.ddata != nullptr);
        }
        const std::size_t cap = pool.capacityScalars();
        pool.defragment();
        EXPECT(pool.capacityScalars() == cap);
        {
            auto big = pool.alloc(cap, 0, 0);           // the merged block holds the whole capacity
            EXPECT(big.m_mem.ddata != nullptr && pool.capacityScalars() == cap);
        }
        pool.clear();
        EXPECT(pool.nBlocks() == 0);
    }
    {   // the requirement of the element calls is zero here (the scratch lives on the device); no GPU is touched by asking
        auto req = fem3Dtet_memory_requirements<Operator<GRAD, FemFix<FEM_P2>>, Operator<GRAD, FemFix<FEM_P2>>>(5, 4);
        EXPECT(req.enoughRawSize() == 0 && req.dSize == 0 && req.iSize == 0);
    }
    {   // FemSpace::dofMap: the local orders of the supported spaces as dof maps (sizes agree with the element matrices)
        using namespace Ani::DofT;
        for (int fem : {FEM_P0, FEM_P1, FEM_P2, FEM_P3})
            for (int vec : {1, 3}) {
                FemSpace s(fem, vec);
                DofMap m = s.dofMap();
                EXPECT(m.NumDofOnTet() == s.dofMapSize());
                for (uint g = 0; g < m.NumDofOnTet(); ++g) EXPECT(m.TetDofID(m[g].getGeomOrder()) == g);
            }
        DofMap th = FemSpace(FEM_P2, 3).dofMap() * FemSpace(FEM_P1).dofMap();   // Taylor-Hood: 30 velocity dofs, then 4 pressure dofs
        EXPECT(th.NumDofOnTet() == 34 && th[30].etype == NODE && th[30].leid == 3 && th[10].nelem == 0 && th[10].leid == 1);
        TetGeomSparsity face1;
        face1.setFace(1, true);
        int on_face = 0;
        for (auto it = th.beginBySparsity(face1); it != th.endBySparsity(); ++it) ++on_face;
        EXPECT(on_face == 3 * (3 + 3) + 3);   // P2^3: 3 nodes + 3 edges of the face per component; P1: 3 nodes
    }
    if (fails) { std::printf("test_host_api: %d FAILED\n", fails); return 1; }
    std::printf("test_host_api: all passed\n");
    return 0;
}
