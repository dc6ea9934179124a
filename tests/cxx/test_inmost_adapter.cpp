// anifem_b200/inmost_adapter.hpp against the mock INMOST of the oracle (oracle/mock_inmost/inmost.h: the bounded INMOST surface
// the reference's inmost_interface uses).  CPU only: the mesh arrays handed to Assembler::SetMesh and the CSR <-> Sparse::Matrix
// conversions are checked, no context is created.  TEST INFRASTRUCTURE use of oracle/: the mock stands in for INMOST.
#include <cstdio>
#include <algorithm>

#include "inmost.h"

#include "anifem_b200/inmost_adapter.hpp"

using namespace Ani;
static int fails = 0;
#define EXPECT(c)                                                                  \
    do {                                                                           \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } \
    } while (0)

int main() {
    // two tetrahedra sharing a face, one of them negatively oriented on purpose
    const double xyz[5 * 3] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1};
    const long tets[2 * 4] = {0, 1, 2, 3, 1, 3, 2, 4};
    INMOST::Mesh m;
    m.BuildTets(5, xyz, 2, tets);
    b200::InmostMeshArrays a = b200::mesh_to_arrays(&m);
    EXPECT(a.x.size() == 5 && a.v[0].size() == 2 && a.node.size() == 5 && a.cell.size() == 2);
    for (int n = 0; n < 5; ++n) EXPECT(a.x[n] == xyz[3 * n] && a.y[n] == xyz[3 * n + 1] && a.z[n] == xyz[3 * n + 2] && a.node_gid[n] == n);
    for (int e = 0; e < 2; ++e) {   // the same four nodes as the cell (any order: the library orients the tets itself)
        long got[4] = {a.v[0][e], a.v[1][e], a.v[2][e], a.v[3][e]}, want[4] = {tets[4 * e], tets[4 * e + 1], tets[4 * e + 2], tets[4 * e + 3]};
        std::sort(got, got + 4); std::sort(want, want + 4);
        EXPECT(std::equal(got, got + 4, want));
    }
    bool thrown = false;
    try { b200::mesh_to_arrays(nullptr); } catch (std::runtime_error&) { thrown = true; }
    EXPECT(thrown);

    // CSR -> INMOST::Sparse::Matrix -> CSR round trip, add semantics, rhs
    CsrMatrix A;
    A.row_begin = 10; A.row_end = 13;
    A.rowptr = {0, 2, 3, 6};
    A.colind = {10, 12, 11, 3, 10, 12};
    A.val = {4.0, -1.0, 2.5, 7.0, -1.0, 3.0};
    INMOST::Sparse::Matrix M("A");
    b200::csr_to_inmost(A, M);
    EXPECT(M.GetFirstIndex() == 10 && M.GetLastIndex() == 13 && M[12].Size() == 3 && M[12].get_safe(3) == 7.0 && M[10].get_safe(12) == -1.0);
    CsrMatrix B = b200::inmost_to_csr(M);
    EXPECT(B.row_begin == A.row_begin && B.row_end == A.row_end && B.rowptr == A.rowptr && B.colind == A.colind && B.val == A.val);
    b200::csr_to_inmost(A, M, /*add*/ true);
    CsrMatrix C = b200::inmost_to_csr(M);
    EXPECT(C.colind == A.colind);
    for (std::size_t k = 0; k < A.val.size(); ++k) EXPECT(C.val[k] == 2 * A.val[k]);
    M[11][12] = 9.0;   // an entry appended out of order by INMOST's find-or-insert is sorted on the way back
    CsrMatrix D = b200::inmost_to_csr(M);
    EXPECT(D.rowptr[2] - D.rowptr[1] == 2 && D.colind[static_cast<std::size_t>(D.rowptr[1])] == 11 && D.colind[static_cast<std::size_t>(D.rowptr[1]) + 1] == 12);
    INMOST::Sparse::Vector v("b");
    b200::rhs_to_inmost({1.0, 2.0, 3.0}, 10, v);
    b200::rhs_to_inmost({1.0, 2.0, 3.0}, 10, v, true);
    EXPECT(v.GetFirstIndex() == 10 && v.GetLastIndex() == 13 && v[11] == 4.0);
    if (fails) { std::printf("test_inmost_adapter: %d FAILED\n", fails); return 1; }
    std::printf("test_inmost_adapter: all passed\n");
    return 0;
}
