// CPU test helper for anifem_b200/enumerator.hpp (no GPU, no library): reads a connectivity file, writes the elem -> dof table.
//   test_enum <type 0..5> <nnode> <ntet> <tets.i32: 4 arrays of ntet> <out.i64> fem0 vec0 [fem1 vec1 ...]
#include <cstdio>
#include <cstdlib>

#include "anifem_b200/enumerator.hpp"

int main(int argc, char** argv) {
    if (argc < 8 || (argc - 6) % 2) { std::fprintf(stderr, "usage: test_enum type nnode ntet tets.i32 out.i64 fem vec [fem vec ...]\n"); return 2; }
    const int type = std::atoi(argv[1]);
    const long long nnode = std::atoll(argv[2]), ntet = std::atoll(argv[3]);
    std::vector<int32_t> v((size_t)4 * ntet);
    FILE* f = std::fopen(argv[4], "rb");
    if (!f || std::fread(v.data(), 4, v.size(), f) != v.size()) { std::fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    std::fclose(f);
    std::vector<Ani::EnumVar> vars;
    for (int k = 6; k + 1 < argc; k += 2) vars.push_back({std::atoi(argv[k]), std::atoi(argv[k + 1])});
    try {
        const Ani::DofEnumeration en = Ani::enumerate_dofs((Ani::ASSEMBLING_TYPE)type, nnode, ntet, v.data(), v.data() + ntet, v.data() + 2 * ntet,
                                                         v.data() + 3 * ntet, vars);
        f = std::fopen(argv[5], "wb");
        const int64_t head[2] = {en.nrows, en.nloc};
        std::fwrite(head, 8, 2, f);
        std::fwrite(en.elem2dof.data(), 8, en.elem2dof.size(), f);
        std::fclose(f);
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
