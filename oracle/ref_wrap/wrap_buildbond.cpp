// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/build_bond.cpp.
#include "wrap_common.h"
#include "build_bond.cpp"
#include <cstring>
extern "C" {
// build_bond.cpp:9 build_bond -> number of bonds; out must hold 2 * N * M ints
int ref_build_bond(const int *verlet, int N, int M, const double *dist, const int *nn, const int *types,
                   const double *cutoff, int ntype, int *out, int num_t)
{
    auto r = build_bond(A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), A1I(types, N), A2D(cutoff, ntype, ntype), num_t);
    const size_t n = r.shape(0);
    std::memcpy(out, r.data(), sizeof(int) * 2 * n);
    delete[] r.data();
    return (int)n;
}
}
