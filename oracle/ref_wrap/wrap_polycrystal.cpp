// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/polycrystal.cpp.
#include "wrap_common.h"
#include "polycrystal.cpp"
#include <cstring>
extern "C" {
// polycrystal.cpp:21 transform_and_filter -> number of atoms kept; out3 must hold 3 N doubles
int ref_transform_and_filter(const double *x, const double *y, const double *z, int N, const double *R9,
                             const double *center3, const double *target3, const double *coeffs, int nfaces,
                             double *out3, int num_t)
{
    auto r = transform_and_filter(A1D(x, N), A1D(y, N), A1D(z, N), A2D(R9, 3, 3), A1D(center3, 3), A1D(target3, 3),
                                  A2D(coeffs, nfaces, 4), num_t);
    const size_t n = r.shape(0);
    std::memcpy(out3, r.data(), sizeof(double) * 3 * n);
    delete[] r.data();
    return (int)n;
}
}
