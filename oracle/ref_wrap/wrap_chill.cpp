// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/chill_plus.cpp.
#include "wrap_common.h"
#include "chill_plus.cpp"
extern "C" {
// chill_plus.cpp:76 compute_chill_plus
void ref_compute_chill_plus(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
                            const double *dist, const int *nn, double rc, int *pattern, int num_t)
{
    compute_chill_plus(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), rc,
                       W1I(pattern, N), num_t);
}
}
