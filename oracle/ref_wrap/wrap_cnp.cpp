// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/common_neighbor_parameter.cpp.
#include "wrap_common.h"
#include "common_neighbor_parameter.cpp"
extern "C" {
// common_neighbor_parameter.cpp:10 compute_cnp
void ref_cnp(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
             const double *dist, const int *nn, double *cnp, double rc, int num_t)
{
    compute_cnp(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), W1D(cnp, N), rc,
                num_t);
}
}
