// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/centro_symmetry_parameter.cpp.
#include "wrap_common.h"
#include "centro_symmetry_parameter.cpp"
extern "C" {
// centro_symmetry_parameter.cpp:12 get_csp
void ref_csp(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M, int nnei,
             double *csp, int num_t)
{
    get_csp(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), nnei, W1D(csp, N), num_t);
}
}
