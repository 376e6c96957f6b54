// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/ackland_jones_analysis.cpp.
#include "wrap_common.h"
#include "ackland_jones_analysis.cpp"
extern "C" {
// ackland_jones_analysis.cpp:9 compute_aja
void ref_aja(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
             const double *dist, int Md, int *aja, int num_t)
{
    compute_aja(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, Md), W1I(aja, N), num_t);
}
}
