// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/cna.cpp.
#include "wrap_common.h"
#include "cna.cpp"
extern "C" {
// cna.cpp:429 FixedCNA
void ref_fcna(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
              const int *nn, int *pattern, double rc, int num_t)
{
    FixedCNA(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A1I(nn, N), W1I(pattern, N), rc, num_t);
}
// cna.cpp:289 AdaptiveCNA
void ref_acna(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
              int *pattern, int num_t)
{
    AdaptiveCNA(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), W1I(pattern, N), num_t);
}
// cna.cpp:163 IdentifyDiamond
void ref_ids(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
             int *new_verlet, int *pattern, int num_t)
{
    IdentifyDiamond(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), W2I(new_verlet, N, 12), W1I(pattern, N), num_t);
}
}
