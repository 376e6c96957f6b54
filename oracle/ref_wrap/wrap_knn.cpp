// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/fast_knn.cpp.
#include "wrap_common.h"
#include "fast_knn.cpp"
extern "C" {
// fast_knn.cpp:846 fast_knn::knn
void ref_knn(const double *x, const double *y, const double *z, int N, BOXARGS, int k, int *indices,
             double *distances, int num_t)
{
    fast_knn::knn(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, k, W2I(indices, N, k), W2D(distances, N, k), num_t);
}
}
