// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/structure_entropy.cpp.
#include "wrap_common.h"
#include "structure_entropy.cpp"
extern "C" {
// structure_entropy.cpp:11 calculate_structure_entropy
void ref_structure_entropy(double rc, double sigma, int use_local_density, double volume, const double *dist, int N,
                           int M, const int *nn, double *entropy, int num_t)
{
    calculate_structure_entropy(rc, sigma, use_local_density != 0, volume, A2D(dist, N, M), A1I(nn, N), W1D(entropy, N),
                                num_t);
}
}
