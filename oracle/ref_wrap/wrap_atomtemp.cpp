// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/atomic_temperature.cpp.
#include "wrap_common.h"
#include "atomic_temperature.cpp"
extern "C" {
// atomic_temperature.cpp:7 compute_temp
void ref_compute_temp(const int *verlet, int N, int M, const double *dist, const double *vx, const double *vy,
                      const double *vz, const double *mass, double *T, double rc, int num_t)
{
    compute_temp(A2I(verlet, N, M), A2D(dist, N, M), A1D(vx, N), A1D(vy, N), A1D(vz, N), A1D(mass, N), W1D(T, N), rc, num_t);
}
}
