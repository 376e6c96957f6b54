// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/bond_analysis.cpp.
#include "wrap_common.h"
#include "bond_analysis.cpp"
extern "C" {
// bond_analysis.cpp:7 compute_bond
void ref_compute_bond(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
                      const double *dist, const int *nn, int *blen, int *bang, double delta_r, double delta_theta,
                      double rc, int nbins, int num_t)
{
    compute_bond(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), W1I(blen, nbins),
                 W1I(bang, nbins), delta_r, delta_theta, rc, nbins, num_t);
}
// bond_analysis.cpp:120 compute_adf
void ref_compute_adf(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
                     const double *dist, const int *nn, double delta_theta, const double *rc_list, const int *pair_list,
                     int npair, const int *type_list, int nbins, int *bang, int num_t)
{
    compute_adf(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), delta_theta,
                A2D(rc_list, npair, 4), A2I(pair_list, npair, 3), A1I(type_list, N), nbins, W2I(bang, npair, nbins), num_t);
}
}
