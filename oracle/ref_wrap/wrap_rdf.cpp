// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/radial_distribution_function.cpp.
#include "wrap_common.h"
#include "radial_distribution_function.cpp"
extern "C" {
// radial_distribution_function.cpp:22 _rdf
void ref_rdf(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list, double *g,
             int ntype, double rc, int nbin)
{
    _rdf(A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), A1I(type_list, N), W3D(g, ntype, ntype, nbin), rc, nbin);
}
// radial_distribution_function.cpp:56 _rdf_single_species
void ref_rdf_single(const int *verlet, int N, int M, const double *dist, const int *nn, double *g, double rc, int nbin)
{
    _rdf_single_species(A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), W1D(g, nbin), rc, nbin);
}
// radial_distribution_function.cpp:143 _rdf_streaming
void ref_rdf_streaming(const double *x, const double *y, const double *z, int N, const int *type_list, BOXARGS,
                       double *g, int ntype, double rc, int nbin, int num_t)
{
    _rdf_streaming(A1D(x, N), A1D(y, N), A1D(z, N), A1I(type_list, N), BOXPASS, W3D(g, ntype, ntype, nbin), rc, nbin,
                   num_t);
}
}
