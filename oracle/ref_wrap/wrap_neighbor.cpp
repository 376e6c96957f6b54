// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/neighbor.cpp.
#include "wrap_common.h"
#include "neighbor.cpp" // resolved through -I/root/reference/src
#include <cstdlib>
#include <cstring>
extern "C" {
// neighbor.cpp:351 build_neighbor
void ref_build_neighbor(const double *x, const double *y, const double *z, int N, BOXARGS, double rc,
                        int *verlet, double *dist, int *nn, int M, int num_t)
{
    build_neighbor(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, rc, W2I(verlet, N, M), W2D(dist, N, M), W1I(nn, N), num_t);
}
// neighbor.cpp:189 build_neighbor_without_max_neigh; caller frees with ref_free
int ref_build_neighbor_auto(const double *x, const double *y, const double *z, int N, BOXARGS, double rc,
                            int **verlet, double **dist, int **nn, int num_t)
{
    auto t = build_neighbor_without_max_neigh(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, rc, num_t);
    *verlet = std::get<0>(t).data();
    *dist = std::get<1>(t).data();
    *nn = std::get<2>(t).data();
    return (int)std::get<0>(t).shape(1);
}
void ref_free_int(int *p) { delete[] p; }
void ref_free_double(double *p) { delete[] p; }
// neighbor.cpp:745 sort_verlet_by_distance
void ref_sort_verlet_by_distance(int *verlet, double *dist, int N, int M, int k, int num_t)
{
    sort_verlet_by_distance(W2I(verlet, N, M), W2D(dist, N, M), k, num_t);
}
// neighbor.cpp:675 wrap_positions
void ref_wrap_positions(double *x, double *y, double *z, int N, BOXARGS, int num_t)
{
    wrap_positions(W1D(x, N), W1D(y, N), W1D(z, N), BOXPASS, num_t);
}
// neighbor.cpp:704 average_by_neighbor
void ref_average_by_neighbor(double rc, const int *verlet, int N, int M, const double *dist, const int *nn,
                             const double *value, double *value_ave, int include_self, int num_t)
{
    average_by_neighbor(rc, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), A1D(value, N), W1D(value_ave, N),
                        include_self != 0, num_t);
}
// neighbor.cpp:390 filter_overlap_atom -> keep flags (1 byte per atom)
void ref_filter_overlap_atom(const double *x, const double *y, const double *z, int N, BOXARGS, double rc,
                             unsigned char *keep, int num_t)
{
    auto f = filter_overlap_atom(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, rc, num_t);
    for (int i = 0; i < N; ++i) keep[i] = f.data()[i] ? 1 : 0;
    delete[] f.data();
}
}
