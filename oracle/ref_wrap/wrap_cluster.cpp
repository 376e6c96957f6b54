// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/cluster.cpp.
#include "wrap_common.h"
#include "cluster.cpp"
extern "C" {
// cluster.cpp:9 get_cluster (particleClusters pre-filled with -1 by the caller)
int ref_get_cluster(const int *verlet, int N, int M, const double *dist, const int *nn, double rc, int *clusters)
{
    return get_cluster(A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), rc, W1I(clusters, N));
}
// cluster.cpp:62 get_cluster_by_bond
int ref_get_cluster_by_bond(const int *verlet, int N, int M, const int *nn, int *clusters)
{
    return get_cluster_by_bond(A2I(verlet, N, M), A1I(nn, N), W1I(clusters, N));
}
// cluster.cpp:114 filter_by_type
void ref_filter_by_type(int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                        const int *t1, const int *t2, const double *r, int npair, int num_t)
{
    filter_by_type(W2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), A1I(type_list, N), A1I(t1, npair), A1I(t2, npair),
                   A1D(r, npair), num_t);
}
}
