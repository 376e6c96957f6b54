// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/steinhardt_bond_orientation.cpp.
#include "wrap_common.h"
#include "steinhardt_bond_orientation.cpp"
extern "C" {
// steinhardt_bond_orientation.cpp:677 get_sq
void ref_get_sq(const double *x, const double *y, const double *z, int N, BOXARGS, const int *verlet, int M,
                const double *dist, const int *nn, const double *weight, int wrows, int wcols, const int *llist,
                int ndeg, int nnn, int lmax, int wl, int wlhat, int average, int use_voronoi, double rc,
                int use_weight, double *qlm_r, double *qlm_i, double *qnarray, int ncol, int num_t)
{
    get_sq(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N),
           A2D(weight, wrows, wcols), A1I(llist, ndeg), nnn, lmax, wl != 0, wlhat != 0, average != 0,
           use_voronoi != 0, rc, use_weight != 0, W3D(qlm_r, N, ndeg, 2 * lmax + 1),
           W3D(qlm_i, N, ndeg, 2 * lmax + 1), W2D(qnarray, N, ncol), num_t);
}
// steinhardt_bond_orientation.cpp:578 identifySolidLiquid
void ref_solid_liquid(int Q6index, const double *Q6, const int *verlet, int N, int M, const double *dist,
                      const int *nn, const double *qlm_r, const double *qlm_i, int ndeg, int nz, double threshold,
                      int n_bond, int *solidliquid, int *nbond, int use_voronoi, int nnn, double rc, int num_t)
{
    identifySolidLiquid(Q6index, A1D(Q6, N), A2I(verlet, N, M), A2D(dist, N, M), A1I(nn, N), A3D(qlm_r, N, ndeg, nz),
                        A3D(qlm_i, N, ndeg, nz), threshold, n_bond, W1I(solidliquid, N), W1I(nbond, N),
                        use_voronoi != 0, nnn, rc, num_t);
}
}
