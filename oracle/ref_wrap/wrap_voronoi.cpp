// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/voronoi.cpp (voro++ vendored under extern/voro++).
#include "wrap_common.h"
#include "voronoi.cpp"
extern "C" {
// voronoi.cpp:16 get_voronoi_volume_number_radius
void ref_voronoi_volume_number_radius(const double *x, const double *y, const double *z, int N, BOXARGS, double *volume,
                                      int *neighbor_number, double *cavity_radius, int num_t)
{
    get_voronoi_volume_number_radius(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, W1D(volume, N), W1I(neighbor_number, N),
                                     W1D(cavity_radius, N), num_t);
}
// voronoi.cpp:307 get_voronoi_neighbor; caller frees with ref_voro_free_*; returns the row width
int ref_voronoi_neighbor(const double *x, const double *y, const double *z, int N, BOXARGS, double a_thr, double r_thr,
                         int **verlet, double **dist, double **area, int **nn, int num_t)
{
    auto t = get_voronoi_neighbor(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, a_thr, r_thr, num_t);
    *verlet = std::get<0>(t).data();
    *dist = std::get<1>(t).data();
    *area = std::get<2>(t).data();
    *nn = std::get<3>(t).data();
    return (int)std::get<0>(t).shape(1);
}
// voronoi.cpp:73 get_voronoi_volume_number_radius_tri (box9 = the LAMMPS-aligned cell, rotation9 row-major)
void ref_voronoi_volume_number_radius_tri(const double *x, const double *y, const double *z, int N, BOXARGS,
                                          const double *rotation9, double *volume, int *neighbor_number,
                                          double *cavity_radius, int need_rotation, int num_t)
{
    get_voronoi_volume_number_radius_tri(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2D(rotation9, 3, 3), W1D(volume, N),
                                         W1I(neighbor_number, N), W1D(cavity_radius, N), need_rotation != 0, num_t);
}
// voronoi.cpp:149 get_voronoi_neighbor_tri
int ref_voronoi_neighbor_tri(const double *x, const double *y, const double *z, int N, BOXARGS, const double *rotation9,
                             int need_rotation, double a_thr, double r_thr, int **verlet, double **dist, double **area,
                             int **nn, int num_t)
{
    auto t = get_voronoi_neighbor_tri(A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2D(rotation9, 3, 3), need_rotation != 0,
                                      a_thr, r_thr, num_t);
    *verlet = std::get<0>(t).data();
    *dist = std::get<1>(t).data();
    *area = std::get<2>(t).data();
    *nn = std::get<3>(t).data();
    return (int)std::get<0>(t).shape(1);
}
void ref_voro_free_int(int *p) { delete[] p; }
void ref_voro_free_double(double *p) { delete[] p; }
}
