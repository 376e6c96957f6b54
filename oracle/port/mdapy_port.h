/* oracle/port/mdapy_port.h -- TEST INFRASTRUCTURE ONLY (see mdapy_port.c). */
#ifndef MDAPY_PORT_H
#define MDAPY_PORT_H
#ifdef __cplusplus
extern "C" {
#endif

void port_build_neighbor(const double *x, const double *y, const double *z, int N, const double *box9,
                         const double *origin3, const int *boundary3, double rc, int *verlet, double *dist, int *nn,
                         int M, int num_t);
void port_sort_verlet_by_distance(int *verlet, double *dist, int N, int M, int k, int num_t);
void port_knn(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, int k, int *indices, double *distances, int num_t);
void port_fcna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, const int *nn, int *pattern, double rc, int num_t);
void port_acna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, int *pattern, int num_t);
void port_ids(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, int *new_verlet, int *pattern, int num_t);
void port_csp(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, int nnei, double *csp, int num_t);
void port_aja(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, const double *dist, int Md, int *aja, int num_t);
void port_get_sq(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
                 const int *boundary3, const int *verlet, int M, const double *dist, const int *nn,
                 const double *weight, const int *llist, int ndeg, int nnn, int lmax, int wl, int wlhat, int average,
                 int use_voronoi, double rc, int use_weight, double *qlm_r, double *qlm_i, double *qnarray, int ncol,
                 int num_t);
void port_solid_liquid(int q6index, const double *Q6, const int *verlet, int N, int M, const double *dist,
                       const int *nn, const double *qlm_r, const double *qlm_i, int ndeg, int nz, double threshold,
                       int n_bond, int *solidliquid, int *nbond, int use_voronoi, int nnn, double rc, int num_t);
void port_rdf(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list, double *g,
              int ntype, double rc, int nbin);
void port_rdf_single(const int *verlet, int N, int M, const double *dist, const int *nn, double *g, double rc,
                     int nbin);
void port_rdf_streaming(const double *x, const double *y, const double *z, int N, const int *type_list,
                        const double *box9, const double *origin3, const int *boundary3, double *g, int ntype,
                        double rc, int nbin, int num_t);
void port_repeat_cell(double *new_pos, const double *old_box, const double *old_pos, int n_old, int nx, int ny,
                      int nz, int num_t);

void port_cnp(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, const double *dist, const int *nn, double *cnp, double rc,
              int num_t);
void port_wcp(const int *verlet, int N, int M, const int *nn, const int *type_list, int T, double *WCP, int num_t);
void port_average_by_neighbor(double rc, const int *verlet, int N, int M, const double *dist, const int *nn,
                              const double *value, double *value_ave, int include_self, int num_t);

int port_get_cluster(const int *verlet, int N, int M, const double *dist, const int *nn, double rc, int *clusters);
void port_filter_by_type(int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                         const int *t1, const int *t2, const double *r, int npair, int num_t);

void port_structure_entropy(double rc, double sigma, int use_local_density, double volume, const double *dist, int N,
                            int M, const int *nn, double *entropy, int num_t);

void port_compute_temp(const int *verlet, int N, int M, const double *dist, const double *vx, const double *vy,
                       const double *vz, const double *mass, double *T, double rc, int num_t);

void port_compute_bond(const double *x, const double *y, const double *z, int N, const double *box9,
                       const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                       const int *nn, int *blen, int *bang, double delta_r, double delta_theta, double rc, int nbins,
                       int num_t);
void port_compute_adf(const double *x, const double *y, const double *z, int N, const double *box9,
                      const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                      const int *nn, double delta_theta, const double *rcs, const int *pairs, int npair,
                      const int *types, int nbins, int *bang, int num_t);

void port_wrap_positions(double *x, double *y, double *z, int N, const double *box9, const double *origin3,
                         const int *boundary3, int num_t);

#ifdef __cplusplus
}
#endif
#endif
