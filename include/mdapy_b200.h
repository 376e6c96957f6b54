/* include/mdapy_b200.h -- C ABI of libmdapy_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for mdapy's neighbour + structural-descriptor
 * hot path.  Every entry point in section A replaces one nanobind function of
 * the reference (file:line given; signatures per SURVEY.md section 8b): same
 * argument order and meaning, NumPy buffers become plain pointers + extents,
 * `num_t` (the OpenMP team size) is accepted and ignored.  Arrays are
 * C-contiguous, double = f64, int = int32, positions are three separate
 * vectors, box rows are lattice vectors, boundary[d] = 1 means periodic.
 * Section A takes HOST pointers and round-trips over PCIe per call; section B
 * is the device-resident handle the Python `System` uses so that chained
 * cal_* calls keep the lists in HBM.
 *
 * Every function returns 0 on success or an MDB_ERR_* code; the message is
 * available from mdb_last_error().  MDB_ERR_VALUE maps to Python ValueError,
 * everything else to RuntimeError (the reference surfaces std::runtime_error
 * from src/box.h:85,186 the same way).
 */
#ifndef MDAPY_B200_H
#define MDAPY_B200_H

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#include <stddef.h>
#pragma GCC visibility push(default)
#endif

#define MDB_OK 0
#define MDB_ERR_CUDA 1
#define MDB_ERR_VALUE 2
#define MDB_ERR_BOX 3
#define MDB_ERR_STATE 4

const char *mdb_last_error(void);
const char *mdb_version(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
long long mdb_launch_count(void);
int mdb_device_count(int *count);
/* Released device blocks and pinned host blocks are cached for the next frame; this returns them to the
 * driver.  mdb_host_alloc/mdb_host_free hand out page-locked host memory for result columns (device ->
 * host copies into it run at full PCIe rate and need no staging). */
int mdb_trim_cache(void);
int mdb_host_alloc(size_t bytes, void **ptr);
void mdb_host_free(void *ptr);

/* ------------------------------------------------------------------------
 * Section A: host-pointer drop-ins, one per reference nanobind function
 * ---------------------------------------------------------------------- */

/* _neighbor.build_neighbor, src/neighbor.cpp:351.  verlet/dist/nn are caller
 * allocated; rows are fully (re)written: -1 / rc+1.0 padding as
 * src/mdapy/neighbor.py:125-129 prefills them.  nn always holds the true
 * count, even beyond M (caller raises ValueError, neighbor.py:135-142). */
int mdb_build_neighbor(const double *x, const double *y, const double *z, int N, const double *box9,
                       const double *origin3, const int *boundary3, double rc, int *verlet, double *dist,
                       int *nn, int M, int num_t);

/* _neighbor.build_neighbor_without_max_neigh, src/neighbor.cpp:189.  Two-step
 * form of the capsule-owning return: the first call computes the lists on the
 * device and reports M = max(count, 1); the caller allocates (N, M) arrays and
 * fetches them with mdb_neighbor_auto_fetch, which also releases the handle. */
int mdb_build_neighbor_without_max_neigh(const double *x, const double *y, const double *z, int N,
                                         const double *box9, const double *origin3, const int *boundary3,
                                         double rc, int num_t, void **handle, int *M);
int mdb_neighbor_auto_fetch(void *handle, int *verlet, double *dist, int *nn);

/* _neighbor.sort_verlet_by_distance, src/neighbor.cpp:745 (in place). */
int mdb_sort_verlet_by_distance(int *verlet, double *dist, int N, int M, int k, int num_t);

/* _fast_knn.knn, src/fast_knn.cpp:846.  k <= 24; periodic images are distinct
 * neighbours; rows ascending in distance, short rows padded -1 / -1.0. */
int mdb_knn(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
            const int *boundary3, int k, int *indices, double *distances, int num_t);

/* _cna.fcna, src/cna.cpp:429.  pattern is fully written (0 = other). */
int mdb_fcna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
             const int *boundary3, const int *verlet, int M, const int *nn, int *pattern, double rc, int num_t);
/* _cna.acna, src/cna.cpp:289.  verlet rows: >= 14 neighbours, ascending distance. */
int mdb_acna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
             const int *boundary3, const int *verlet, int M, int *pattern, int num_t);
/* _cna.ids, src/cna.cpp:163 (IdentifyDiamond).  verlet rows: >= 4 neighbours, ascending distance.
 * new_verlet (N x 12, may be NULL) receives the second-shell list; pattern is fully written (0-6). */
int mdb_ids(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
            const int *boundary3, const int *verlet, int M, int *new_verlet, int *pattern, int num_t);

/* _csp.get_csp, src/centro_symmetry_parameter.cpp:12. */
int mdb_get_csp(const double *x, const double *y, const double *z, int N, const double *box9,
                const double *origin3, const int *boundary3, const int *verlet, int M, int nnei, double *csp,
                int num_t);

/* _aja.compute_aja, src/ackland_jones_analysis.cpp:9. */
int mdb_compute_aja(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    int Md, int *aja, int num_t);

/* _ptm.get_ptm, src/polyhedral_template_matching.cpp:135.  verlet rows: nearest neighbours in ascending
 * distance (18 wanted).  output (N, ocols): type, ordering, rmsd, interatomic distance, qw, qx, qy, qz;
 * ptm_indices (N, icols): the atom, then its matched neighbours, -1 padded.  All eight structures of the
 * reference: sc / fcc / hcp / ico / bcc and the two-shell dcub / dhex / graphene (neighbours of neighbours).
 * ptm_indices follow THIS library's template point order (mdb_identify_sftb_fcc
 * carries the matching layer table). */
int mdb_get_ptm(const char *structure, const double *x, const double *y, const double *z, int N, const double *box9,
                const double *origin3, const int *boundary3, const int *verlet, int M, const int *atom_types,
                int ntypes, double rmsd_threshold, double *output, int ocols, int *ptm_indices, int icols,
                int num_t);

/* _sbo.get_sq, src/steinhardt_bond_orientation.cpp:677.  qlm_r/qlm_i (N, ndeg, 2*lmax+1) are inout
 * (zeroed by the caller, steinhardt_bond_orientation.py:228-229), qnarray (N, ncol) is out; rc is the
 * value the Python wrapper passes (1e9 / 1e10 for the nnn / voronoi neighbour sources).  llist is
 * int32 (the reference wrapper hands int64 and relies on nanobind's cast: convert before calling). */
int mdb_get_sq(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, const double *dist, const int *nn,
               const double *weight, const int *llist, int ndeg, int nnn, int lmax, int wl, int wlhat, int average,
               int use_voronoi, double rc, int use_weight, double *qlm_r, double *qlm_i, double *qnarray, int ncol,
               int num_t);
/* _sbo.identifySolidLiquid, src/steinhardt_bond_orientation.cpp:578.  solidliquid must be pre-zeroed. */
int mdb_identify_solid_liquid(int Q6index, const double *Q6, const int *verlet, int N, int M, const double *dist,
                              const int *nn, const double *qlm_r, const double *qlm_i, int ndeg, int nz,
                              double threshold, int n_bond, int *solidliquid, int *nbond, int use_voronoi, int nnn,
                              double rc, int num_t);

/* _rdf._rdf / _rdf._rdf_single_species / _rdf._rdf_streaming,
 * src/radial_distribution_function.cpp:22, 56, 143.  g is ACCUMULATED into (caller zeroes it). */
int mdb_rdf(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list, double *g,
            int ntype, double rc, int nbin);
int mdb_rdf_single_species(const int *verlet, int N, int M, const double *dist, const int *nn, double *g, double rc,
                           int nbin);
int mdb_rdf_streaming(const double *x, const double *y, const double *z, int N, const int *type_list,
                      const double *box9, const double *origin3, const int *boundary3, double *g, int ntype,
                      double rc, int nbin, int num_t);

/* _neighbor.wrap_positions(x, y, z (in place), box, origin, boundary, num_t) -- src/neighbor.cpp:675 */
int mdb_wrap_positions(double *x, double *y, double *z, int N, const double *box9, const double *origin3,
                       const int *boundary3, int num_t);

/* ---- further neighbour-list consumers ---- */
/* _cnp.compute_cnp(x,y,z,box,origin,boundary,verlet_list,distance_list,neighbor_number,cnp,rc,num_t)
 * -- src/common_neighbor_parameter.cpp:10 */
int mdb_compute_cnp(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    const int *nn, double *cnp, double rc, int num_t);
/* _wcp.get_wcp(verlet_list, neighbor_number, type_list, Ntype, WCP[Ntype,Ntype], num_t)
 * -- src/warren_cowley_parameter.cpp:9 */
int mdb_get_wcp(const int *verlet, int N, int M, const int *nn, const int *type_list, int Ntype, double *WCP,
                int num_t);
/* _neighbor.average_by_neighbor(rc, verlet_list, distance_list, neighbor_number, value, value_ave,
 * include_self, num_t) -- src/neighbor.cpp:704 */
int mdb_average_by_neighbor(double rc, const int *verlet, int N, int M, const double *dist, const int *nn,
                            const double *value, double *value_ave, int include_self, int num_t);

/* _cluster.get_cluster(verlet_list, distance_list, neighbor_number, rc, particleClusters) -> cluster count
 * -- src/cluster.cpp:9;  _cluster.get_cluster_by_bond(verlet_list, neighbor_number, particleClusters) -- :62;
 * _cluster.filter_by_type(verlet_list (inout), distance_list, neighbor_number, type_list, type1, type2, r, num_t)
 * -- :114 */
int mdb_get_cluster(const int *verlet, int N, int M, const double *dist, const int *nn, double rc,
                    int *particle_clusters, int *cluster_number);
int mdb_get_cluster_by_bond(const int *verlet, int N, int M, const int *nn, int *particle_clusters,
                            int *cluster_number);
int mdb_filter_by_type(int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                       const int *type1, const int *type2, const double *r, int npair, int num_t);

/* _structure_entropy.calculate_structure_entropy(rc, sigma, use_local_density, volume, distance_list,
 * neighbor_number, entropy, num_t) -- src/structure_entropy.cpp:11 */
int mdb_calculate_structure_entropy(double rc, double sigma, int use_local_density, double volume, const double *dist,
                                    int N, int M, const int *nn, double *entropy, int num_t);

/* _atomtemp.compute_temp(verlet_list, distance_list, vx, vy, vz, mass_list, T, rc, num_t)
 * -- src/atomic_temperature.cpp:7 (velocities in A/ps, masses in g/mol, T in K) */
int mdb_compute_temp(const int *verlet, int N, int M, const double *dist, const double *vx, const double *vy,
                     const double *vz, const double *mass, double *T, double rc, int num_t);

/* _bond_analysis.compute_bond(x,y,z,box,origin,boundary,verlet,dist,nn,bond_length_distribution I[nbins] (+=),
 * bond_angle_distribution I[nbins] (+=), delta_r, delta_theta, rc, nbins, num_t) -- src/bond_analysis.cpp:7 */
int mdb_compute_bond(const double *x, const double *y, const double *z, int N, const double *box9,
                     const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                     const int *nn, int *bond_length_distribution, int *bond_angle_distribution, double delta_r,
                     double delta_theta, double rc, int nbins, int num_t);
/* _bond_analysis.compute_adf(x,y,z,box,origin,boundary,verlet,dist,nn,delta_theta,rc_list D[Npair,4],
 * pair_list I[Npair,3], type_list, nbins, bond_angle_distribution I[Npair,nbins] (+=), num_t) -- :120 */
int mdb_compute_adf(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    const int *nn, double delta_theta, const double *rc_list, const int *pair_list, int npair,
                    const int *type_list, int nbins, int *bond_angle_distribution, int num_t);

/* ------------------------------------------------------------------------
 * Section B: device-resident system handle
 * ---------------------------------------------------------------------- */
typedef struct MdbSystem mdb_system;

int mdb_system_create(int device, mdb_system **out);
void mdb_system_destroy(mdb_system *s);
/* run on a caller-owned CUDA stream (cudaStream_t as void*, e.g. torch's current stream) */
int mdb_system_set_stream(mdb_system *s, void *cuda_stream);
int mdb_system_synchronize(mdb_system *s);

/* upload host coordinates (H2D) / borrow device coordinates (no copy; must outlive use) */
int mdb_system_set_atoms(mdb_system *s, const double *x, const double *y, const double *z, int N,
                         const double *box9, const double *origin3, const int *boundary3);
int mdb_system_set_atoms_device(mdb_system *s, const double *dx, const double *dy, const double *dz, int N,
                                const double *box9, const double *origin3, const int *boundary3);

/* Decomposed frame (one rank per GPU, mdapy_b200/distributed.py): the first n_owned atoms are
 * owned, the rest are ghosts; dgid holds GLOBAL atom ids (device int32).  Only the window of
 * nplanes global x cell planes starting at plane0 (ghost, owned..., ghost; periodic wrap) is
 * stored.  Lists and per-atom outputs then have n_owned rows; rows are ordered exactly like the
 * single-GPU build (cells are the GLOBAL grid, ties inside a cell by global id) and
 * mdb_system_fetch_neighbor exports global ids. */
int mdb_system_set_slab_device(mdb_system *s, const double *dx, const double *dy, const double *dz,
                               const int *dgid, int n_local, int n_owned, int plane0, int nplanes,
                               const double *box9, const double *origin3, const int *boundary3);
/* Decomposed frame: fraction (0, 1] of the box volume the local atoms occupy.  A density hint only
 * (cell size of the k-nearest search grid); results do not depend on it. */
int mdb_system_set_local_fraction(mdb_system *s, double fraction);
/* cell grid of the cut-off search for (box, rc): src/neighbor.cpp:367-370 */
int mdb_cell_grid(const double *box9, const double *origin3, const int *boundary3, double rc, int *n3);
/* global x cell plane of each atom (device arrays), same arithmetic as src/neighbor.cpp:30-62 */
int mdb_cell_planes_device(const double *dx, const double *dy, const double *dz, int N, const double *box9,
                           const double *origin3, const int *boundary3, double rc, int *dplane, void *cuda_stream);
/* Device-side halo exchange of a decomposed frame (mdapy_b200/distributed.py; no host round trip per frame):
 * pack appends the atoms of the first / last `halo` owned x cell planes [lo, hi) to two send buffers of `cap` rows
 * of 4 doubles (x, y, z raw, global id); row 0 is a header holding the row count.  unpack appends the rows of one
 * or two received buffers behind the n_owned owned atoms of dx/dy/dz/dgid (room = free rows there) and leaves the
 * new atom count in *dtotal (-1: a buffer overflowed).  All pointers are device pointers. */
int mdb_slab_pack_device(const double *dx, const double *dy, const double *dz, const int *dgid, int N,
                         const double *box9, const double *origin3, const int *boundary3, double rc, int lo, int hi,
                         int halo, double *send_left, double *send_right, int cap, int *dcounts, void *cuda_stream);
int mdb_slab_unpack_device(const double *recv_a, const double *recv_b, int cap, double *dx, double *dy, double *dz,
                           int *dgid, int n_owned, int room, int *dtotal, void *cuda_stream);

/* cut-off list kept on the device.  max_neigh <= 0: size automatically
 * (neighbor.cpp:189 semantics, M = max(count,1)).  Returns row width and the
 * largest count; with max_neigh > 0 and max_count > max_neigh the list is
 * truncated exactly like the reference and the caller should raise. */
int mdb_system_build_neighbor(mdb_system *s, double rc, int max_neigh, int *M, int *max_count);
/* k-nearest list (sorted, width k) kept on the device; replaces the cached list like
 * System.build_nearest_neighbor does (src/mdapy/system.py:1226-1263) */
int mdb_system_build_knn(mdb_system *s, int k);
int mdb_system_sort_neighbor(mdb_system *s, int k);
int mdb_system_neighbor_min_count(mdb_system *s, int *min_count);
/* D2H of the cached list; any pointer may be NULL */
int mdb_system_fetch_neighbor(mdb_system *s, int *verlet, double *dist, int *nn);
/* replace the cached list by host arrays (H2D), e.g. a user-provided list */
int mdb_system_put_neighbor(mdb_system *s, const int *verlet, const double *dist, const int *nn, int M, double rc,
                            int kind);
/* raw device pointers of the cached list (for zero-copy consumers); M via mdb_system_build_neighbor */
int mdb_system_neighbor_device(mdb_system *s, int **verlet, double **dist, int **nn, int *M);

/* descriptors on the cached list; results stay on the device when out == NULL */
int mdb_system_fcna(mdb_system *s, double rc, int *pattern_host);
/* Neighbour search (neighbor.cpp:130-186) and FixedCNA (cna.cpp:429-506) in ONE pass that never writes the
 * list: what System.cal_common_neighbor_analysis(rc) needs when nothing else reads the list (it is built
 * lazily on first access).  *used = 0: frame not eligible (triclinic / tiny box), nothing was computed. */
int mdb_system_fused_cna(mdb_system *s, double rc, int *pattern_host, int *used);

/* ---- CHILL+ and build_bond ---------------------------------------------------------------
 * mdb_compute_chill_plus <- _chill_plus.compute_chill_plus(x, y, z, box, origin, boundary, verlet_list,
 *                           distance_list, neighbor_number, rc, pattern, num_t)      (src/chill_plus.cpp:76)
 * mdb_build_bond         <- _build_bond.build_bond(verlet_list, distance_list, neighbor_number, type_list,
 *                           cutoff_matrix, num_t) -> (Nbond, 2)                       (src/build_bond.cpp:9)
 *                           two calls: bonds == NULL returns *nbond, then a buffer of 2 * nbond ints is filled;
 *                           rows come in (i, list slot) order (the reference's order depends on the OpenMP
 *                           schedule; its caller sorts and de-duplicates, system.py:1405-1410). */
int mdb_compute_chill_plus(const double *x, const double *y, const double *z, int N, const double *box9,
                           const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                           const int *nn, double rc, int *pattern, int num_t);
int mdb_build_bond(const int *verlet, int N, int M, const double *dist, const int *nn, const int *types,
                   const double *cutoff_matrix, int ntype, int *bonds, int *nbond, int num_t);
int mdb_system_chill_plus(mdb_system *s, double rc, int *pattern_host);
int mdb_system_build_bond(mdb_system *s, const int *types, const double *cutoff_matrix, int ntype, int *bonds_host,
                          int *nbond);

/* ---- FCC planar faults on the PTM result -------------------------------------------------
 * mdb_identify_sftb_fcc <- _fccpft.identify_sftb_fcc(hcp_indices, hcp_neighbors, ptm_indices[N,12], structure_types,
 *                          fault_types, identify_esf, num_t)          (src/identify_fcc_planar_faults.cpp:43)
 *   fault_types: 0 non-HCP, 1 other, 2 intrinsic SF, 3 coherent twin boundary, 4 multi-layer SF, 5 extrinsic SF.
 *   hcp_indices / hcp_neighbors (the reference's scratch) are accepted and ignored.  index_order names the HCP
 *   template point order of ptm_indices: 0 = this library's PTM, 1 = the reference's (extern/ptm).
 * mdb_system_planar_faults: the same on the handle's last PTM result (device resident). */
int mdb_identify_sftb_fcc(const int *hcp_indices, int n_hcp, int *hcp_neighbors, const int *ptm_indices,
                          const int *structure_types, int N, int *fault_types, int identify_esf, int index_order,
                          int num_t);
int mdb_system_planar_faults(mdb_system *s, int identify_esf, int *fault_host);

/* ---- builders: benchmark-size inputs are generated in HBM ---------------------------------
 * mdb_repeat_cell            <- _repeat_cell.repeat_cell(new_pos, old_box, old_pos, nx, ny, nz, num_t)
 *                               (src/repeat_cell.cpp:19; new_pos: 3 * n_old * nx*ny*nz doubles, cell-major, iz fastest)
 * mdb_transform_and_filter   <- _polycrystal.transform_and_filter(x, y, z, rotation, center, target, coeffs, num_t)
 *                               (src/polycrystal.cpp:21; out_pos must hold 3 N doubles, *count rows are returned)
 * mdb_filter_overlap_atom    <- _neighbor.filter_overlap_atom(x, y, z, box, origin, boundary, rc, num_t)
 *                               (src/neighbor.cpp:390; keep: one byte per atom, 0 = the higher index of a close pair)
 * mdb_system_set_atoms_lattice: the replicated crystal is generated straight into the handle (no host copy);
 *                               box = cell rows scaled by (nx, ny, nz), like build_crystal. */
int mdb_repeat_cell(double *new_pos, const double *old_box9, const double *old_pos, int n_old, int nx, int ny, int nz,
                    int num_t);
int mdb_transform_and_filter(const double *x, const double *y, const double *z, int N, const double *rotation9,
                             const double *center3, const double *target3, const double *coeffs, int nfaces,
                             double *out_pos, int *count, int num_t);
int mdb_filter_overlap_atom(const double *x, const double *y, const double *z, int N, const double *box9,
                            const double *origin3, const int *boundary3, double rc, unsigned char *keep, int num_t);
int mdb_system_set_atoms_lattice(mdb_system *s, const double *cell9, const double *basis_pos, int n_basis, int nx, int ny,
                                 int nz, const double *origin3, const int *boundary3);
int mdb_system_positions_device(mdb_system *s, double **dx, double **dy, double **dz, int *N);
int mdb_system_fetch_positions(mdb_system *s, double *x, double *y, double *z);
int mdb_system_filter_overlap(mdb_system *s, double rc, unsigned char *keep_host);
/* transform_and_filter on the atoms already held by the handle (the replicated lattice block of a polycrystal
 * build is generated once on the device and reused for every grain) */
int mdb_system_transform_and_filter(mdb_system *s, const double *rotation9, const double *center3,
                                    const double *target3, const double *coeffs, int nfaces, double *out_pos, int *count);
int mdb_system_acna(mdb_system *s, int *pattern_host);
int mdb_system_ids(mdb_system *s, int *pattern_host);
int mdb_system_csp(mdb_system *s, int nnei, double *csp_host);
int mdb_system_aja(mdb_system *s, int *aja_host);
/* Steinhardt q_l (+ w_l, w_l-hat) on the cached list; q_lm stay on the device for solid_liquid.
 * Host outputs may be NULL.  rc rule as in steinhardt_bond_orientation.py:238-245. */
int mdb_system_steinhardt(mdb_system *s, const int *llist, int ndeg, int nnn, double rc, int average, int wl,
                          int wlhat, int use_voronoi, const double *weight_host, double *qnarray_host,
                          double *qlm_r_host, double *qlm_i_host);
int mdb_system_solid_liquid(mdb_system *s, int q6index, double threshold, int n_bond, int use_voronoi, int nnn,
                            double rc, int *solidliquid_host, int *nbond_host);
/* RDF counts accumulated into g_host: list kernels (streaming = 0; types_host NULL -> single species)
 * or straight from positions (streaming = 1) */
int mdb_system_rdf(mdb_system *s, const int *types_host, int ntype, double rc, int nbin, int streaming,
                   double *g_host);
/* PTM on the cached (sorted, >= 18 wide) list; types_host may be NULL; outputs (n_rows, 8) / (n_rows, 18) */
int mdb_system_ptm(mdb_system *s, const char *structure, const int *types_host, double rmsd_threshold,
                   double *output_host, int *indices_host);
/* device pointers to the most recent int32 / f64 per-atom result */
/* list consumers of SURVEY.md 8f.1 on the cached list (results copied to the host arrays) */
int mdb_system_cnp(mdb_system *s, double rc, double *cnp_host);
int mdb_system_wcp(mdb_system *s, const int *types_host, int ntype, double *wcp_host);
int mdb_system_average_by_neighbor(mdb_system *s, double rc, const double *value_host, int include_self,
                                   double *value_ave_host);
/* cluster ids (1-based, numbered by smallest member index) on the cached list.  npair == 0: one cut-off rc
 * (get_cluster); npair > 0: type-pair cut-offs, bonds filtered like filter_by_type and then joined like
 * get_cluster_by_bond (types_host: n_local ints in the units of type1/type2). */
int mdb_system_cluster(mdb_system *s, double rc, const int *types_host, const int *type1, const int *type2,
                       const double *r, int npair, int *cluster_host, int *cluster_number);
/* structure entropy on the cached list; average_rc > 0 also fills entropy_ave_host (average_by_neighbor
 * with include_self, structure_entropy.py:133-145).  Either host pointer may be NULL. */
int mdb_system_structure_entropy(mdb_system *s, double rc, double sigma, int use_local_density, double volume,
                                 double average_rc, double *entropy_host, double *entropy_ave_host);
/* self-check of the correctly rounded small-integer division used by the Legendre recurrences: counts the
 * host values a[i] (uploaded) whose device quotient differs from a[i] / d.  Test hook. */
int mdb_system_check_small_division(mdb_system *s, const double *a_host, int n, int d, long long *mismatches);
/* atomic temperature on the cached list; vx, vy, vz, mass: n_local host doubles */
int mdb_system_atomic_temperature(mdb_system *s, const double *vx, const double *vy, const double *vz,
                                  const double *mass, double rc, double *T_host);
/* bond-length / bond-angle histograms and the per-triplet angular distribution on the cached list; the int32
 * host histograms are ACCUMULATED into, like the reference */
int mdb_system_bond_analysis(mdb_system *s, double delta_r, double delta_theta, double rc, int nbins,
                             int *bond_length_host, int *bond_angle_host);
int mdb_system_adf(mdb_system *s, double delta_theta, const double *rc_list, const int *pair_list, int npair,
                   const int *types_host, int nbins, int *bond_angle_host);
int mdb_system_result_device(mdb_system *s, int **i32, double **f64);

/* per-kernel device times (ms) of the most recent build_neighbor / fcna, measured with CUDA events */
int mdb_system_set_profiling(mdb_system *s, int on);
int mdb_system_last_times(mdb_system *s, float *t_binning_ms, float *t_neighbor_ms, float *t_cna_ms);

/* ---- Voronoi cells --------------------------------------------------------------------------
 * Orthogonal boxes with any mix of periodic / open axes (open axes end at the box faces, like voro++'s container_3d);
 * triclinic boxes with all three axes periodic (voro++'s container_triclinic is periodic: the reference's Python
 * side triples the open axes first, src/mdapy/voronoi.py:148-152, and so does mdapy_b200/voronoi.py).
 * mdb_get_voronoi_volume_number_radius <- _voronoi.get_voronoi_volume_number_radius(x, y, z, box, origin, boundary,
 *                                          volume, neighbor_number, cavity_radius, num_t)   (src/voronoi.cpp:16)
 * mdb_system_voronoi_neighbor + mdb_system_voronoi_fetch <- _voronoi.get_voronoi_neighbor(x, y, z, box, origin,
 *   boundary, a_face_area_threshold, r_face_area_threshold, num_t) -> (verlet_list, distance_list, face_area,
 *   neighbor_number)                                                                        (src/voronoi.cpp:307)
 *   The reference allocates (N, max faces) arrays itself; here the first call computes and reports that width M, the
 *   caller allocates, the second call copies.  Values as the reference's: wall faces and faces at or below the area
 *   threshold hold -1 / 10000.0 / 0.0 in place, neighbor_number counts every face, distances are minimum-image.
 *   Rows list the faces in this library's insertion order, not voro++'s vertex-table order (same set per row). */
int mdb_get_voronoi_volume_number_radius(const double *x, const double *y, const double *z, int N, const double *box9,
                                         const double *origin3, const int *boundary3, double *volume,
                                         int *neighbor_number, double *cavity_radius, int num_t);
int mdb_system_voronoi_volume(mdb_system *s, double *volume_host, int *neighbor_number_host, double *cavity_radius_host);
int mdb_system_voronoi_neighbor(mdb_system *s, double a_face_area_threshold, double r_face_area_threshold, int *M);
int mdb_system_voronoi_fetch(mdb_system *s, int *verlet_host, double *distance_host, double *face_area_host,
                             int *neighbor_number_host);

/* ================================================================================================
 * Section C: device group -- one process, several GPUs (mdapy_b200/csrc/group.cu)
 * ================================================================================================
 * The reference runs src/neighbor.cpp + src/cna.cpp as OpenMP loops over one address space; the group is the
 * multi-GPU form of that call for System(..., devices=[...]): ONE unpartitioned host frame in, labels in the
 * original atom order out.  Member d uploads the d-th contiguous chunk of x, y, z over its own PCIe link; a
 * routing kernel pushes every atom (and the boundary-plane ghosts) into the slab of the member that owns its x
 * cell plane with peer stores over NVLink; each member runs the fused neighbour + CNA kernel on its slab; the
 * labels travel back the same way.  No host-side partitioning, no NCCL, no PyTorch.  A device may be listed
 * more than once (several slabs on one GPU; used by the single-GPU tests).  A group works on one frame at a time:
 * calls on the same group must not overlap (different groups, and groups next to mdb_system handles, may). */
typedef struct mdb_group mdb_group;
int mdb_group_create(const int *devices, int ndev, mdb_group **out);
void mdb_group_destroy(mdb_group *g);
int mdb_group_size(mdb_group *g);
/* start uploading one host frame (pageable or page-locked); the arrays must stay valid until the next group
 * call returns */
int mdb_group_set_atoms(mdb_group *g, const double *x, const double *y, const double *z, int N, const double *box9,
                        const double *origin3, const int *boundary3);
/* FixedCNA labels of the uploaded frame (same values as mdb_compute_fcna on the list of mdb_build_neighbor_*),
 * original atom order, N ints.  *members = devices the frame ran on (1: too few x cell planes to decompose, or
 * the fused kernel declined the frame -- the chunks were gathered on the first member with peer copies). */
int mdb_group_fused_cna(mdb_group *g, double rc, int *pattern_host, int *members);
/* host-clock phase times of the last frame, ms: upload issue, route, slab compute, label push, label download,
 * set_atoms-to-labels total */
int mdb_group_last_times(mdb_group *g, float *ms6);
int mdb_group_member_atoms(mdb_group *g, int member, int *n_owned, int *n_local);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MDAPY_B200_H */
