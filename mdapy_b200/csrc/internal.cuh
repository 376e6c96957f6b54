// mdapy_b200/csrc/internal.cuh -- device-side state behind the C ABI (include/mdapy_b200.h).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <utility>
#include <stdexcept>
#include <string>
#include "box.cuh"
#include "../../include/mdapy_b200.h"

enum MdbStatusInternal {
    MDB_OK_ = 0,
    MDB_ERR_CUDA_ = 1,     // a CUDA runtime call failed          -> RuntimeError
    MDB_ERR_VALUE_ = 2,    // invalid argument / max_neigh small  -> ValueError
    MDB_ERR_BOX_ = 3,      // zero-volume cell (box.h:185)        -> RuntimeError
    MDB_ERR_STATE_ = 4,    // call order (no list built, ...)     -> RuntimeError
};

void mdb_set_error(const char *fmt, ...);

// every kernel launch goes through here so the library can report how many of
// its own kernels ran (bench.py "gpu_launches")
extern std::atomic<long long> g_mdb_launches;
#define MDB_LAUNCH(kern, grid, block, smem, st, ...)              \
    do {                                                          \
        kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);     \
        ++g_mdb_launches;                                         \
    } while (0)

struct MdbError {
    int code;
};

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            mdb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,    \
                          cudaGetErrorString(_e));                                                  \
            throw MdbError{MDB_ERR_CUDA};                                                           \
        }                                                                                           \
    } while (0)

#define MDB_REQUIRE(cond, code, ...)                                                                \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            mdb_set_error(__VA_ARGS__);                                                             \
            throw MdbError{code};                                                                   \
        }                                                                                           \
    } while (0)

// Device / pinned-host block caches (capi.cu).  A System is created per frame by the Python layer
// (system.py mirrors the reference's one-System-per-file usage); cudaMalloc/cudaFree of ~20 GB of
// buffers per frame would dominate the end-to-end time, so released blocks are parked per device and
// reused by the next System.  mdb_trim_cache() returns everything to the driver.
// Frees are STREAM ORDERED: a released block carries an event recorded on the stream that last used it,
// and whoever takes the block from the cache waits for that event on its own stream (or on the host when
// it has none) -- work still queued on the old owner's stream can never overlap the new owner's.
void *mdb_pool_alloc(size_t bytes, size_t *got, cudaStream_t user);
void mdb_pool_free(void *p, size_t bytes, cudaStream_t last_user);

// grow-only device buffer; allocation cost is paid once per high-water mark
struct DevBuf {
    void *p{nullptr};
    size_t cap{0};
    const cudaStream_t *owner{nullptr};  // the stream of the handle this buffer belongs to (MdbSystem::bind_buffers)
    cudaStream_t stream() const { return owner ? *owner : nullptr; }
    template <class T> T *ensure(size_t count)
    {
        const size_t bytes = count * sizeof(T);
        if (bytes > cap) {
            release();
            p = mdb_pool_alloc(bytes + bytes / 16 + 256, &cap, stream());
        }
        return static_cast<T *>(p);
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
    void release()
    {
        if (p) mdb_pool_free(p, cap, stream());
        p = nullptr;
        cap = 0;
    }
};

// 32-byte sorted record: one atom of the cell-ordered copy (raw coordinates,
// original index).  32-byte alignment makes any run of records a legal source
// for cp.async.bulk (TMA 1-D) and two 16-byte vector loads otherwise.
struct __align__(32) SortedAtom {
    double x, y, z;
    int idx;
    int cell;
};

enum ListKind { LIST_NONE = 0, LIST_CUTOFF = 1, LIST_KNN = 2 };

struct MdbSystem {
    int device{0};
    cudaStream_t stream{nullptr};
    bool own_stream{false};
    cudaStream_t copy_stream{nullptr};  // device -> host copies that overlap the next chunk of a kernel
    cudaEvent_t chunk_ev[2]{};

    // atoms (raw coordinates, SoA as mdapy's polars columns: neighbor.py:104-106)
    int N{0};
    DBox box{};
    bool has_box{false};
    DevBuf bx, by, bz;             // owned copies (host upload path)
    const double *x{nullptr}, *y{nullptr}, *z{nullptr};  // device pointers in use (owned or borrowed)
    // rows of every list / per-atom output.  n_rows == N on one GPU; in a decomposed frame the
    // first n_rows atoms are owned, the rest are ghosts that only serve as neighbours.
    int n_rows{0};
    const int *gid{nullptr};  // optional global ids (order inside a cell, exported lists)
    int slab_x0{0}, slab_nx{0};  // stored window of global x cell planes (0,0 = whole grid)
    double local_frac{1.0};      // fraction of the box volume the local atoms occupy (density hint, kNN grid)

    // cell binning for a given rc
    double bin_rc{-1.0};
    CellGrid grid{};
    DevBuf cell_count, cell_start, perm, perm_tmp, sorted, scan_tmp, big_cells, counters;

    // neighbour list (device resident)
    int list_kind{LIST_NONE};
    double list_rc{-1.0};
    int M{0};
    int max_count{0};
    bool has_dist{true};  // false: the caller handed in a bare verlet list; distances are recomputed on first use
    DevBuf verlet, dist, nn, verlet_tmp, dist_tmp;
    // width of the previous automatic build on this handle (frames of a trajectory reuse one handle: the
    // next frame starts from this width instead of sampling tiles again)
    double hint_rc{-1.0};
    int hint_M{0};
    bool hint_uniform{false};

    // per-atom outputs kept on device until fetched
    DevBuf out_i32, out_f64, out_f64b, out_f64c, scratch, scratch2;
    DevBuf wx, wy, wz;  // kNN: wrapped coordinates (fast_knn.cpp wrap arithmetic)
    // Voronoi cells (voronoi.cu): raw (neighbour id, face area) rows of width vor_W, face counts, and the
    // reference-shaped arrays of width vor_M built from them
    DevBuf vor_id, vor_area, vor_nn, vor_verlet, vor_dist, vor_farea;
    int vor_W{0}, vor_M{0};
    // Steinhardt state kept for identifySolidLiquid / repeated reads
    DevBuf qlm_r, qlm_i, qn, types, weight, ptm_out, ptm_idx;
    int sbo_ndeg{0}, sbo_nz{0}, sbo_ncol{0};

    // timing of the last call, per kernel (ms), filled when profiling is on
    bool profile{false};
    float t_bin{0}, t_neigh{0}, t_cna{0};
    cudaEvent_t ev[4]{};

    // every device buffer of the handle (stream binding, destruction)
    template <class F> void for_each_buffer(F f)
    {
        DevBuf *all[] = {&bx, &by, &bz, &cell_count, &cell_start, &perm, &perm_tmp, &sorted, &scan_tmp, &big_cells,
                         &counters, &verlet, &dist, &nn, &verlet_tmp, &dist_tmp, &out_i32, &out_f64, &out_f64b,
                         &out_f64c, &scratch, &scratch2, &wx, &wy, &wz, &qlm_r, &qlm_i, &qn, &types, &weight,
                         &ptm_out, &ptm_idx, &vor_id, &vor_area, &vor_nn, &vor_verlet, &vor_dist, &vor_farea};
        for (DevBuf *b : all) f(*b);
    }
    void bind_buffers()
    {
        for_each_buffer([this](DevBuf &b) { b.owner = &stream; });
    }
};

// ---- host -> device upload of caller columns (staging.cu): pageable sources are staged through page-locked ring
// buffers by `threads` host threads, page-locked sources go down as one asynchronous copy
void mdb_h2d(int n, void *const *dst, const void *const *src, const size_t *bytes, cudaStream_t consumer, int threads);
int mdb_upload_threads();

// ---- kernels launchers (one per .cu) ---------------------------------------
void launch_binning(MdbSystem &s, double rc);
void finish_binning(MdbSystem &s, int nc, const double *X, const double *Y, const double *Z);
void launch_knn(MdbSystem &s, int k);
int launch_voronoi(MdbSystem &s, bool want_rows, double *volume, int *nfaces, double *radius);
void launch_voronoi_rows(MdbSystem &s, const int *nfaces, int M, double a_thr, double r_thr, int *verlet, double *dist,
                         double *area);
void launch_cell_planes(const double *x, const double *y, const double *z, int N, const DBox &b, const CellGrid &g,
                        int *plane, cudaStream_t st);
void launch_translate_ids(MdbSystem &s, const int *local_ids, int *global_ids, size_t n);
void launch_slab_pack(const double *x, const double *y, const double *z, const int *gid, int N, const DBox &b,
                      const CellGrid &g, int lo, int hi, int halo, double *left, double *right, int cap, int *counts,
                      cudaStream_t st);
void launch_slab_unpack(const double *a, const double *b, int cap, double *x, double *y, double *z, int *gid, int n_owned,
                        int room, int *total, cudaStream_t st);
void launch_neighbor(MdbSystem &s, double rc, int M, bool count_only);
void launch_compact_rows(MdbSystem &s, int M_from, int M_to);
bool tiled_neighbor_plan(const MdbSystem &s, int &T);
void launch_neighbor_tiled(MdbSystem &s, double rc, int M, int T, bool count_only, int sample_stride);
int neighbor_tiled_max(MdbSystem &s, int *min_count = nullptr);
bool launch_fused_cna(MdbSystem &s, double rc, int *pattern, int *n_fallback);
void launch_sort_rows(MdbSystem &s, int *verlet, double *dist, int N, int M, int k);
void launch_fcna(MdbSystem &s, const int *verlet, const int *nn, int M, double rc, int *pattern, int first = 0,
                 int count = -1);
void launch_acna(MdbSystem &s, const int *verlet, int M, int *pattern);
void launch_ids(MdbSystem &s, const int *verlet, int M, int *second_out, int *pattern);
void launch_csp(MdbSystem &s, const int *verlet, int M, int nnei, double *csp);
void launch_aja(MdbSystem &s, const int *verlet, int M, const double *dist, int Md, int *aja);
void launch_steinhardt(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M,
                       const double *weight, const int *llist, int ndeg, int nnn, int lmax, bool wl, bool wlhat,
                       bool average, bool use_voronoi, double rc, bool use_weight, double *qr, double *qi, double *qn);
void launch_solid_liquid(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, int q6index,
                         const double *Q6, const double *qr, const double *qi, int ndeg, int nz, double threshold,
                         int n_bond, bool use_voronoi, int nnn, double rc, int *solid, int *nbond);
void launch_rdf_list(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int N, int M,
                     const int *types, int ntype, double rc, int nbin, double *g);
void launch_rdf_streaming(MdbSystem &s, const int *types, int ntype, double rc, int nbin, double *g);
void launch_cnp(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, double *cnp);
void launch_wcp_counts(MdbSystem &s, const int *verlet, const int *nn, int M, const int *types, int T,
                       unsigned long long *counts);
void launch_average_by_neighbor(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc,
                                const double *value, bool include_self, double *out);
void device_exclusive_scan(MdbSystem &s, const int *in, int *out, int n);
void launch_filter_by_type(MdbSystem &s, int *verlet, const double *dist, const int *nn, int M, const int *types,
                           const int *t1, const int *t2, const double *r, int npair);
int launch_cluster(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, int *cluster);
void launch_structure_entropy(MdbSystem &s, const double *dist, const int *nn, int M, double rc, double sigma,
                              bool use_local_density, double volume, double *entropy);
long long sbo_div_small_mismatches(MdbSystem &s, const double *a_dev, int n, int d);
void launch_atomic_temperature(MdbSystem &s, const int *verlet, const double *dist, int M, const double *vx,
                               const double *vy, const double *vz, const double *mass, double rc, double *T);
void launch_bond_hist(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double delta_r,
                      double delta_theta, double rc, int nbins, unsigned long long *hist);
void launch_adf_hist(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double delta_theta,
                     const double *rcs, const int *pairs, int npair, const int *types, int nbins,
                     unsigned long long *hist);
void launch_wrap_positions(MdbSystem &s, double *x, double *y, double *z, int N);
int ptm_parse_flags(const char *structure);
void launch_ptm(MdbSystem &s, int flags, const int *verlet, int M, const int *types, double rmsd_threshold,
                double *output, int ocols, int *indices, int icols);
void launch_repeat_cell(MdbSystem &s, const double *old_pos_dev, int n_old, const DBox &b, int nx, int ny, int nz,
                        double *ox, double *oy, double *oz, double *o3);
int launch_transform_and_filter(MdbSystem &s, const double *x, const double *y, const double *z, int N, const double *R9,
                                const double *center3, const double *target3, const double *planes_dev, int nfaces,
                                double *out3);
void launch_filter_overlap(MdbSystem &s, double rc, unsigned char *keep);
void launch_chill_plus(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, int *pattern);
int launch_build_bond(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, const int *types_dev,
                      const double *cutoff_dev, int ntype, int **bonds_dev);
void launch_planar_faults(MdbSystem &s, const int *type, int N, const int *ptm_idx, int stride, int col0, int order,
                          bool identify_esf, int *fault);
void launch_types_from_ptm_output(MdbSystem &s, const double *out, int ocols, int N, int *type);
int device_max_int(MdbSystem &s, const int *v, size_t n);
int device_min_int(MdbSystem &s, const int *v, size_t n);
