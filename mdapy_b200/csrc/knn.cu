// mdapy_b200/csrc/knn.cu
//
// Exact k-nearest-neighbour search (k <= 24) on a uniform cell grid.  Replaces
// the reference's kd-trees (src/fast_knn.cpp:208-568 KdTree, 588-794
// OrthoKdTree, 846-916 knn).  Semantics kept:
//   * atoms and queries are wrapped with the reference's own arithmetic
//     (ortho: fast_knn.cpp:689-703; triclinic: wrap_triclinic 86-99, which
//     ignores the origin),
//   * periodic images are distinct neighbours: the candidate set is every
//     (atom j, image shift s) with |s_d| <= nimages on periodic axes
//     (build_pbc_shifts 801-841), d2 = |a_j - (q - s)|^2 in that op order,
//   * self is skipped iff idx == self && d2 == 0.0 (641, 538),
//   * output ascending in d2, distance = sqrt(d2), short rows -1 / -1.0.
// Among EQUAL d2 the reference's order depends on libstdc++'s nth_element
// traversal; here ties keep visit order (documented tie-sensitivity).
//
// Search: rings of cells around the query's cell; ring r is final once the
// k-th best d2 <= (r * min perpendicular cell width)^2, or when every axis has
// run out of allowed images.
#include "internal.cuh"
#include <algorithm>

namespace {

constexpr int KNN_MAX_K = 24;

struct KnnGrid {
    int n[3];
    int total;
    int img[3];        // allowed image shifts per axis (0 on open axes)
    double lo[3];      // lower bound of the gridded coordinate (reduced for triclinic, absolute-origin for ortho)
    double scale[3];   // cells per unit of the gridded coordinate
    double w_perp[3];  // perpendicular cell width in length units
};

__device__ __forceinline__ unsigned long long f2key(double v)
{
    unsigned long long b = __double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static inline double key2f(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

// gridded coordinate of a wrapped position
__device__ __forceinline__ void grid_coord(const DBox &b, double x, double y, double z, double &fx, double &fy,
                                           double &fz)
{
    if (b.triclinic) {
        fx = x * b.hinv[0] + y * b.hinv[3] + z * b.hinv[6];
        fy = x * b.hinv[1] + y * b.hinv[4] + z * b.hinv[7];
        fz = x * b.hinv[2] + y * b.hinv[5] + z * b.hinv[8];
    } else {
        fx = x - b.origin[0];
        fy = y - b.origin[1];
        fz = z - b.origin[2];
    }
}

// wrap with the reference's arithmetic and reduce min/max of the gridded coordinate
__global__ void __launch_bounds__(256) k_knn_wrap(const double *__restrict__ x, const double *__restrict__ y,
                                                  const double *__restrict__ z, int N, DBox b,
                                                  double *__restrict__ wx, double *__restrict__ wy,
                                                  double *__restrict__ wz, unsigned long long *__restrict__ mm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double fx = 0, fy = 0, fz = 0;
    const bool live = i < N;
    if (live) {
        double px = x[i], py = y[i], pz = z[i];
        if (b.triclinic) {
            // fast_knn.cpp:86-99
            double r[3];
            r[0] = px * b.hinv[0] + py * b.hinv[3] + pz * b.hinv[6];
            r[1] = px * b.hinv[1] + py * b.hinv[4] + pz * b.hinv[7];
            r[2] = px * b.hinv[2] + py * b.hinv[5] + pz * b.hinv[8];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (b.pbc[d]) {
                    const double s = floor(r[d]);
                    if (s != 0.0) {
                        r[d] -= s;
                        px -= s * b.h[d * 3 + 0];
                        py -= s * b.h[d * 3 + 1];
                        pz -= s * b.h[d * 3 + 2];
                    }
                }
            }
        } else {
            // fast_knn.cpp:682-703 (invL = 1.0 / L on the host side of the reference too)
            if (b.pbc[0]) {
                const double s = floor((px - b.origin[0]) * b.hinv[0]);
                if (s != 0.0) px -= s * b.h[0];
            }
            if (b.pbc[1]) {
                const double s = floor((py - b.origin[1]) * b.hinv[4]);
                if (s != 0.0) py -= s * b.h[4];
            }
            if (b.pbc[2]) {
                const double s = floor((pz - b.origin[2]) * b.hinv[8]);
                if (s != 0.0) pz -= s * b.h[8];
            }
        }
        wx[i] = px;
        wy[i] = py;
        wz[i] = pz;
        grid_coord(b, px, py, pz, fx, fy, fz);
    }
    unsigned long long k[6];
    const double f[3] = {fx, fy, fz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        k[2 * d] = live ? f2key(f[d]) : ~0ull;    // min
        k[2 * d + 1] = live ? f2key(f[d]) : 0ull;  // max
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            k[2 * d] = min(k[2 * d], __shfl_xor_sync(0xffffffffu, k[2 * d], o));
            k[2 * d + 1] = max(k[2 * d + 1], __shfl_xor_sync(0xffffffffu, k[2 * d + 1], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(mm + 2 * d, k[2 * d]);
            atomicMax(mm + 2 * d + 1, k[2 * d + 1]);
        }
    }
}

__global__ void __launch_bounds__(256) k_knn_cell(const double *__restrict__ wx, const double *__restrict__ wy,
                                                  const double *__restrict__ wz, int N, DBox b, KnnGrid g,
                                                  int *__restrict__ cell_of_atom, int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double fx, fy, fz;
    grid_coord(b, wx[i], wy[i], wz[i], fx, fy, fz);
    const int ic = clampi((int)floor((fx - g.lo[0]) * g.scale[0]), 0, g.n[0] - 1);
    const int jc = clampi((int)floor((fy - g.lo[1]) * g.scale[1]), 0, g.n[1] - 1);
    const int kc = clampi((int)floor((fz - g.lo[2]) * g.scale[2]), 0, g.n[2] - 1);
    const int c = (ic * g.n[1] + jc) * g.n[2] + kc;
    cell_of_atom[i] = c;
    atomicAdd(&count[c], 1);
}

__device__ __forceinline__ int floor_div(int a, int n)
{
    int q = a / n;
    if ((a % n) < 0) --q;
    return q;
}

__device__ __forceinline__ double2 ldg2(const SortedAtom *p, int half) { return __ldg(reinterpret_cast<const double2 *>(p) + half); }

__global__ void __launch_bounds__(128) k_knn_query(const SortedAtom *__restrict__ sorted,
                                                   const int *__restrict__ cell_start, int N, DBox b, KnnGrid g,
                                                   int k, int *__restrict__ out_idx, double *__restrict__ out_d)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double2 m0 = ldg2(sorted + s, 0), m1 = ldg2(sorted + s, 1);
    const double qx = m0.x, qy = m0.y, qz = m1.x;
    const int self = __double2loint(m1.y);
    const int cell = __double2hiint(m1.y);
    int c[3];
    c[2] = cell % g.n[2];
    c[1] = (cell / g.n[2]) % g.n[1];
    c[0] = cell / (g.n[2] * g.n[1]);

    double bd[KNN_MAX_K];
    int bi[KNN_MAX_K];
    int nb = 0;

    int amin[3], amax[3];  // allowed unwrapped cell range per axis
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        amin[d] = -g.n[d] * g.img[d];
        amax[d] = g.n[d] * (1 + g.img[d]) - 1;
    }
    int plo[3] = {1, 1, 1}, phi[3] = {0, 0, 0};  // previous (already scanned) block: empty
    for (int r = 1;; ++r) {
        int lo[3], hi[3];
        bool exhausted = true;
        double bound = 1e300;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = max(c[d] - r, amin[d]);
            hi[d] = min(c[d] + r, amax[d]);
            const bool both_clipped = (c[d] - r < amin[d]) && (c[d] + r > amax[d]);
            if (!both_clipped) {
                exhausted = false;
                bound = fmin(bound, r * g.w_perp[d]);
            }
        }
        for (int ii = lo[0]; ii <= hi[0]; ++ii) {
            const int m0i = floor_div(ii, g.n[0]);
            const int ci = ii - m0i * g.n[0];
            const bool in0 = ii >= plo[0] && ii <= phi[0];
            for (int jj = lo[1]; jj <= hi[1]; ++jj) {
                const int m1i = floor_div(jj, g.n[1]);
                const int cj = jj - m1i * g.n[1];
                const bool in1 = in0 && jj >= plo[1] && jj <= phi[1];
                for (int kk = lo[2]; kk <= hi[2]; ++kk) {
                    if (in1 && kk >= plo[2] && kk <= phi[2]) continue;
                    const int m2i = floor_div(kk, g.n[2]);
                    const int ck = kk - m2i * g.n[2];
                    // image shift, fast_knn.cpp:824-832
                    double sx, sy, sz;
                    if (b.triclinic) {
                        sx = m0i * b.h[0] + m1i * b.h[3] + m2i * b.h[6];
                        sy = m0i * b.h[1] + m1i * b.h[4] + m2i * b.h[7];
                        sz = m0i * b.h[2] + m1i * b.h[5] + m2i * b.h[8];
                    } else {
                        sx = m0i * b.h[0];
                        sy = m1i * b.h[4];
                        sz = m2i * b.h[8];
                    }
                    // image m of a cell sits at a + m*L; the reference moves the QUERY instead:
                    // q = qw - s (fast_knn.cpp:764-768, 421-425)
                    const double tx = qx - sx, ty = qy - sy, tz = qz - sz;
                    const int cc = (ci * g.n[1] + cj) * g.n[2] + ck;
                    const int beg = __ldg(cell_start + cc), end = __ldg(cell_start + cc + 1);
                    for (int q = beg; q < end; ++q) {
                        const double2 a0 = ldg2(sorted + q, 0), a1 = ldg2(sorted + q, 1);
                        const double dx = a0.x - tx, dy = a0.y - ty, dz = a1.x - tz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        const int j = __double2loint(a1.y);
                        if (j == self && d2 == 0.0) continue;
                        if (nb == k && !(d2 < bd[k - 1])) continue;
                        int pos = nb < k ? nb : k - 1;
                        while (pos > 0 && bd[pos - 1] > d2) {
                            bd[pos] = bd[pos - 1];
                            bi[pos] = bi[pos - 1];
                            --pos;
                        }
                        bd[pos] = d2;
                        bi[pos] = j;
                        if (nb < k) ++nb;
                    }
                }
            }
        }
        if (exhausted) break;
        const double safe = bound * (1.0 - 1e-12);
        if (nb == k && bd[k - 1] <= safe * safe) break;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            plo[d] = lo[d];
            phi[d] = hi[d];
        }
    }
    int *orow = out_idx + (size_t)self * k;
    double *drow = out_d + (size_t)self * k;
    for (int t = 0; t < k; ++t) {
        orow[t] = t < nb ? bi[t] : -1;
        drow[t] = t < nb ? sqrt(bd[t]) : -1.0;
    }
}

__global__ void __launch_bounds__(256) k_fill_int(int *__restrict__ p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// Cell width of the search grid in units of the expected k-th neighbour distance.  0.6: the second ring of
// cells (5^3 cells, (3 w)^3 = 5.8 r_k^3... per axis 2.5 w) already bounds the k-th distance, and the scanned
// volume is ~2.3x smaller than with one ring of cells 1.1 r_k wide.  MDB_KNN_CELL overrides (experiments).
static double knn_cell_factor()
{
    const char *env = getenv("MDB_KNN_CELL");
    if (env) {
        const double v = atof(env);
        if (v > 0.05 && v < 10.0) return v;
    }
    return 0.6;
}

void launch_knn(MdbSystem &s, int k)
{
    MDB_REQUIRE(k >= 1 && k <= KNN_MAX_K, MDB_ERR_VALUE, "k must be in [1, %d], got %d.", KNN_MAX_K, k);
    MDB_REQUIRE(s.N > 0 && s.x, MDB_ERR_STATE, "no atoms uploaded");
    // Decomposed frame: the search runs over ALL local atoms (owned + ghosts) with the global box, so
    // wrapped coordinates, image shifts and distances are the single-GPU ones; a row is complete iff its
    // k-th distance does not reach past the halo (checked by mdapy_b200/distributed.py).
    const int N = s.N;
    cudaStream_t st = s.stream;
    const DBox &b = s.box;
    double *wx = s.wx.ensure<double>(N), *wy = s.wy.ensure<double>(N), *wz = s.wz.ensure<double>(N);
    unsigned long long *mm = s.scratch2.ensure<unsigned long long>(8);
    const unsigned long long init[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int nbk = (N + 255) / 256;
    MDB_LAUNCH(k_knn_wrap, nbk, 256, 0, st, s.x, s.y, s.z, N, b, wx, wy, wz, mm);
    unsigned long long hmm[6];
    CUDA_TRY(cudaMemcpyAsync(hmm, mm, sizeof(hmm), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    // image count per periodic axis: fast_knn.cpp:806-816
    int nimages = 1;
    if (b.any_pbc) {
        const long long cl = std::min<long long>(std::max<long long>(N, 50), 200);
        nimages = (int)(200 / cl);
        if (nimages < 1) nimages = 1;
        if (nimages < 2 && b.triclinic) nimages = 2;
    }
    KnnGrid g;
    double range[3], lo[3];
    for (int d = 0; d < 3; ++d) {
        const double fmin_ = key2f(hmm[2 * d]), fmax_ = key2f(hmm[2 * d + 1]);
        const double full = b.triclinic ? 1.0 : b.h[d * 4];
        if (b.pbc[d]) {
            lo[d] = 0.0;
            range[d] = full;
        } else {
            lo[d] = fmin_;
            range[d] = fmax_ - fmin_;
            if (!(range[d] > 1e-12 * (std::fabs(full) + 1e-300))) range[d] = 1e-9 * std::fabs(full) + 1e-300;
        }
        g.img[d] = b.pbc[d] ? nimages : 0;
    }
    // occupied volume and expected distance of the k-th neighbour
    double frac = 1.0;
    double len[3];
    for (int d = 0; d < 3; ++d) {
        const double full = b.triclinic ? 1.0 : b.h[d * 4];
        frac *= range[d] / full;
        len[d] = std::fabs(b.thick[d]) * (range[d] / std::fabs(full));  // perpendicular extent along d
    }
    // a slab occupies only local_frac of the (periodic) box: keep the cell size tied to the LOCAL density
    const double vol = std::fabs(dbox_volume(b)) * std::fabs(frac) * (s.local_frac > 0 && s.local_frac < 1 ? s.local_frac : 1.0);
    const double rho = vol > 0 ? N / vol : 1.0;
    double wt = knn_cell_factor() * std::cbrt(3.0 * (k + 1) / (4.0 * 3.14159265358979323846 * rho));
    // thin (quasi 2-D / 1-D) extents: do not let a degenerate axis inflate the density estimate
    for (int d = 0; d < 3; ++d)
        if (len[d] < wt && !b.pbc[d]) {
            double area = 1.0;
            int nd = 0;
            for (int e = 0; e < 3; ++e)
                if (e != d && len[e] >= wt) {
                    area *= len[e];
                    ++nd;
                }
            if (nd == 2) wt = std::max(wt, 1.1 * std::sqrt((k + 1) / (3.14159265358979323846 * N / area)));
        }
    double total = 1.0;
    for (int d = 0; d < 3; ++d) {
        int n = (int)std::floor(len[d] / wt);
        if (n < 1) n = 1;
        if (n > 2048) n = 2048;
        g.n[d] = n;
        total *= n;
    }
    while (total > 4.0 * N + 64) {  // keep the grid O(N)
        int dmax = 0;
        for (int d = 1; d < 3; ++d)
            if (g.n[d] > g.n[dmax]) dmax = d;
        total /= g.n[dmax];
        g.n[dmax] = std::max(1, g.n[dmax] / 2);
        total *= g.n[dmax];
    }
    g.total = g.n[0] * g.n[1] * g.n[2];
    for (int d = 0; d < 3; ++d) {
        g.lo[d] = lo[d];
        g.scale[d] = g.n[d] / range[d];
        g.w_perp[d] = len[d] / g.n[d];
    }
    int *count = s.cell_count.ensure<int>((size_t)g.total + 1);
    int *cell_of_atom = s.perm_tmp.ensure<int>(N);
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)g.total + 1), st));
    MDB_LAUNCH(k_knn_cell, nbk, 256, 0, st, wx, wy, wz, N, b, g, cell_of_atom, count);
    finish_binning(s, g.total, wx, wy, wz);
    s.bin_rc = -1.0;  // the sorted copy now holds kNN-wrapped coordinates

    int *idx = s.verlet.ensure<int>((size_t)N * k);
    double *dist = s.dist.ensure<double>((size_t)N * k);
    int *nn = s.nn.ensure<int>(N);
    MDB_LAUNCH(k_knn_query, (N + 127) / 128, 128, 0, st, s.sorted.as<SortedAtom>(), s.cell_start.as<int>(), N, b, g,
               k, idx, dist);
    MDB_LAUNCH(k_fill_int, nbk, 256, 0, st, nn, N, k);
    CUDA_TRY(cudaGetLastError());
    s.M = k;
    s.max_count = k;
    s.list_kind = LIST_KNN;
    s.has_dist = true;
    s.list_rc = -1.0;
}
