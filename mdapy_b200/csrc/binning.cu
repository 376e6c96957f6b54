// mdapy_b200/csrc/binning.cu
//
// Atom -> cell binning.  Replaces the reference's serial linked-list build
// (src/neighbor.cpp:64-100 build_cell) with a counting sort on the device:
//   1. k_cell_count   cell id per atom (same wrap + floor arithmetic as
//                     neighbor.cpp:30-62) and a per-cell population histogram
//   2. exclusive scan cell_start[c] (hand-written 3-phase block scan)
//   3. k_scatter      atoms to their cell segment
//   4. k_order_cells  ascending original index inside every cell, so that a
//                     backwards walk reproduces the reference's head-insertion
//                     chains (descending index, neighbor.cpp:97-98) exactly
//   5. k_gather       32-byte SortedAtom records in cell order
// The sorted copy is what every neighbour-type kernel streams.
#include "internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SMALL_CELL = 48;  // insertion-sorted by one thread; larger cells go to the rank-sort path

__global__ void __launch_bounds__(256) k_cell_count(const double *__restrict__ x, const double *__restrict__ y,
                                                    const double *__restrict__ z, int N, DBox box, CellGrid g,
                                                    int *__restrict__ cell_of_atom, int *__restrict__ count,
                                                    int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[i], yi = y[i], zi = z[i];
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_of(box, g, xi, yi, zi, ic, jc, kc);
    int c = cell_linear(g, ic, jc, kc);
    if (c < 0) {  // atom outside the stored slab window: caller error, flagged and parked in cell 0
        atomicAdd(bad, 1);
        c = 0;
    }
    cell_of_atom[i] = c;
    atomicAdd(&count[c], 1);
}

// exclusive scan of one tile per block; tile totals to sums[]
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int *__restrict__ in, int *__restrict__ out, int n,
                                                             int *__restrict__ sums)
{
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        local += v[k];
    }
    // inclusive warp scan of the per-thread totals
    int incl = local;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - local + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && sums) sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int *__restrict__ out, int n, const int *__restrict__ offs)
{
    const int add = offs[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// in -> out exclusive scan over n ints, tmp holds the tile-sum pyramid
void exclusive_scan(const int *in, int *out, int n, int *tmp, cudaStream_t st)
{
    const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 1) {
        MDB_LAUNCH(k_scan_tiles, 1, SCAN_THREADS, 0, st, in, out, n, nullptr);
        return;
    }
    int *sums = tmp;
    MDB_LAUNCH(k_scan_tiles, tiles, SCAN_THREADS, 0, st, in, out, n, sums);
    exclusive_scan(sums, sums, tiles, tmp + ((tiles + 31) / 32) * 32, st);  // in place is safe: tile reads precede writes
    MDB_LAUNCH(k_scan_add, tiles, SCAN_THREADS, 0, st, out, n, sums);
}

size_t scan_tmp_ints(int n)
{
    size_t total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += ((n + 31) / 32) * 32;
    }
    return total + 32;
}

__global__ void __launch_bounds__(256) k_scatter(const int *__restrict__ cell_of_atom, int N,
                                                 const int *__restrict__ cell_start, int *__restrict__ cursor,
                                                 int *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = cell_of_atom[i];
    const int slot = cell_start[c] + atomicAdd(&cursor[c], 1);
    perm[slot] = i;
}

// ascending original index within each cell; big cells are deferred
// order key: the atom's own index, or its global id when the frame is decomposed
__device__ __forceinline__ int okey(const int *__restrict__ gid, int i) { return gid ? gid[i] : i; }

__global__ void __launch_bounds__(256) k_order_cells(const int *__restrict__ cell_start, int ncell, int *__restrict__ perm,
                                                     int *__restrict__ big_cells, int *__restrict__ n_big,
                                                     const int *__restrict__ gid)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    const int n = e - b;
    if (n <= 1) return;
    if (n > SMALL_CELL) {
        big_cells[atomicAdd(n_big, 1)] = c;
        return;
    }
    for (int a = b + 1; a < e; ++a) {
        const int v = perm[a];
        const int kv = okey(gid, v);
        int q = a - 1;
        while (q >= b && okey(gid, perm[q]) > kv) {
            perm[q + 1] = perm[q];
            --q;
        }
        perm[q + 1] = v;
    }
}

// rank sort (all keys distinct) of one over-full cell per block
// (grid-stride over the deferred list; the list length stays on the device)
__global__ void __launch_bounds__(256) k_order_big(const int *__restrict__ cell_start, const int *__restrict__ big_cells,
                                                   const int *__restrict__ n_big, const int *__restrict__ perm,
                                                   int *__restrict__ perm_out, const int *__restrict__ gid)
{
    for (int w = blockIdx.x; w < *n_big; w += gridDim.x) {
        const int c = big_cells[w];
        const int b = cell_start[c], e = cell_start[c + 1];
        for (int a = b + threadIdx.x; a < e; a += blockDim.x) {
            const int v = perm[a];
            const int kv = okey(gid, v);
            int rank = 0;
            for (int q = b; q < e; ++q) rank += (okey(gid, perm[q]) < kv);
            perm_out[b + rank] = v;
        }
    }
}

__global__ void __launch_bounds__(256) k_copy_big(const int *__restrict__ cell_start, const int *__restrict__ big_cells,
                                                  const int *__restrict__ n_big, int *__restrict__ perm,
                                                  const int *__restrict__ perm_out)
{
    for (int w = blockIdx.x; w < *n_big; w += gridDim.x) {
        const int c = big_cells[w];
        const int b = cell_start[c], e = cell_start[c + 1];
        for (int a = b + threadIdx.x; a < e; a += blockDim.x) perm[a] = perm_out[a];
    }
}

__global__ void __launch_bounds__(256) k_gather(const double *__restrict__ x, const double *__restrict__ y,
                                                const double *__restrict__ z, const int *__restrict__ perm,
                                                const int *__restrict__ cell_of_atom, int N,
                                                SortedAtom *__restrict__ sorted)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int i = perm[s];
    SortedAtom a;
    a.x = x[i];
    a.y = y[i];
    a.z = z[i];
    a.idx = i;
    a.cell = cell_of_atom[i];
    // two 16-byte stores
    double2 *dst = reinterpret_cast<double2 *>(sorted + s);
    dst[0] = make_double2(a.x, a.y);
    double2 hi;
    hi.x = a.z;
    hi.y = __hiloint2double(a.cell, a.idx);
    dst[1] = hi;
}

// global x cell plane of every atom (ownership / ghost selection of a decomposed frame)
__global__ void __launch_bounds__(256) k_cell_planes(const double *__restrict__ x, const double *__restrict__ y,
                                                     const double *__restrict__ z, int N, DBox box, CellGrid g,
                                                     int *__restrict__ plane)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[i], yi = y[i], zi = z[i];
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_of(box, g, xi, yi, zi, ic, jc, kc);
    plane[i] = ic;
}

// Halo exchange of a decomposed frame, device side (no host round trip): atoms of the first / last `halo` owned
// planes are appended to two fixed-capacity send buffers of (x, y, z, global id) rows; row 0 of a buffer is its
// header (row count).  The receiver appends both received buffers behind its owned atoms.
__global__ void __launch_bounds__(256) k_slab_pack(const double *__restrict__ x, const double *__restrict__ y,
                                                   const double *__restrict__ z, const int *__restrict__ gid, int N,
                                                   DBox box, CellGrid g, int lo, int hi, int halo, double4 *__restrict__ left,
                                                   double4 *__restrict__ right, int cap, int *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double xr = x[i], yr = y[i], zr = z[i];
    double xi = xr, yi = yr, zi = zr;
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_of(box, g, xi, yi, zi, ic, jc, kc);
    const double4 rec = make_double4(xr, yr, zr, (double)gid[i]);   // raw coordinates travel (neighbor.cpp uses x[j] raw)
    if (ic >= lo && ic < lo + halo) {
        const int slot = atomicAdd(counts, 1);
        if (slot < cap - 1) left[slot + 1] = rec;
    }
    if (ic >= hi - halo && ic < hi) {
        const int slot = atomicAdd(counts + 1, 1);
        if (slot < cap - 1) right[slot + 1] = rec;
    }
}

__global__ void k_slab_headers(double4 *left, double4 *right, const int *counts)
{
    if (threadIdx.x == 0) left[0] = make_double4((double)counts[0], 0.0, 0.0, 0.0);
    if (threadIdx.x == 1) right[0] = make_double4((double)counts[1], 0.0, 0.0, 0.0);
}

__global__ void __launch_bounds__(256) k_slab_unpack(const double4 *__restrict__ a, const double4 *__restrict__ b, int cap,
                                                     double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                                                     int *__restrict__ gid, int n_owned, int room, int *__restrict__ total)
{
    const int ca = (int)a[0].x, cb = b ? (int)b[0].x : 0;
    const bool bad = ca > cap - 1 || cb > cap - 1 || ca + cb > room || ca < 0 || cb < 0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *total = bad ? -1 : n_owned + ca + cb;
    if (bad || i >= ca + cb) return;
    const double4 r = i < ca ? a[1 + i] : b[1 + i - ca];
    x[n_owned + i] = r.x;
    y[n_owned + i] = r.y;
    z[n_owned + i] = r.z;
    gid[n_owned + i] = (int)r.w;
}

__global__ void __launch_bounds__(256) k_translate_ids(const int *__restrict__ in, int *__restrict__ out, size_t n,
                                                       const int *__restrict__ gid)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int j = in[t];
        out[t] = j >= 0 ? gid[j] : j;
    }
}

}  // namespace

// wrap_positions, src/neighbor.cpp:675-702: every atom wrapped into the primary cell (box.h:131-176)
__global__ void __launch_bounds__(256) k_wrap_positions(double *__restrict__ x, double *__restrict__ y,
                                                        double *__restrict__ z, int N, DBox box)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[i], yi = y[i], zi = z[i];
    wrap_into_box(box, xi, yi, zi);
    x[i] = xi;
    y[i] = yi;
    z[i] = zi;
}

void launch_wrap_positions(MdbSystem &s, double *x, double *y, double *z, int N)
{
    MDB_LAUNCH(k_wrap_positions, (N + 255) / 256, 256, 0, s.stream, x, y, z, N, s.box);
    CUDA_TRY(cudaGetLastError());
}

// in -> out exclusive prefix sum over n ints on the system's stream (scratch: s.scan_tmp)
void device_exclusive_scan(MdbSystem &s, const int *in, int *out, int n)
{
    int *tmp = s.scan_tmp.ensure<int>(scan_tmp_ints(n));
    exclusive_scan(in, out, n, tmp, s.stream);
    CUDA_TRY(cudaGetLastError());
}

void launch_cell_planes(const double *x, const double *y, const double *z, int N, const DBox &b, const CellGrid &g,
                        int *plane, cudaStream_t st)
{
    if (N <= 0) return;
    MDB_LAUNCH(k_cell_planes, (N + 255) / 256, 256, 0, st, x, y, z, N, b, g, plane);
    CUDA_TRY(cudaGetLastError());
}

void launch_slab_pack(const double *x, const double *y, const double *z, const int *gid, int N, const DBox &b,
                      const CellGrid &g, int lo, int hi, int halo, double *left, double *right, int cap, int *counts,
                      cudaStream_t st)
{
    CUDA_TRY(cudaMemsetAsync(counts, 0, 2 * sizeof(int), st));
    if (N > 0)
        MDB_LAUNCH(k_slab_pack, (N + 255) / 256, 256, 0, st, x, y, z, gid, N, b, g, lo, hi, halo,
                   reinterpret_cast<double4 *>(left), reinterpret_cast<double4 *>(right), cap, counts);
    MDB_LAUNCH(k_slab_headers, 1, 32, 0, st, reinterpret_cast<double4 *>(left), reinterpret_cast<double4 *>(right), counts);
    CUDA_TRY(cudaGetLastError());
}

void launch_slab_unpack(const double *a, const double *b, int cap, double *x, double *y, double *z, int *gid, int n_owned,
                        int room, int *total, cudaStream_t st)
{
    const int maxrows = 2 * cap;
    MDB_LAUNCH(k_slab_unpack, (maxrows + 255) / 256, 256, 0, st, reinterpret_cast<const double4 *>(a),
               reinterpret_cast<const double4 *>(b), cap, x, y, z, gid, n_owned, room, total);
    CUDA_TRY(cudaGetLastError());
}

void launch_translate_ids(MdbSystem &s, const int *local_ids, int *global_ids, size_t n)
{
    if (!n) return;
    MDB_LAUNCH(k_translate_ids, 1184, 256, 0, s.stream, local_ids, global_ids, n, s.gid);
    CUDA_TRY(cudaGetLastError());
}

// Common tail of every binning: per-cell counts are in s.cell_count, cell ids in
// s.perm_tmp; produces s.cell_start, s.perm and the SortedAtom records gathered from X/Y/Z.
void finish_binning(MdbSystem &s, int nc, const double *X, const double *Y, const double *Z)
{
    const int N = s.N;
    cudaStream_t st = s.stream;
    int *count = s.cell_count.as<int>();
    int *cell_of_atom = s.perm_tmp.as<int>();
    int *start = s.cell_start.ensure<int>((size_t)nc + 1);
    int *perm = s.perm.ensure<int>(N);
    int *scan_tmp = s.scan_tmp.ensure<int>(scan_tmp_ints(nc + 1));
    int *counters = s.counters.ensure<int>(8);
    SortedAtom *sorted = s.sorted.ensure<SortedAtom>(N);
    const int nb = (N + 255) / 256;
    CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * 4, st));
    exclusive_scan(count, start, nc + 1, scan_tmp, st);
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)nc, st));
    MDB_LAUNCH(k_scatter, nb, 256, 0, st, cell_of_atom, N, start, count, perm);
    // worst case every cell is "big": N / SMALL_CELL entries
    int *big = s.big_cells.ensure<int>((size_t)N / SMALL_CELL + 2);
    MDB_LAUNCH(k_order_cells, (nc + 255) / 256, 256, 0, st, start, nc, perm, big, counters, s.gid);
    // over-full cells (rare: > SMALL_CELL atoms in one cell) are rank-sorted by whole blocks
    int *tmp = s.scratch.ensure<int>(N);
    MDB_LAUNCH(k_order_big, 296, 256, 0, st, start, big, counters, perm, tmp, s.gid);
    MDB_LAUNCH(k_copy_big, 296, 256, 0, st, start, big, counters, perm, tmp);
    MDB_LAUNCH(k_gather, nb, 256, 0, st, X, Y, Z, perm, cell_of_atom, N, sorted);
    CUDA_TRY(cudaGetLastError());
}

void launch_binning(MdbSystem &s, double rc)
{
    MDB_REQUIRE(s.N > 0 && s.x, MDB_ERR_STATE, "no atoms uploaded");
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g", rc);
    CellGrid g = cellgrid_make(s.box, rc);
    if (s.slab_nx > 0) {
        MDB_REQUIRE(s.slab_nx <= g.n[0] && s.slab_x0 >= 0 && s.slab_x0 < g.n[0], MDB_ERR_VALUE,
                    "slab window [%d,+%d) does not fit the %d x-planes of the cell grid", s.slab_x0, s.slab_nx,
                    g.n[0]);
        g.x0 = s.slab_x0;
        g.nxl = s.slab_nx;
    }
    g.total = g.nxl * g.n[1] * g.n[2];
    MDB_REQUIRE((double)g.nxl * g.n[1] * g.n[2] < 2.0e9, MDB_ERR_VALUE, "cell grid %dx%dx%d too large", g.nxl,
                g.n[1], g.n[2]);
    s.grid = g;
    const int N = s.N, nc = g.total;
    cudaStream_t st = s.stream;
    int *count = s.cell_count.ensure<int>((size_t)nc + 1);
    int *cell_of_atom = s.perm_tmp.ensure<int>(N);
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)nc + 1), st));
    int *counters = s.counters.ensure<int>(8);
    CUDA_TRY(cudaMemsetAsync(counters + 7, 0, sizeof(int), st));
    MDB_LAUNCH(k_cell_count, (N + 255) / 256, 256, 0, st, s.x, s.y, s.z, N, s.box, g, cell_of_atom, count,
               counters + 7);
    if (s.slab_nx > 0) {
        int bad = 0;
        CUDA_TRY(cudaMemcpyAsync(&bad, counters + 7, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        MDB_REQUIRE(bad == 0, MDB_ERR_VALUE, "%d atoms lie outside the slab window (owned + ghost planes)", bad);
    }
    finish_binning(s, nc, s.x, s.y, s.z);
    s.bin_rc = rc;
}
