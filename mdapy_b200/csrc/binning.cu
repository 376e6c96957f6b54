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
                                                    int *__restrict__ cell_of_atom, int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[i], yi = y[i], zi = z[i];
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_of(box, g, xi, yi, zi, ic, jc, kc);
    const int c = (ic * g.n[1] + jc) * g.n[2] + kc;
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
__global__ void __launch_bounds__(256) k_order_cells(const int *__restrict__ cell_start, int ncell, int *__restrict__ perm,
                                                     int *__restrict__ big_cells, int *__restrict__ n_big)
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
        int q = a - 1;
        while (q >= b && perm[q] > v) {
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
                                                   int *__restrict__ perm_out)
{
    for (int w = blockIdx.x; w < *n_big; w += gridDim.x) {
        const int c = big_cells[w];
        const int b = cell_start[c], e = cell_start[c + 1];
        for (int a = b + threadIdx.x; a < e; a += blockDim.x) {
            const int v = perm[a];
            int rank = 0;
            for (int q = b; q < e; ++q) rank += (perm[q] < v);
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

}  // namespace

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
    CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * 8, st));
    exclusive_scan(count, start, nc + 1, scan_tmp, st);
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)nc, st));
    MDB_LAUNCH(k_scatter, nb, 256, 0, st, cell_of_atom, N, start, count, perm);
    // worst case every cell is "big": N / SMALL_CELL entries
    int *big = s.big_cells.ensure<int>((size_t)N / SMALL_CELL + 2);
    MDB_LAUNCH(k_order_cells, (nc + 255) / 256, 256, 0, st, start, nc, perm, big, counters);
    // over-full cells (rare: > SMALL_CELL atoms in one cell) are rank-sorted by whole blocks
    int *tmp = s.scratch.ensure<int>(N);
    MDB_LAUNCH(k_order_big, 296, 256, 0, st, start, big, counters, perm, tmp);
    MDB_LAUNCH(k_copy_big, 296, 256, 0, st, start, big, counters, perm, tmp);
    MDB_LAUNCH(k_gather, nb, 256, 0, st, X, Y, Z, perm, cell_of_atom, N, sorted);
    CUDA_TRY(cudaGetLastError());
}

void launch_binning(MdbSystem &s, double rc)
{
    MDB_REQUIRE(s.N > 0 && s.x, MDB_ERR_STATE, "no atoms uploaded");
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g", rc);
    const CellGrid g = cellgrid_make(s.box, rc);
    MDB_REQUIRE((double)g.n[0] * g.n[1] * g.n[2] < 2.0e9, MDB_ERR_VALUE, "cell grid %dx%dx%d too large", g.n[0],
                g.n[1], g.n[2]);
    s.grid = g;
    const int N = s.N, nc = g.total;
    cudaStream_t st = s.stream;
    int *count = s.cell_count.ensure<int>((size_t)nc + 1);
    int *cell_of_atom = s.perm_tmp.ensure<int>(N);
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)nc + 1), st));
    MDB_LAUNCH(k_cell_count, (N + 255) / 256, 256, 0, st, s.x, s.y, s.z, N, s.box, g, cell_of_atom, count);
    finish_binning(s, nc, s.x, s.y, s.z);
    s.bin_rc = rc;
}
