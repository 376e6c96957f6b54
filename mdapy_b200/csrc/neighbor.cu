// mdapy_b200/csrc/neighbor.cu
//
// Fixed-radius neighbour build on the cell-sorted copy.  Replaces
// src/neighbor.cpp:102-187 (build_verlet_list) / 189-349 (dynamic sizer).
//
// Row contract kept from the reference: row i (ORIGINAL atom
// index) lists original indices j with d2 <= rc*rc (inclusive), distance
// sqrt(d2); slots beyond M are never written but always counted.  Row ORDER
// is also the reference's: the 27 cells in (x,y,z)-nested stencil order and
// descending atom index inside a cell (head-insertion chains), which makes
// every downstream tie-break (selection sort, first-k) bit-identical.
//
// d2 arithmetic (Appendix A of SURVEY.md): xi wrapped, x[j] raw,
// xij = x[j] - xi, min-image, xij*xij + yij*yij + zij*zij evaluated left to
// right with no FMA contraction.
#include "internal.cuh"

namespace {

__device__ __forceinline__ SortedAtom load_sorted(const SortedAtom *__restrict__ p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 lo = __ldg(q), hi = __ldg(q + 1);
    SortedAtom a;
    a.x = lo.x;
    a.y = lo.y;
    a.z = hi.x;
    a.idx = __double2loint(hi.y);
    a.cell = __double2hiint(hi.y);
    return a;
}

// One thread per atom, walking the cell-sorted records through L1/L2.
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(128) k_neighbor_direct(const SortedAtom *__restrict__ sorted,
                                                         const int *__restrict__ cell_start, int N, DBox box,
                                                         CellGrid g, double rcsq, int M, int n_rows,
                                                         int *__restrict__ verlet, double *__restrict__ dist,
                                                         int *__restrict__ nn)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const SortedAtom me = load_sorted(sorted + s);
    if (me.idx >= n_rows) return;  // ghost atom of a decomposed frame: neighbour only
    double xi = me.x, yi = me.y, zi = me.z;
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_decode(g, me.cell, ic, jc, kc);
    int *vrow = verlet + (size_t)me.idx * M;
    double *drow = dist + (size_t)me.idx * M;
    int cnt = 0;
    for (int di = -1; di <= 1; ++di) {
        const int ci = wrap_cell(ic + di, g.n[0]);
        for (int dj = -1; dj <= 1; ++dj) {
            const int cj = wrap_cell(jc + dj, g.n[1]);
            for (int dk = -1; dk <= 1; ++dk) {
                const int ck = wrap_cell(kc + dk, g.n[2]);
                const int c = cell_linear(g, ci, cj, ck);
                if (c < 0) continue;  // cannot happen for an owned atom with both ghost planes present
                const int b = __ldg(cell_start + c), e = __ldg(cell_start + c + 1);
                for (int q = e - 1; q >= b; --q) {  // descending original index
                    if (q == s) continue;
                    const SortedAtom o = load_sorted(sorted + q);
                    double dx = o.x - xi, dy = o.y - yi, dz = o.z - zi;
                    min_image(box, dx, dy, dz);
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 <= rcsq) {
                        if (!COUNT_ONLY && cnt < M) {
                            vrow[cnt] = o.idx;
                            drow[cnt] = sqrt(d2);
                        }
                        ++cnt;
                    }
                }
            }
        }
    }
    nn[me.idx] = cnt;
}

__global__ void __launch_bounds__(256) k_fill_rows(int *__restrict__ verlet, double *__restrict__ dist, size_t total,
                                                   double pad)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        verlet[t] = -1;
        dist[t] = pad;
    }
}

// copy rows from stride M_from to stride M_to (M_to <= M_from)
__global__ void __launch_bounds__(256) k_compact_rows(const int *__restrict__ vin, const double *__restrict__ din,
                                                      int *__restrict__ vout, double *__restrict__ dout, int N,
                                                      int M_from, int M_to)
{
    const size_t total = (size_t)N * M_to;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (total < 0xffffffffull) {  // 32-bit index arithmetic (the common case)
        const unsigned mt = (unsigned)M_to;
        for (; t < total; t += stride) {
            const unsigned i = (unsigned)t / mt, k = (unsigned)t - i * mt;
            vout[t] = vin[(size_t)i * M_from + k];
            dout[t] = din[(size_t)i * M_from + k];
        }
        return;
    }
    for (; t < total; t += stride) {
        const size_t i = t / M_to, k = t % M_to;
        vout[t] = vin[i * M_from + k];
        dout[t] = din[i * M_from + k];
    }
}

__global__ void __launch_bounds__(256) k_reduce_minmax(const int *__restrict__ v, size_t n, int *__restrict__ out)
{
    int mx = INT_MIN, mn = INT_MAX;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int a = v[t];
        mx = max(mx, a);
        mn = min(mn, a);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mx);
        atomicMin(out + 1, mn);
    }
}

}  // namespace

static void minmax_int(MdbSystem &s, const int *v, size_t n, int &mx, int &mn)
{
    int *c = s.counters.ensure<int>(8);
    const int init[2] = {INT_MIN, INT_MAX};
    CUDA_TRY(cudaMemcpyAsync(c + 4, init, sizeof(init), cudaMemcpyHostToDevice, s.stream));
    const int nb = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    if (n) MDB_LAUNCH(k_reduce_minmax, nb, 256, 0, s.stream, v, n, c + 4);
    int res[2];
    CUDA_TRY(cudaMemcpyAsync(res, c + 4, sizeof(res), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    mx = res[0];
    mn = res[1];
}

int device_max_int(MdbSystem &s, const int *v, size_t n)
{
    int mx, mn;
    minmax_int(s, v, n, mx, mn);
    return n ? mx : 0;
}

int device_min_int(MdbSystem &s, const int *v, size_t n)
{
    int mx, mn;
    minmax_int(s, v, n, mx, mn);
    return n ? mn : 0;
}

// Writes s.verlet / s.dist / s.nn with row stride M (prefilled -1 / rc+1 as
// neighbor.py:125-129 and neighbor.cpp:320-328 do).
void launch_neighbor(MdbSystem &s, double rc, int M, bool count_only)
{
    MDB_REQUIRE(s.bin_rc == rc, MDB_ERR_STATE, "binning for rc=%g missing", rc);
    const int N = s.N, R = s.n_rows;
    cudaStream_t st = s.stream;
    int *nn = s.nn.ensure<int>(R);
    const double rcsq = rc * rc;
    const int nb = (N + 127) / 128;
    if (count_only) {
        MDB_LAUNCH(k_neighbor_direct<true>, nb, 128, 0, st, s.sorted.as<SortedAtom>(), s.cell_start.as<int>(), N, s.box, s.grid,
                                                    rcsq, 0, s.n_rows, nullptr, nullptr, nn);
    } else {
        MDB_REQUIRE(M > 0, MDB_ERR_VALUE, "max_neigh must be positive, got %d", M);
        int *verlet = s.verlet.ensure<int>((size_t)R * M);
        double *dist = s.dist.ensure<double>((size_t)R * M);
        MDB_LAUNCH(k_fill_rows, 1184, 256, 0, st, verlet, dist, (size_t)R * M, rc + 1.0);
        MDB_LAUNCH(k_neighbor_direct<false>, nb, 128, 0, st, s.sorted.as<SortedAtom>(), s.cell_start.as<int>(), N, s.box,
                                                     s.grid, rcsq, M, s.n_rows, verlet, dist, nn);
    }
    CUDA_TRY(cudaGetLastError());
}

void launch_compact_rows(MdbSystem &s, int M_from, int M_to)
{
    const int N = s.n_rows;
    int *vout = s.verlet_tmp.ensure<int>((size_t)N * M_to);
    double *dout = s.dist_tmp.ensure<double>((size_t)N * M_to);
    MDB_LAUNCH(k_compact_rows, 1184, 256, 0, s.stream, s.verlet.as<int>(), s.dist.as<double>(), vout, dout, N, M_from, M_to);
    CUDA_TRY(cudaGetLastError());
    std::swap(s.verlet, s.verlet_tmp);
    std::swap(s.dist, s.dist_tmp);
}
