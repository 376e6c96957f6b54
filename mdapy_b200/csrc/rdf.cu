// mdapy_b200/csrc/rdf.cu
//
// Pair-distance histograms.  Replaces src/radial_distribution_function.cpp:22-54 (_rdf),
// 56-85 (_rdf_single_species) and 143-317 (_rdf_streaming).  Counts are integers, so the result
// does not depend on summation order; they are accumulated as 64-bit integers (per-block shared
// histograms, then global atomics) and ADDED to the caller's f64 histogram at the end, like the
// reference's `+=`.
//
// List kernels use the stored distances: `dis < rc` strict, bin (int)(dis/dr), dr = rc/nbin.
// The reference has no `k < nbin` guard there (Appendix D.5, an out-of-bounds write when dis/dr
// rounds up to nbin); such a pair is dropped here.
// The streaming kernel walks the cut-off cell grid straight from positions: xi wrapped, x[j] raw,
// min-image, r2 < rc^2 strict, k = (int)(sqrt(r2)/dr), k < nbin (cpp:214-256).  Pair membership
// does not depend on the cell decomposition, so the reference's "cell list vs all pairs" switch
// (cpp:167-176) needs no counterpart.
#include "internal.cuh"

namespace {

constexpr int RDF_SMEM_BINS = 8192;  // 64 KB of 64-bit... kept as 32-bit counters in shared memory (32 KB)

__device__ __forceinline__ void hist_add(unsigned *sh, unsigned long long *gl, bool use_sh, int slot, unsigned v)
{
    if (use_sh)
        atomicAdd(&sh[slot], v);
    else
        atomicAdd(&gl[slot], (unsigned long long)v);
}

// floor(d / dr) as the reference computes it (the IEEE quotient, truncated), without the division sequence for
// almost every entry: k0 = trunc(d * (1/dr)) can only differ from it when d / dr lies within a few ulps of an
// integer; those entries (fractional part of the product within 1e-9 of 0 or 1) take the real division.
__device__ __forceinline__ int rdf_bin(double d, double dr, double inv_dr)
{
    const double t = d * inv_dr;
    const double f = t - floor(t);
    if (f < 1e-9 || f > 1.0 - 1e-9) return (int)(d / dr);
    return (int)t;
}

__device__ __forceinline__ void hist_flush(unsigned *sh, unsigned long long *gl, int nslot)
{
    __syncthreads();
    for (int t = threadIdx.x; t < nslot; t += blockDim.x) {
        const unsigned v = sh[t];
        if (v) atomicAdd(&gl[t], (unsigned long long)v);
    }
}

// The same counts with one thread per LIST ENTRY (flat index over the N x M arrays): the distance and index
// rows are read once, fully coalesced, and the 32 entries of a warp belong to one or two atoms, so their bins
// differ (thread-per-row put the j-th neighbours of 32 atoms -- in a crystal: the same shell, the same bin -- into
// one atomic instruction).
__global__ void __launch_bounds__(256) k_rdf_list_flat(const int *__restrict__ verlet, const double *__restrict__ dist,
                                                       const int *__restrict__ nn, int N, int M,
                                                       const int *__restrict__ types, int ntype, double rc, int nbin,
                                                       const int *__restrict__ gid, unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned sh[];
    const int nslot = (types ? ntype * ntype : 1) * nbin;
    const bool use_sh = nslot <= RDF_SMEM_BINS;
    if (use_sh) {
        for (int t = threadIdx.x; t < nslot; t += blockDim.x) sh[t] = 0;
        __syncthreads();
    }
    const double dr = rc / nbin, inv_dr = 1.0 / dr, inv_M = 1.0 / M;
    const size_t total = (size_t)N * M, step = (size_t)gridDim.x * blockDim.x;
    auto one = [&](size_t e, double d, int j) {
        if (!(d < rc)) return;
        // row of entry e without a 64-bit integer division: the double quotient is off by at most one
        int i = (int)((double)e * inv_M);
        if ((size_t)(i + 1) * M <= e) ++i;
        else if ((size_t)i * M > e) --i;
        const int q = (int)(e - (size_t)i * M);
        if (q >= nn[i]) return;
        const int k = rdf_bin(d, dr, inv_dr);
        if (k >= nbin || k < 0) return;
        if (types)
            hist_add(sh, hist, use_sh, (types[i] * ntype + types[j]) * nbin + k, 1u);
        else if (gid ? gid[j] > gid[i] : j > i)
            hist_add(sh, hist, use_sh, k, 2u);
    };
    // four entries per thread and trip, all eight loads issued before the first is used (eight per trip is slower:
    // the registers cost more residency than the extra loads in flight gain)
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; e + 3 * step < total; e += 4 * step) {
        const double d0 = dist[e], d1 = dist[e + step], d2 = dist[e + 2 * step], d3 = dist[e + 3 * step];
        const int j0 = verlet[e], j1 = verlet[e + step], j2 = verlet[e + 2 * step], j3 = verlet[e + 3 * step];
        one(e, d0, j0);
        one(e + step, d1, j1);
        one(e + 2 * step, d2, j2);
        one(e + 3 * step, d3, j3);
    }
    for (; e < total; e += step) one(e, dist[e], verlet[e]);
    if (use_sh) hist_flush(sh, hist, nslot);
}

// _rdf (typed) and _rdf_single_species (types == nullptr)
__global__ void __launch_bounds__(256) k_rdf_list(const int *__restrict__ verlet, const double *__restrict__ dist,
                                                  const int *__restrict__ nn, int N, int M,
                                                  const int *__restrict__ types, int ntype, double rc, int nbin,
                                                  const int *__restrict__ gid, unsigned long long *__restrict__ hist)
{
    // gid: global ids of a decomposed frame -- the single-species "j > i counts twice" rule compares
    // ORIGINAL indices, so every unordered pair is counted by exactly one rank
    extern __shared__ unsigned sh[];
    const int nslot = (types ? ntype * ntype : 1) * nbin;
    const bool use_sh = nslot <= RDF_SMEM_BINS;
    if (use_sh) {
        for (int t = threadIdx.x; t < nslot; t += blockDim.x) sh[t] = 0;
        __syncthreads();
    }
    const double dr = rc / nbin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int cnt = min(nn[i], M);
        const int it = types ? types[i] : 0;
        for (int q = 0; q < cnt; ++q) {
            const double d = dist[(size_t)i * M + q];
            if (!(d < rc)) continue;
            const int j = verlet[(size_t)i * M + q];
            const int k = (int)(d / dr);
            if (k >= nbin || k < 0) continue;
            if (types)
                hist_add(sh, hist, use_sh, (it * ntype + types[j]) * nbin + k, 1u);
            else if (gid ? gid[j] > gid[i] : j > i)
                hist_add(sh, hist, use_sh, k, 2u);
        }
    }
    if (use_sh) hist_flush(sh, hist, nslot);
}

__device__ __forceinline__ SortedAtom load_sorted2(const SortedAtom *__restrict__ p)
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

__global__ void __launch_bounds__(128) k_rdf_stream(const SortedAtom *__restrict__ sorted,
                                                    const int *__restrict__ cell_start, int N, int n_rows, DBox box,
                                                    CellGrid g,
                                                    const int *__restrict__ types, int ntype, double rc, int nbin,
                                                    unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned sh[];
    const int nslot = ntype * ntype * nbin;
    const bool use_sh = nslot <= RDF_SMEM_BINS;
    if (use_sh) {
        for (int t = threadIdx.x; t < nslot; t += blockDim.x) sh[t] = 0;
        __syncthreads();
    }
    const double dr = rc / nbin, rcsq = rc * rc;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const SortedAtom me = s < N ? load_sorted2(sorted + s) : SortedAtom{};
    if (s < N && me.idx < n_rows) {  // ghosts of a decomposed frame are neighbours only
        double xi = me.x, yi = me.y, zi = me.z;
        if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
        int ic, jc, kc;
        cell_decode(g, me.cell, ic, jc, kc);
        const int it = types[me.idx];
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
                for (int dk = -1; dk <= 1; ++dk) {
                    const int c = cell_linear(g, wrap_cell(ic + di, g.n[0]), wrap_cell(jc + dj, g.n[1]),
                                              wrap_cell(kc + dk, g.n[2]));
                    if (c < 0) continue;  // outside the stored slab window (never for an owned atom)
                    const int b = __ldg(cell_start + c), e = __ldg(cell_start + c + 1);
                    for (int q = b; q < e; ++q) {
                        if (q == s) continue;
                        const SortedAtom o = load_sorted2(sorted + q);
                        double dx = o.x - xi, dy = o.y - yi, dz = o.z - zi;
                        min_image(box, dx, dy, dz);
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        if (r2 < rcsq) {
                            const int k = (int)(sqrt(r2) / dr);
                            if (k < nbin) hist_add(sh, hist, use_sh, (it * ntype + types[o.idx]) * nbin + k, 1u);
                        }
                    }
                }
    }
    if (use_sh) hist_flush(sh, hist, nslot);
}

__global__ void __launch_bounds__(256) k_hist_accumulate(const unsigned long long *__restrict__ hist, int n,
                                                         double *__restrict__ g)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) g[t] += (double)hist[t];
}

}  // namespace

// g (device, f64) += counts.  types == nullptr selects the single-species kernel.
void launch_rdf_list(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int N, int M,
                     const int *types, int ntype, double rc, int nbin, double *g)
{
    MDB_REQUIRE(nbin > 0 && rc > 0, MDB_ERR_VALUE, "nbin and rc must be positive");
    const int nslot = (types ? ntype * ntype : 1) * nbin;
    unsigned long long *hist = s.scratch2.ensure<unsigned long long>(nslot);
    cudaStream_t st = s.stream;
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * nslot, st));
    const size_t smem = nslot <= RDF_SMEM_BINS ? sizeof(unsigned) * nslot : 0;
    int nb = (N + 255) / 256;
    if (nb > 1184) nb = 1184;
    const char *mode = getenv("MDB_RDF");
    if (mode && !strcmp(mode, "rows")) {
        MDB_LAUNCH(k_rdf_list, nb, 256, smem, st, verlet, dist, nn, N, M, types, ntype, rc, nbin, s.gid, hist);
    } else {
        const size_t total = (size_t)N * M;
        size_t nbf = (total + 255) / 256;
        if (nbf > 148 * 16) nbf = 148 * 16;
        MDB_LAUNCH(k_rdf_list_flat, (int)nbf, 256, smem, st, verlet, dist, nn, N, M, types, ntype, rc, nbin, s.gid, hist);
    }
    MDB_LAUNCH(k_hist_accumulate, (nslot + 255) / 256, 256, 0, st, hist, nslot, g);
    CUDA_TRY(cudaGetLastError());
}

void launch_rdf_streaming(MdbSystem &s, const int *types, int ntype, double rc, int nbin, double *g)
{
    MDB_REQUIRE(nbin > 0 && rc > 0, MDB_ERR_VALUE, "nbin and rc must be positive");
    if (s.bin_rc != rc) launch_binning(s, rc);
    const int nslot = ntype * ntype * nbin;
    unsigned long long *hist = s.scratch2.ensure<unsigned long long>(nslot);
    cudaStream_t st = s.stream;
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * nslot, st));
    const size_t smem = nslot <= RDF_SMEM_BINS ? sizeof(unsigned) * nslot : 0;
    MDB_LAUNCH(k_rdf_stream, (s.N + 127) / 128, 128, smem, st, s.sorted.as<SortedAtom>(), s.cell_start.as<int>(), s.N,
               s.n_rows, s.box, s.grid, types, ntype, rc, nbin, hist);
    MDB_LAUNCH(k_hist_accumulate, (nslot + 255) / 256, 256, 0, st, hist, nslot, g);
    CUDA_TRY(cudaGetLastError());
}
