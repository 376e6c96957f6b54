// mdapy_b200/csrc/cluster.cu
//
// Cluster analysis (connected components of the bond graph) on the device.  Replaces
// src/cluster.cpp:9-60 (get_cluster), 62-112 (get_cluster_by_bond) and 114-150 (filter_by_type).
//
// The reference floods breadth-first from every still-unlabelled seed in ascending atom index and hands
// out cluster ids 1, 2, ... in that order, so for a symmetric bond list (every cut-off list is one, and
// filter_by_type removes bonds symmetrically) the id of a cluster is the RANK of its smallest atom index
// among all clusters' smallest indices.  Here: lock-free union-find that always hooks the larger root
// under the smaller one (so the root IS the smallest index), a flatten pass, a prefix sum over the root
// flags, and id = rank + 1.  Deterministic and identical to the serial result.
#include "internal.cuh"

namespace {

__device__ __forceinline__ int uf_find(int *parent, int v)
{
    // L2 loads (__ldcg): the L1 of another SM is not coherent with this SM's view of freshly hooked roots
    int p = __ldcg(parent + v);
    while (p != v) {  // path halving; parents only ever decrease, so stale reads are still ancestors
        const int g = __ldcg(parent + p);
        if (g != p) parent[v] = g;
        v = p;
        p = g;
    }
    return v;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }  // a > b: hook a under b
        const int old = atomicCAS(parent + a, a, b);
        if (old == a) return;
    }
}

__global__ void __launch_bounds__(256) k_iota(int *__restrict__ p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// dist != nullptr: bond iff distance <= rc (get_cluster); else bond iff the entry is >= 0 (get_cluster_by_bond)
__global__ void __launch_bounds__(128) k_cluster_union(const int *__restrict__ verlet, const double *__restrict__ dist,
                                                       const int *__restrict__ nn, int N, int M, double rc,
                                                       int *__restrict__ parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = min(nn[i], M);
    for (int q = 0; q < c; ++q) {
        const int j = verlet[(size_t)i * M + q];
        const bool bond = dist ? (dist[(size_t)i * M + q] <= rc) : (j > -1);
        if (bond && j >= 0 && j < N && j != i) uf_union(parent, i, j);
    }
}

__global__ void __launch_bounds__(256) k_cluster_flatten(int *__restrict__ parent, int N, int *__restrict__ is_root)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int r = uf_find(parent, i);
    parent[i] = r;
    is_root[i] = r == i ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_cluster_assign(const int *__restrict__ parent, const int *__restrict__ rank,
                                                        int N, int *__restrict__ cluster)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) cluster[i] = rank[parent[i]] + 1;
}

// cluster.cpp:114-150: entry (i, jj) becomes -1 when some k has type1[k] == type[i], type2[k] == type[j] and
// distance > r[k]
__global__ void __launch_bounds__(128) k_filter_by_type(int *__restrict__ verlet, const double *__restrict__ dist,
                                                        const int *__restrict__ nn, int N, int M,
                                                        const int *__restrict__ types, const int *__restrict__ t1,
                                                        const int *__restrict__ t2, const double *__restrict__ r,
                                                        int npair)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = min(nn[i], M);
    const int ti = types[i];
    for (int q = 0; q < c; ++q) {
        const int j = verlet[(size_t)i * M + q];
        if (j < 0) continue;
        const int tj = types[j];
        const double d = dist[(size_t)i * M + q];
        bool cut = false;
        for (int k = 0; k < npair; ++k) cut |= (t1[k] == ti) & (t2[k] == tj) & (d > r[k]);
        if (cut) verlet[(size_t)i * M + q] = -1;
    }
}

}  // namespace

void launch_filter_by_type(MdbSystem &s, int *verlet, const double *dist, const int *nn, int M, const int *types,
                           const int *t1, const int *t2, const double *r, int npair)
{
    const int N = s.n_rows;
    MDB_LAUNCH(k_filter_by_type, (N + 127) / 128, 128, 0, s.stream, verlet, dist, nn, N, M, types, t1, t2, r, npair);
    CUDA_TRY(cudaGetLastError());
}

// cluster (device, N ints) <- ids 1..count; returns count.  dist == nullptr selects the by-bond rule.
int launch_cluster(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, int *cluster)
{
    const int N = s.n_rows;
    cudaStream_t st = s.stream;
    int *parent = s.scratch.ensure<int>((size_t)2 * N + 2);
    int *flag = parent + N;  // N + 1 entries: the scan of [flags | 0] puts the total at index N
    int *rank = s.perm_tmp.ensure<int>((size_t)N + 1);
    const int nb = (N + 255) / 256;
    MDB_LAUNCH(k_iota, nb, 256, 0, st, parent, N);
    MDB_LAUNCH(k_cluster_union, (N + 127) / 128, 128, 0, st, verlet, dist, nn, N, M, rc, parent);
    CUDA_TRY(cudaMemsetAsync(flag + N, 0, sizeof(int), st));
    MDB_LAUNCH(k_cluster_flatten, nb, 256, 0, st, parent, N, flag);
    device_exclusive_scan(s, flag, rank, N + 1);
    MDB_LAUNCH(k_cluster_assign, nb, 256, 0, st, parent, rank, N, cluster);
    int count = 0;
    CUDA_TRY(cudaMemcpyAsync(&count, rank + N, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return count;
}
