// mdapy_b200/csrc/planar_faults.cu
//
// FCC planar-fault identification on the PTM result: src/identify_fcc_planar_faults.cpp:43-241.
// Per HCP-like atom: 0 non-HCP, 1 other / indeterminate, 2 intrinsic stacking fault, 3 coherent twin boundary,
// 4 multi-layer stacking fault, 5 extrinsic stacking fault.
//
// The algorithm only needs, for every HCP atom, WHICH of its 12 matched neighbours lie in its basal plane, above
// and below it -- i.e. the z layer of the HCP template point every neighbour was mapped to.  That is a table over
// the template's point order; `order` selects the table: 0 = this library's template (ptm_tables.h: six in-plane
// points, three below, three above), 1 = the reference's (identify_fcc_planar_faults.cpp:63-65).  Everything
// else (shared-basal-neighbour test, counting) is order-free, and the result is invariant under the HCP
// template's symmetry operations, so either PTM implementation feeds it.
//
// The third step of the reference is a SERIAL sweep in ascending atom order whose in-place updates are read by
// later atoms (lines 147-189); it is reproduced by one thread (the HCP atoms of a frame are its defects: few).
#include "internal.cuh"

namespace {

struct PftTables {
    int layer_dir[12];
    int basal[6];
    int outofplane[6];
};

__host__ PftTables pft_tables(int order)
{
    PftTables t;
    if (order == 1) {
        const int ld[12] = {0, 0, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1};
        const int ba[6] = {0, 1, 5, 6, 7, 8}, op[6] = {2, 3, 4, 9, 10, 11};
        for (int i = 0; i < 12; ++i) t.layer_dir[i] = ld[i];
        for (int i = 0; i < 6; ++i) t.basal[i] = ba[i], t.outofplane[i] = op[i];
    } else {
        const int ld[12] = {0, 0, 0, 0, 0, 0, -1, -1, -1, 1, 1, 1};
        for (int i = 0; i < 12; ++i) t.layer_dir[i] = ld[i];
        for (int i = 0; i < 6; ++i) t.basal[i] = i, t.outofplane[i] = 6 + i;
    }
    return t;
}

__global__ void __launch_bounds__(256) k_pft_flag(const int *__restrict__ type, int N, int *__restrict__ flag,
                                                  int *__restrict__ fault)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    if (i == N) {
        flag[N] = 0;
        return;
    }
    flag[i] = type[i] == 2;
    fault[i] = 0;
}

__global__ void __launch_bounds__(256) k_pft_list(const int *__restrict__ flag, const int *__restrict__ rank, int N,
                                                  int *__restrict__ hcp_idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && flag[i]) hcp_idx[rank[i]] = i;
}

// hcp_neigh[r][j]: rank of the j-th matched neighbour among the HCP atoms, or -(atom)-1 when it is not HCP
__global__ void __launch_bounds__(256) k_pft_neigh(const int *__restrict__ hcp_idx, int n_hcp, const int *__restrict__ ptm_idx,
                                                   int stride, int col0, const int *__restrict__ type,
                                                   const int *__restrict__ rank, int *__restrict__ hcp_neigh)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_hcp * 12) return;
    const int r = t / 12, j = t - r * 12;
    const int b = ptm_idx[(size_t)hcp_idx[r] * stride + col0 + j];
    hcp_neigh[t] = (b >= 0 && type[b] == 2) ? rank[b] : -b - 1;
}

__device__ __forceinline__ bool are_stacked(const int *__restrict__ hn, int a, int b, const PftTables &T)
{
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j)
            if (hn[a * 12 + T.basal[i]] == hn[b * 12 + T.basal[j]]) return false;
    return true;
}

__global__ void __launch_bounds__(128) k_pft_classify(const int *__restrict__ hcp_idx, int n_hcp, const int *__restrict__ hn,
                                                      const int *__restrict__ type, PftTables T, int *__restrict__ fault)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hcp) return;
    int n_basal = 0, n_pos = 0, n_neg = 0, n_fcc_pos = 0, n_fcc_neg = 0;
    for (int j = 0; j < 12; ++j) {
        const int nb = hn[i * 12 + j];
        if (nb >= 0) {
            if (T.layer_dir[j] == 0) ++n_basal;
            else if (are_stacked(hn, i, nb, T)) {
                if (T.layer_dir[j] == 1) ++n_pos;
                else ++n_neg;
            }
        } else if (T.layer_dir[j] != 0) {
            if (type[-nb - 1] == 1) {
                if (T.layer_dir[j] > 0) ++n_fcc_pos;
                else ++n_fcc_neg;
            }
        }
    }
    int f;
    if ((n_pos != 0 && n_neg == 0) || (n_pos == 0 && n_neg != 0)) f = 2;
    else if (n_basal >= 1 && n_pos == 0 && n_neg == 0 && n_fcc_pos != 0 && n_fcc_neg != 0) f = 3;
    else if (n_pos != 0 && n_neg != 0) f = 4;
    else f = 1;
    fault[hcp_idx[i]] = f;
}

// the reference's serial third step, lines 147-189 (in-place updates seen by later atoms)
__global__ void k_pft_sweep(const int *__restrict__ hcp_idx, int n_hcp, const int *__restrict__ hn, PftTables T,
                            int *__restrict__ fault)
{
    if (blockIdx.x || threadIdx.x) return;
    for (int i = 0; i < n_hcp; ++i) {
        const int a = hcp_idx[i];
        const int f = fault[a];
        if (f == 3 || f == 1) {
            int n_isf = 0, n_twin = 0;
            for (int jj = 0; jj < 6; ++jj) {
                const int nb = hn[i * 12 + T.basal[jj]];
                if (nb >= 0) {
                    const int nf = fault[hcp_idx[nb]];
                    if (nf == 2) ++n_isf;
                    else if (nf == 3) ++n_twin;
                }
            }
            if (n_isf != 0 && n_twin == 0) fault[a] = 2;
            else if (n_isf == 0 && n_twin != 0) fault[a] = 3;
        } else if (f == 4) {
            for (int jj = 0; jj < 6; ++jj) {
                const int nb = hn[i * 12 + T.outofplane[jj]];
                if (nb >= 0) {
                    const int na = hcp_idx[nb];
                    if (fault[na] == 2) fault[na] = 4;
                }
            }
        }
    }
}

// extrinsic stacking faults, lines 196-240: a twin-boundary atom with an FCC neighbour that itself has 5-6 FCC and
// 5-6 HCP neighbours
__global__ void __launch_bounds__(128) k_pft_esf(const int *__restrict__ hcp_idx, int n_hcp, const int *__restrict__ ptm_idx,
                                                 int stride, int col0, const int *__restrict__ type, int *__restrict__ fault)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hcp) return;
    const int a = hcp_idx[i];
    if (fault[a] != 3) return;
    const int *row = ptm_idx + (size_t)a * stride + col0;
    for (int j = 0; j < 12; ++j) {
        const int b = row[j];
        if (b < 0 || type[b] != 1) continue;
        int nf = 0, nh = 0;
        const int *rowj = ptm_idx + (size_t)b * stride + col0;
        for (int k = 0; k < 12; ++k) {
            const int c = rowj[k];
            if (c < 0) continue;
            const int t = type[c];
            nf += t == 1;
            nh += t == 2;
        }
        if (nf >= 5 && nf <= 6 && nh >= 5 && nh <= 6) {
            fault[a] = 5;
            return;
        }
    }
}

__global__ void __launch_bounds__(256) k_pft_types_from_output(const double *__restrict__ out, int ocols, int N,
                                                               int *__restrict__ type)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) type[i] = (int)out[(size_t)i * ocols];
}

}  // namespace

// type: structure types [N] (device); ptm_idx: [N][stride] with the 12 matched neighbours at columns col0..col0+11
void launch_planar_faults(MdbSystem &s, const int *type, int N, const int *ptm_idx, int stride, int col0, int order,
                          bool identify_esf, int *fault)
{
    if (N <= 0) return;
    const PftTables T = pft_tables(order);
    int *flag = s.scratch.ensure<int>((size_t)N + 1);
    int *rank = s.scratch2.ensure<int>((size_t)N + 1);
    MDB_LAUNCH(k_pft_flag, (N + 256) / 256, 256, 0, s.stream, type, N, flag, fault);
    device_exclusive_scan(s, flag, rank, N + 1);
    int n_hcp = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_hcp, rank + N, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (n_hcp <= 0) return;
    int *hcp_idx = s.verlet_tmp.ensure<int>((size_t)n_hcp * 13);
    int *hn = hcp_idx + n_hcp;
    MDB_LAUNCH(k_pft_list, (N + 255) / 256, 256, 0, s.stream, flag, rank, N, hcp_idx);
    MDB_LAUNCH(k_pft_neigh, (n_hcp * 12 + 255) / 256, 256, 0, s.stream, hcp_idx, n_hcp, ptm_idx, stride, col0, type, rank, hn);
    MDB_LAUNCH(k_pft_classify, (n_hcp + 127) / 128, 128, 0, s.stream, hcp_idx, n_hcp, hn, type, T, fault);
    MDB_LAUNCH(k_pft_sweep, 1, 32, 0, s.stream, hcp_idx, n_hcp, hn, T, fault);
    if (identify_esf) MDB_LAUNCH(k_pft_esf, (n_hcp + 127) / 128, 128, 0, s.stream, hcp_idx, n_hcp, ptm_idx, stride, col0, type, fault);
    CUDA_TRY(cudaGetLastError());
}

void launch_types_from_ptm_output(MdbSystem &s, const double *out, int ocols, int N, int *type)
{
    MDB_LAUNCH(k_pft_types_from_output, (N + 255) / 256, 256, 0, s.stream, out, ocols, N, type);
    CUDA_TRY(cudaGetLastError());
}
