// mdapy_b200/csrc/descriptors.cu
//
// Per-atom descriptors that consume a sorted neighbour list:
//   k_sort_rows  partial selection sort of the first k slots
//                (src/neighbor.cpp:745-778 sort_verlet_by_distance)
//   k_csp        centro-symmetry parameter (src/centro_symmetry_parameter.cpp:12-92)
//   k_aja        Ackland-Jones analysis    (src/ackland_jones_analysis.cpp:9-172)
// All arithmetic follows the reference operation order.
#include "internal.cuh"

namespace {

// The reference's selection sort is not stable (it swaps), so the exact swap
// sequence is replayed: first minimum wins (strict <), scan covers the whole
// row including the rc+1 padding.
__global__ void __launch_bounds__(128) k_sort_rows(int *__restrict__ verlet, double *__restrict__ dist, int N, int M,
                                                   int k)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int *v = verlet + (size_t)i * M;
    double *d = dist + (size_t)i * M;
    const int eff = k < M ? k : M;
    for (int j = 0; j < eff; ++j) {
        int mi = j;
        double md = d[j];
        for (int q = j + 1; q < M; ++q) {
            const double dq = d[q];
            if (dq < md) {
                md = dq;
                mi = q;
            }
        }
        if (mi != j) {
            const double td = d[j];
            d[j] = md;
            d[mi] = td;
            const int tv = v[j];
            v[j] = v[mi];
            v[mi] = tv;
        }
    }
}

constexpr int CSP_MAX_N = 64;

// csp = sum of the N/2 smallest |r_j + r_k|^2 over all neighbour pairs, added in ascending order
// (std::partial_sort then a forward sum, centro_symmetry_parameter.cpp:82-90).
__global__ void __launch_bounds__(128) k_csp(const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, int N, DBox box,
                                             const int *__restrict__ verlet, int M, int nnei,
                                             double *__restrict__ csp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double xi = x[i], yi = y[i], zi = z[i];
    double rx[CSP_MAX_N], ry[CSP_MAX_N], rz[CSP_MAX_N];
    const int *row = verlet + (size_t)i * M;
    for (int a = 0; a < nnei; ++a) {
        const int j = row[a];
        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
        min_image(box, dx, dy, dz);
        rx[a] = dx;
        ry[a] = dy;
        rz[a] = dz;
    }
    const int half = nnei / 2;
    double best[CSP_MAX_N / 2];  // ascending
    int nb = 0;
    for (int a = 0; a < nnei; ++a)
        for (int b = a + 1; b < nnei; ++b) {
            const double sx = rx[a] + rx[b], sy = ry[a] + ry[b], sz = rz[a] + rz[b];
            const double v = sx * sx + sy * sy + sz * sz;
            if (nb == half && !(v < best[half - 1])) continue;
            int pos = nb < half ? nb : half - 1;
            while (pos > 0 && best[pos - 1] > v) {
                best[pos] = best[pos - 1];
                --pos;
            }
            best[pos] = v;
            if (nb < half) ++nb;
        }
    double sum = 0.0;
    for (int q = 0; q < half; ++q) sum += best[q];
    csp[i] = sum;
}

__global__ void __launch_bounds__(128) k_aja(const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, int N, DBox box,
                                             const int *__restrict__ verlet, int M, const double *__restrict__ dist,
                                             int Md, int *__restrict__ aja)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double *di = dist + (size_t)i * Md;
    const int *vi = verlet + (size_t)i * M;
    double dloc[14];
    for (int j = 0; j < 14; ++j) dloc[j] = di[j];
    double r0_sq = 0.0;
    for (int j = 0; j < 6; ++j) r0_sq += dloc[j] * dloc[j];
    r0_sq /= 6.0;
    int N0 = 0, N1 = 0;
    const double r145 = 1.45 * r0_sq, r155 = 1.55 * r0_sq;
    for (int j = 0; j < 14; ++j) {
        const double r2 = dloc[j] * dloc[j];
        if (r2 < r155) {
            ++N1;
            if (r2 < r145) ++N0;
        }
    }
    const double xi = x[i], yi = y[i], zi = z[i];
    double rx[14], ry[14], rz[14];
    for (int j = 0; j < N0; ++j) {
        const int a = vi[j];
        double dx = x[a] - xi, dy = y[a] - yi, dz = z[a] - zi;
        min_image(box, dx, dy, dz);
        rx[j] = dx;
        ry[j] = dy;
        rz[j] = dz;
    }
    int al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < N0; ++j)
        for (int k = j + 1; k < N0; ++k) {
            const double dot = rx[j] * rx[k] + ry[j] * ry[k] + rz[j] * rz[k];
            const double c = dot / (dloc[j] * dloc[k]);
            int b;
            if (c < -0.945) b = 0;
            else if (c < -0.915) b = 1;
            else if (c < -0.755) b = 2;
            else if (c < -0.195) b = 3;
            else if (c < 0.195) b = 4;
            else if (c < 0.245) b = 5;
            else if (c < 0.795) b = 6;
            else b = 7;
            ++al[b];
        }
    const double sigma_cp = fabs(1.0 - al[6] / 24.0);
    const int s56m4 = al[5] + al[6] - al[4];
    double sigma_bcc = sigma_cp + 1.0;
    if (s56m4 != 0) sigma_bcc = 0.35 * al[4] / static_cast<double>(s56m4);
    double sigma_fcc = 0.61 * (abs(al[0] + al[1] - 6) + al[2]) / 6.0;
    double sigma_hcp = (fabs(al[0] - 3.0) + abs(al[0] + al[1] + al[2] + al[3] - 9)) / 12.0;
    if (al[0] == 7) sigma_bcc = 0.0;
    else if (al[0] == 6) sigma_fcc = 0.0;
    else if (al[0] <= 3) sigma_hcp = 0.0;
    int t;
    if (al[7] > 0) t = 0;
    else if (al[4] < 3) t = (N1 > 13 || N1 < 11) ? 0 : 4;
    else if (sigma_bcc <= sigma_cp) t = (N1 < 11) ? 0 : 3;
    else if (N1 > 12 || N1 < 11) t = 0;
    else t = (sigma_fcc < sigma_hcp) ? 1 : 2;
    aja[i] = t;
}

}  // namespace

void launch_sort_rows(MdbSystem &s, int *verlet, double *dist, int N, int M, int k)
{
    if (N <= 0 || M <= 0 || k <= 0) return;
    MDB_LAUNCH(k_sort_rows, (N + 127) / 128, 128, 0, s.stream, verlet, dist, N, M, k);
    CUDA_TRY(cudaGetLastError());
}

void launch_csp(MdbSystem &s, const int *verlet, int M, int nnei, double *csp)
{
    MDB_REQUIRE(nnei > 0 && nnei % 2 == 0, MDB_ERR_VALUE, "N must be a positive even number: %d.", nnei);
    MDB_REQUIRE(nnei <= CSP_MAX_N, MDB_ERR_VALUE, "N=%d exceeds the device limit %d", nnei, CSP_MAX_N);
    MDB_REQUIRE(nnei <= M, MDB_ERR_VALUE, "N=%d exceeds neighbour row width %d", nnei, M);
    MDB_LAUNCH(k_csp, (s.n_rows + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, s.n_rows, s.box, verlet, M, nnei, csp);
    CUDA_TRY(cudaGetLastError());
}

void launch_aja(MdbSystem &s, const int *verlet, int M, const double *dist, int Md, int *aja)
{
    MDB_REQUIRE(M >= 14 && Md >= 14, MDB_ERR_VALUE, "Ackland-Jones needs >= 14 sorted neighbours, row width is %d", M);
    MDB_LAUNCH(k_aja, (s.n_rows + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, s.n_rows, s.box, verlet, M, dist, Md, aja);
    CUDA_TRY(cudaGetLastError());
}
