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

// The same swap sequence with the block's 128 rows staged in shared memory: the rows are read and written once,
// coalesced (they are contiguous in HBM), and the k scans of a row run on shared memory.  Element q of thread t
// lives at [q * 129 + t]: the padding keeps both the cooperative copy (consecutive q) and the per-thread scans
// (consecutive t) on distinct banks.
constexpr int SORT_LD = 129;
__global__ void __launch_bounds__(128) k_sort_rows_staged(int *__restrict__ verlet, double *__restrict__ dist, int N, int M,
                                                          int k)
{
    extern __shared__ __align__(16) unsigned char sort_smem[];
    double *d = reinterpret_cast<double *>(sort_smem);
    int *v = reinterpret_cast<int *>(d + (size_t)M * SORT_LD);
    const int tid = threadIdx.x;
    const size_t row0 = (size_t)blockIdx.x * 128;
    const int rows = (int)((size_t)N - row0 < 128 ? (size_t)N - row0 : 128);
    const int total = rows * M;
    const size_t base = row0 * M;
    for (int e = tid; e < total; e += 128) {
        const int r = e / M, q = e - r * M;
        d[q * SORT_LD + r] = dist[base + e];
        v[q * SORT_LD + r] = verlet[base + e];
    }
    __syncthreads();
    if (tid < rows) {
        const int eff = k < M ? k : M;
        for (int j = 0; j < eff; ++j) {
            int mi = j;
            double md = d[j * SORT_LD + tid];
            for (int q = j + 1; q < M; ++q) {
                const double dq = d[q * SORT_LD + tid];
                if (dq < md) {
                    md = dq;
                    mi = q;
                }
            }
            if (mi != j) {
                const double td = d[j * SORT_LD + tid];
                d[j * SORT_LD + tid] = md;
                d[mi * SORT_LD + tid] = td;
                const int tv = v[j * SORT_LD + tid];
                v[j * SORT_LD + tid] = v[mi * SORT_LD + tid];
                v[mi * SORT_LD + tid] = tv;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < total; e += 128) {
        const int r = e / M, q = e - r * M;
        dist[base + e] = d[q * SORT_LD + r];
        verlet[base + e] = v[q * SORT_LD + r];
    }
}

constexpr int CSP_MAX_N = 64;

// csp for a compile-time neighbour count: vectors and the running N/2 smallest pair values stay in registers
// (every loop unrolls).  Same arithmetic and the same order of additions as k_csp below.
template <int NN>
__global__ void __launch_bounds__(128) k_csp_fixed(const double *__restrict__ x, const double *__restrict__ y,
                                                   const double *__restrict__ z, int N, DBox box,
                                                   const int *__restrict__ verlet, int M, double *__restrict__ csp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double xi = x[i], yi = y[i], zi = z[i];
    double rx[NN], ry[NN], rz[NN];
    const int *row = verlet + (size_t)i * M;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const int j = row[a];
        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
        min_image(box, dx, dy, dz);
        rx[a] = dx, ry[a] = dy, rz[a] = dz;
    }
    constexpr int HALF = NN / 2;
    double best[HALF];   // ascending; +inf until filled (a value equal to best[HALF-1] is not inserted, as in k_csp)
#pragma unroll
    for (int q = 0; q < HALF; ++q) best[q] = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int a = 0; a < NN; ++a)
#pragma unroll
        for (int b = a + 1; b < NN; ++b) {
            const double sx = rx[a] + rx[b], sy = ry[a] + ry[b], sz = rz[a] + rz[b];
            const double v = sx * sx + sy * sy + sz * sz;
            // Insertion into the ascending list, branch free.  With c = number of entries <= v (ties stay in front
            // of the newcomer, like the strict comparison of k_csp's shifting loop): entries above c move up one
            // slot, slot c takes v, the largest entry falls out; v >= every entry changes nothing.
#pragma unroll
            for (int q = HALF - 1; q >= 0; --q) {
                const double below = q > 0 ? best[q - 1] : 0.0;
                best[q] = (q > 0 && below > v) ? below : (best[q] > v ? v : best[q]);
            }
        }
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < HALF; ++q) sum += best[q];
    csp[i] = sum;
}

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


// Common neighbour parameter, src/common_neighbor_parameter.cpp:40-136: for every listed neighbour j
// within rc, R_ij = sum over common neighbours k (within rc of both) of (r_ik + r_jk), each a
// min-image vector; cnp_i = sum_j |R_ij|^2 / N_i, 1000 when no neighbour lies within rc.  Same loop
// nesting and summation order as the reference.
// CACHE: the row of atom i (indices + "within rc" flags) sits in shared memory, interleaved per thread, because the
// innermost loop searches it for every neighbour of every neighbour (M^3 probes per atom); rows wider than the
// cache (CNP_CACHE entries) read global memory as before.
constexpr int CNP_CACHE = 48;
template <bool CACHED>
__global__ void __launch_bounds__(128) k_cnp(const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, int N, DBox box,
                                             const int *__restrict__ verlet, const double *__restrict__ dist,
                                             const int *__restrict__ nn, int M, double rc, double *__restrict__ cnp)
{
    __shared__ int sh_idx[CACHED ? CNP_CACHE * 128 : 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int ni = min(nn[i], M);
    const int *vi = verlet + (size_t)i * M;
    const double *di = dist + (size_t)i * M;
    const double xi = x[i], yi = y[i], zi = z[i];
    int *row = sh_idx + threadIdx.x;
    unsigned long long in_rc = 0ull;   // bit h: di[h] <= rc (CACHED: ni <= 48)
    if (CACHED)
        for (int h = 0; h < ni; ++h) {
            row[h * 128] = vi[h];
            if (di[h] <= rc) in_rc |= 1ull << h;
        }
    int cnt = 0;
    double acc = 0.0;
    for (int m = 0; m < ni; ++m) {
        if (CACHED ? !((in_rc >> m) & 1ull) : !(di[m] <= rc)) continue;
        const int j = CACHED ? row[m * 128] : vi[m];
        ++cnt;
        double rx = 0.0, ry = 0.0, rz = 0.0;
        const int nj = min(nn[j], M);
        const int *vj = verlet + (size_t)j * M;
        const double *dj = dist + (size_t)j * M;
        const double xj = x[j], yj = y[j], zj = z[j];
        for (int s = 0; s < nj; ++s) {
            const int k = vj[s];
            for (int h = 0; h < ni; ++h) {
                if ((CACHED ? row[h * 128] : vi[h]) != k) continue;
                if (dj[s] <= rc && (CACHED ? ((in_rc >> h) & 1ull) != 0ull : di[h] <= rc)) {
                    const double xk = x[k], yk = y[k], zk = z[k];
                    double ax = xi - xk, ay = yi - yk, az = zi - zk;
                    double bx = xj - xk, by = yj - yk, bz = zj - zk;
                    min_image(box, ax, ay, az);
                    min_image(box, bx, by, bz);
                    rx += ax + bx;
                    ry += ay + by;
                    rz += az + bz;
                }
                break;  // the reference leaves the h loop at the first index match
            }
        }
        acc += rx * rx + ry * ry + rz * rz;
    }
    cnp[i] = cnt > 0 ? acc / cnt : 1000.0;
}

// Warren-Cowley counts, src/warren_cowley_parameter.cpp:33-49: Z_mn (neighbour-type pairs), Z_m (listed
// neighbours per central type), population per type.  counts = [Zmn (T*T) | Zm (T) | pop (T)], 64-bit.
__global__ void __launch_bounds__(256) k_wcp_counts(const int *__restrict__ verlet, const int *__restrict__ nn, int N,
                                                    int M, const int *__restrict__ types, int T,
                                                    unsigned long long *__restrict__ counts)
{
    extern __shared__ unsigned sh_w[];
    const int nslot = T * T + 2 * T;
    for (int t = threadIdx.x; t < nslot; t += blockDim.x) sh_w[t] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int it = types[i];
        const int c = nn[i];
        atomicAdd(sh_w + T * T + T + it, 1u);
        atomicAdd(sh_w + T * T + it, (unsigned)c);
        for (int q = 0; q < c; ++q) atomicAdd(sh_w + it * T + types[verlet[(size_t)i * M + q]], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nslot; t += blockDim.x)
        if (sh_w[t]) atomicAdd(counts + t, (unsigned long long)sh_w[t]);
}

// average_by_neighbor, src/neighbor.cpp:704-743: own value first (include_self), then the listed
// neighbours within rc in list order.
__global__ void __launch_bounds__(128) k_average_by_neighbor(const int *__restrict__ verlet,
                                                             const double *__restrict__ dist,
                                                             const int *__restrict__ nn, int N, int M, double rc,
                                                             const double *__restrict__ value, int include_self,
                                                             double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double sum = 0.0;
    int n = 0;
    if (include_self) {
        sum += value[i];
        ++n;
    }
    const int c = nn[i];
    for (int q = 0; q < c; ++q)
        if (dist[(size_t)i * M + q] <= rc) {
            sum += value[verlet[(size_t)i * M + q]];
            ++n;
        }
    out[i] = n > 0 ? sum / n : 0.0;
}


// Local structural entropy (pair-entropy fingerprint), src/structure_entropy.cpp:11-103.  Per atom:
// g_m(r_j) = sum_k exp(-(r_j - d_k)^2 / (2 sigma^2)) / prefactor_j over listed distances <= rc, optional
// local-density rescale, integrand (g ln g - g + 1) r^2, trapezoid sum.  Same loop nesting and summation
// order as the reference (bins outer, neighbours inner); exp / log are the device's (<= 1 ulp from libm).
__global__ void __launch_bounds__(128) k_structure_entropy(const double *__restrict__ dist, const int *__restrict__ nn,
                                                           int N, int M, double rc, double sigma, int use_local_density,
                                                           double global_density, int nbins, double *__restrict__ entropy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double MY_PI = 3.14159265358979323846;
    const double step = rc / (nbins - 1);
    const double factor = (4. * MY_PI * global_density * sqrt(2. * MY_PI * sigma * sigma));
    const double sigma_sq = sigma * sigma;
    const double local_vol = 4. / 3. * MY_PI * rc * rc * rc;
    const double *di = dist + (size_t)i * M;
    const int c = nn[i];
    int n_neigh = 0;
    for (int k = 0; k < c; ++k) n_neigh += di[k] <= rc ? 1 : 0;
    double density = global_density, fac = 1.0;
    if (use_local_density) {
        density = n_neigh / local_vol;
        fac = global_density / density;
    }
    const double r1 = 1 * step;
    const double pref1 = (r1 * r1) * factor;
    double sum = 0.0, prev = 0.0;
    for (int j = 0; j < nbins; ++j) {
        const double rj = j * step;
        const double rsq = rj * rj;
        const double pref = j == 0 ? pref1 : rsq * factor;
        double g = 0.0;
        for (int k = 0; k < c; ++k) {
            const double dis = di[k];
            if (dis <= rc) {
                const double delta = rj - dis;
                g += exp(-(delta * delta) / (2.0 * sigma_sq)) / pref;
            }
        }
        if (use_local_density) g *= fac;
        const double integrand = g >= 1e-10 ? (g * log(g) - g + 1.0) * rsq : rsq;
        if (j > 0) sum += prev + integrand;
        prev = integrand;
    }
    entropy[i] = -MY_PI * density * sum * sigma;
}


// Local atomic temperature, src/atomic_temperature.cpp:7-112: mass-weighted mean velocity of the atom and its
// listed neighbours within rc, kinetic energy of the fluctuations about it, T = 2 KE / (3 n k_B).  Velocities
// arrive in A/ps (the host multiplies A/fs by 1e3 * factor like the reference's wrapper).  Same operation order.
__global__ void __launch_bounds__(128) k_atomic_temperature(const int *__restrict__ verlet, const double *__restrict__ dist,
                                                            int N, int M, const double *__restrict__ vx,
                                                            const double *__restrict__ vy, const double *__restrict__ vz,
                                                            const double *__restrict__ mass, double rc,
                                                            double *__restrict__ T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double kb = 1.380649e-23, dim = 3.0, afu = 6.022140857e23;
    const double mass_factor = 1.0 / afu / 1000.0;
    const double vel_conv = 1e4;
    const int *vi = verlet + (size_t)i * M;
    const double *di = dist + (size_t)i * M;
    const double mass_i = mass[i];
    double sx = vx[i] * mass_i, sy = vy[i] * mass_i, sz = vz[i] * mass_i;
    int n_neigh = 1;
    double mass_neigh = mass_i;
    for (int q = 0; q < M; ++q) {
        const int j = vi[q];
        if (j < 0) break;
        if (j != i && di[q] <= rc) {
            const double mj = mass[j];
            sx += vx[j] * mj;
            sy += vy[j] * mj;
            sz += vz[j] * mj;
            ++n_neigh;
            mass_neigh += mj;
        }
    }
    const double mx = sx / mass_neigh, my = sy / mass_neigh, mz = sz / mass_neigh;
    double ke = 0.0;
    double dx = vx[i] - mx, dy = vy[i] - my, dz = vz[i] - mz;
    double vsq = dx * dx + dy * dy + dz * dz;
    ke += 0.5 * mass_i * mass_factor * vsq * vel_conv;
    for (int q = 0; q < M; ++q) {
        const int j = vi[q];
        if (j < 0) break;
        if (j != i && di[q] <= rc) {
            const double mj = mass[j];
            dx = vx[j] - mx;
            dy = vy[j] - my;
            dz = vz[j] - mz;
            vsq = dx * dx + dy * dy + dz * dz;
            ke += 0.5 * mj * mass_factor * vsq * vel_conv;
        }
    }
    T[i] = ke * 2.0 / (dim * n_neigh * kb);
}


// Bond-length / bond-angle histograms, src/bond_analysis.cpp:7-118 (compute_bond): lengths of listed pairs
// j > i within rc, angles j-i-k over listed neighbour pairs jj < kk within rc; angle = acos(clamped cos) * 180 / PI,
// bin = floor(theta / delta), clamped to the last bin.  hist: [nbins lengths | nbins angles] (64-bit).
// acos is the device's (<= 1 ulp from libm): a bin can differ only for an angle within ~1e-14 degrees of an
// edge, i.e. on perfect lattices whose angles sit exactly on bin edges.
__global__ void __launch_bounds__(128) k_bond_hist(const double *__restrict__ x, const double *__restrict__ y,
                                                   const double *__restrict__ z, int N, DBox box,
                                                   const int *__restrict__ verlet, const double *__restrict__ dist,
                                                   const int *__restrict__ nn, int M, double delta_r, double delta_theta,
                                                   double rc, int nbins, unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned sh_b[];
    const bool use_sh = 2 * nbins <= 8192;
    if (use_sh) {
        for (int t = threadIdx.x; t < 2 * nbins; t += blockDim.x) sh_b[t] = 0;
        __syncthreads();
    }
    const double PI = 3.14159265358979323846;
    const double dri = 1.0 / delta_r, dti = 1.0 / delta_theta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int c = nn[i];
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        for (int jj = 0; jj < c; ++jj) {
            if (vi[jj] > i) {
                const double r = di[jj];
                if (r <= rc) {
                    int index = static_cast<int>(floor(r * dri));
                    if (index > nbins - 1) index = nbins - 1;
                    if (use_sh) atomicAdd(sh_b + index, 1u);
                    else atomicAdd(hist + index, 1ull);
                }
            }
        }
        const double xi = x[i], yi = y[i], zi = z[i];
        for (int jj = 0; jj < c; ++jj) {
            const double rij = di[jj];
            if (!(rij <= rc)) continue;
            const int j = vi[jj];
            double ax = x[j] - xi, ay = y[j] - yi, az = z[j] - zi;
            min_image(box, ax, ay, az);
            for (int kk = jj + 1; kk < c; ++kk) {
                const double rik = di[kk];
                if (!(rik <= rc)) continue;
                const int k = vi[kk];
                double bx = x[k] - xi, by = y[k] - yi, bz = z[k] - zi;
                min_image(box, bx, by, bz);
                const double dot = ax * bx + ay * by + az * bz;
                double ct = dot / (rij * rik);
                if (ct > 1.0) ct = 1.0;
                if (ct < -1.0) ct = -1.0;
                const double theta = acos(ct) * 180.0 / PI;
                int index = static_cast<int>(floor(theta * dti));
                if (index > nbins - 1) index = nbins - 1;
                if (use_sh) atomicAdd(sh_b + nbins + index, 1u);
                else atomicAdd(hist + nbins + index, 1ull);
            }
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * nbins; t += blockDim.x)
            if (sh_b[t]) atomicAdd(hist + t, (unsigned long long)sh_b[t]);
    }
}

// Angular distribution function per (centre, j, k) type triplet, src/bond_analysis.cpp:120-240 (compute_adf).
// pairs: [Npair][3] types, rcs: [Npair][4] = r_ij min, max, r_ik min, max.  hist: [Npair][nbins] (64-bit).
__global__ void __launch_bounds__(128) k_adf_hist(const double *__restrict__ x, const double *__restrict__ y,
                                                  const double *__restrict__ z, int N, DBox box,
                                                  const int *__restrict__ verlet, const double *__restrict__ dist,
                                                  const int *__restrict__ nn, int M, double delta_theta,
                                                  const double *__restrict__ rcs, const int *__restrict__ pairs,
                                                  int npair, const int *__restrict__ types, int nbins,
                                                  unsigned long long *__restrict__ hist)
{
    // per-block 32-bit histogram in shared memory when it fits (a few type triplets x <= 8192 / npair bins):
    // thousands of atoms hammering the same ~100 global 64-bit counters serialise otherwise
    extern __shared__ unsigned sh_a[];
    const int nslot = npair * nbins;
    const bool use_sh = nslot <= 8192;
    if (use_sh) {
        for (int t = threadIdx.x; t < nslot; t += blockDim.x) sh_a[t] = 0;
        __syncthreads();
    }
    const double PI = 3.14159265358979323846;
    const double dti = 1.0 / delta_theta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int itype = types[i];
        const int c = nn[i];
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        const double xi = x[i], yi = y[i], zi = z[i];
        for (int m = 0; m < npair; ++m) {
            if (itype != pairs[m * 3]) continue;
            const int jt = pairs[m * 3 + 1], kt = pairs[m * 3 + 2];
            const bool same = jt == kt;
            for (int jj = 0; jj < c; ++jj) {
                const int j = vi[jj];
                if (types[j] != jt) continue;
                const double rij = di[jj];
                if (!(rij <= rcs[m * 4 + 1] && rij >= rcs[m * 4 + 0])) continue;
                double ax = x[j] - xi, ay = y[j] - yi, az = z[j] - zi;
                min_image(box, ax, ay, az);
                for (int kk = same ? jj + 1 : 0; kk < c; ++kk) {
                    if (kk == jj) continue;
                    const int k = vi[kk];
                    if (types[k] != kt) continue;
                    const double rik = di[kk];
                    if (!(rik <= rcs[m * 4 + 3] && rik >= rcs[m * 4 + 2])) continue;
                    double bx = x[k] - xi, by = y[k] - yi, bz = z[k] - zi;
                    min_image(box, bx, by, bz);
                    const double dot = ax * bx + ay * by + az * bz;
                    double ct = dot / (rij * rik);
                    if (ct > 1.0) ct = 1.0;
                    if (ct < -1.0) ct = -1.0;
                    const double theta = acos(ct) * 180.0 / PI;
                    int index = static_cast<int>(floor(theta * dti));
                    if (index < 0) index = 0;
                    if (index >= nbins) index = nbins - 1;
                    if (use_sh) atomicAdd(sh_a + m * nbins + index, 1u);
                    else atomicAdd(hist + (size_t)m * nbins + index, 1ull);
                }
            }
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int t = threadIdx.x; t < nslot; t += blockDim.x)
            if (sh_a[t]) atomicAdd(hist + t, (unsigned long long)sh_a[t]);
    }
}

}  // namespace

void launch_sort_rows(MdbSystem &s, int *verlet, double *dist, int N, int M, int k)
{
    if (N <= 0 || M <= 0 || k <= 0) return;
    const size_t smem = (size_t)M * SORT_LD * (sizeof(double) + sizeof(int));
    const char *mode = getenv("MDB_SORT");
    if (smem <= 200 * 1024 && !(mode && !strcmp(mode, "global"))) {
        CUDA_TRY(cudaFuncSetAttribute(k_sort_rows_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MDB_LAUNCH(k_sort_rows_staged, (N + 127) / 128, 128, smem, s.stream, verlet, dist, N, M, k);
    } else {
        MDB_LAUNCH(k_sort_rows, (N + 127) / 128, 128, 0, s.stream, verlet, dist, N, M, k);
    }
    CUDA_TRY(cudaGetLastError());
}

void launch_csp(MdbSystem &s, const int *verlet, int M, int nnei, double *csp)
{
    MDB_REQUIRE(nnei > 0 && nnei % 2 == 0, MDB_ERR_VALUE, "N must be a positive even number: %d.", nnei);
    MDB_REQUIRE(nnei <= CSP_MAX_N, MDB_ERR_VALUE, "N=%d exceeds the device limit %d", nnei, CSP_MAX_N);
    MDB_REQUIRE(nnei <= M, MDB_ERR_VALUE, "N=%d exceeds neighbour row width %d", nnei, M);
    const int R = s.n_rows, nb = (R + 127) / 128;
    const char *mode = getenv("MDB_CSP");
    const bool generic = mode && !strcmp(mode, "generic");
    if (nnei == 12 && !generic) MDB_LAUNCH(k_csp_fixed<12>, nb, 128, 0, s.stream, s.x, s.y, s.z, R, s.box, verlet, M, csp);
    else if (nnei == 8 && !generic) MDB_LAUNCH(k_csp_fixed<8>, nb, 128, 0, s.stream, s.x, s.y, s.z, R, s.box, verlet, M, csp);
    else if (nnei == 14 && !generic) MDB_LAUNCH(k_csp_fixed<14>, nb, 128, 0, s.stream, s.x, s.y, s.z, R, s.box, verlet, M, csp);
    else MDB_LAUNCH(k_csp, nb, 128, 0, s.stream, s.x, s.y, s.z, R, s.box, verlet, M, nnei, csp);
    CUDA_TRY(cudaGetLastError());
}

void launch_aja(MdbSystem &s, const int *verlet, int M, const double *dist, int Md, int *aja)
{
    MDB_REQUIRE(M >= 14 && Md >= 14, MDB_ERR_VALUE, "Ackland-Jones needs >= 14 sorted neighbours, row width is %d", M);
    MDB_LAUNCH(k_aja, (s.n_rows + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, s.n_rows, s.box, verlet, M, dist, Md, aja);
    CUDA_TRY(cudaGetLastError());
}

void launch_cnp(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, double *cnp)
{
    const int N = s.n_rows;
    if (M <= CNP_CACHE) MDB_LAUNCH(k_cnp<true>, (N + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, N, s.box, verlet, dist, nn, M, rc, cnp);
    else MDB_LAUNCH(k_cnp<false>, (N + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, N, s.box, verlet, dist, nn, M, rc, cnp);
    CUDA_TRY(cudaGetLastError());
}

// counts (device, 64-bit): [T*T | T | T] as described at k_wcp_counts; zeroed here
void launch_wcp_counts(MdbSystem &s, const int *verlet, const int *nn, int M, const int *types, int T,
                       unsigned long long *counts)
{
    const int N = s.n_rows;
    const int nslot = T * T + 2 * T;
    MDB_REQUIRE(T >= 1 && nslot <= 8192, MDB_ERR_VALUE, "Ntype=%d is outside the supported range", T);
    CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * nslot, s.stream));
    int nb = (N + 255) / 256;
    // a block's 32-bit shared counters see at most 256 * ceil(N / (nb*256)) * M increments: keep that < 2^32
    if (nb > 1184) nb = 1184;
    MDB_LAUNCH(k_wcp_counts, nb, 256, sizeof(unsigned) * nslot, s.stream, verlet, nn, N, M, types, T, counts);
    CUDA_TRY(cudaGetLastError());
}

void launch_average_by_neighbor(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc,
                                const double *value, bool include_self, double *out)
{
    const int N = s.n_rows;
    MDB_LAUNCH(k_average_by_neighbor, (N + 127) / 128, 128, 0, s.stream, verlet, dist, nn, N, M, rc, value,
               include_self ? 1 : 0, out);
    CUDA_TRY(cudaGetLastError());
}

void launch_structure_entropy(MdbSystem &s, const double *dist, const int *nn, int M, double rc, double sigma,
                              bool use_local_density, double volume, double *entropy)
{
    const int N = s.n_rows;
    MDB_REQUIRE(rc > 0 && sigma > 0 && volume > 0, MDB_ERR_VALUE, "rc, sigma and the volume must be positive");
    const int nbins = static_cast<int>(std::floor(rc / sigma)) + 1;
    MDB_REQUIRE(nbins >= 2, MDB_ERR_VALUE, "sigma=%g is larger than rc=%g", sigma, rc);
    const double global_density = N / volume;
    MDB_LAUNCH(k_structure_entropy, (N + 127) / 128, 128, 0, s.stream, dist, nn, N, M, rc, sigma,
               use_local_density ? 1 : 0, global_density, nbins, entropy);
    CUDA_TRY(cudaGetLastError());
}

void launch_atomic_temperature(MdbSystem &s, const int *verlet, const double *dist, int M, const double *vx,
                               const double *vy, const double *vz, const double *mass, double rc, double *T)
{
    const int N = s.n_rows;
    MDB_LAUNCH(k_atomic_temperature, (N + 127) / 128, 128, 0, s.stream, verlet, dist, N, M, vx, vy, vz, mass, rc, T);
    CUDA_TRY(cudaGetLastError());
}

// hist (device, 64-bit, 2 * nbins) is zeroed here
void launch_bond_hist(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double delta_r,
                      double delta_theta, double rc, int nbins, unsigned long long *hist)
{
    const int N = s.n_rows;
    MDB_REQUIRE(nbins > 0 && delta_r > 0 && delta_theta > 0, MDB_ERR_VALUE, "nbins and the bin widths must be positive");
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * 2 * nbins, s.stream));
    int nb = (N + 127) / 128;
    if (nb > 148 * 16) nb = 148 * 16;
    const size_t smem = 2 * nbins <= 8192 ? sizeof(unsigned) * 2 * nbins : 0;
    MDB_LAUNCH(k_bond_hist, nb, 128, smem, s.stream, s.x, s.y, s.z, N, s.box, verlet, dist, nn, M, delta_r, delta_theta, rc,
               nbins, hist);
    CUDA_TRY(cudaGetLastError());
}

void launch_adf_hist(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double delta_theta,
                     const double *rcs, const int *pairs, int npair, const int *types, int nbins,
                     unsigned long long *hist)
{
    const int N = s.n_rows;
    MDB_REQUIRE(nbins > 0 && npair > 0 && delta_theta > 0, MDB_ERR_VALUE, "nbins, the triplet table and the bin width are required");
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (size_t)npair * nbins, s.stream));
    int nb = (N + 127) / 128;
    if (nb > 148 * 16) nb = 148 * 16;
    const size_t smem = (size_t)npair * nbins <= 8192 ? sizeof(unsigned) * (size_t)npair * nbins : 0;
    MDB_LAUNCH(k_adf_hist, nb, 128, smem, s.stream, s.x, s.y, s.z, N, s.box, verlet, dist, nn, M, delta_theta, rcs, pairs,
               npair, types, nbins, hist);
    CUDA_TRY(cudaGetLastError());
}
