// mdapy_b200/csrc/chill_bond.cu
//
// Two more consumers of the cut-off list:
//   CHILL+ water-phase identification   src/chill_plus.cpp:76-181  (Nguyen & Molinero 2015)
//   build_bond                          src/build_bond.cpp:9-88    bond pairs (i < j) under a type-pair cut-off matrix
//
// CHILL+ is single-precision in the reference (std::complex<float>): q_3m(i) = sum_j Y_3m(r_ij) over the listed
// neighbours within rc, then c_ij = Re(q_i . q_j*) / (|q_i| |q_j|) per bond, thresholds -> label.  The kernels
// below evaluate the same float expressions in the same order (left-to-right, no FMA contraction); the only
// operations that are not bit-defined by IEEE are atan2f / sinf / cosf / the complex exponential (libm on the
// CPU, CUDA's math library here, <= 2 ulp apart), so a bond correlation can differ in the last float bits and a
// label only when c_ij sits within ~1e-6 of a threshold.
#include "internal.cuh"

namespace {

struct cf {
    float re, im;
};
__device__ __forceinline__ cf cmul(cf a, cf b) { return cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cf conjf_(cf a) { return cf{a.re, -a.im}; }
__device__ __forceinline__ cf scale(float s, cf a) { return cf{s * a.re, s * a.im}; }

struct ChillConst {
    float N0, N1, N2, N3;
};

__device__ __forceinline__ void chill_y3m(const ChillConst &K, double dx, double dy, double dz, cf y[7])
{
    const float r2 = static_cast<float>(dx * dx + dy * dy + dz * dz);
    if (r2 <= 0.0f) {
        for (int i = 0; i < 7; ++i) y[i] = cf{0.0f, 0.0f};
        return;
    }
    const float r = sqrtf(r2);
    const float ct = static_cast<float>(dz) / r;
    const float xy = sqrtf(static_cast<float>(dx * dx + dy * dy));
    const float st = xy / r;
    const float phi = atan2f(static_cast<float>(dy), static_cast<float>(dx));
    const float ct2 = ct * ct, ct3 = ct2 * ct, st2 = st * st, st3 = st2 * st;
    float s1, c1;
    sincosf(phi, &s1, &c1);
    const cf e1{c1, s1};                 // exp(i phi)
    const cf e2 = cmul(e1, e1), e3 = cmul(e2, e1);
    const cf en1 = conjf_(e1), en2 = conjf_(e2), en3 = conjf_(e3);
    const float a1 = K.N1 * st * (5.0f * ct2 - 1.0f);
    const float a2 = K.N2 * st2 * ct;
    const float a3 = K.N3 * st3;
    y[0] = scale(a3, en3);
    y[1] = scale(a2, en2);
    y[2] = scale(a1, en1);
    y[3] = cf{K.N0 * (5.0f * ct3 - 3.0f * ct), 0.0f};
    y[4] = scale(-a1, e1);
    y[5] = scale(a2, e2);
    y[6] = scale(-a3, e3);
}

// q[i][0..6]: one thread per atom (rows of the list), neighbour ids index the N_all local atoms
__global__ void __launch_bounds__(128) k_chill_q(const double *__restrict__ x, const double *__restrict__ y,
                                                 const double *__restrict__ z, int rows, DBox box,
                                                 const int *__restrict__ verlet, const double *__restrict__ dist,
                                                 const int *__restrict__ nn, int M, double rc, ChillConst K,
                                                 float2 *__restrict__ q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int n = min(nn[i], M);
    const double xi = x[i], yi = y[i], zi = z[i];
    cf acc[7];
    for (int k = 0; k < 7; ++k) acc[k] = cf{0.0f, 0.0f};
    for (int jj = 0; jj < n; ++jj) {
        if (dist[(size_t)i * M + jj] > rc) continue;
        const int j = verlet[(size_t)i * M + jj];
        if (j < 0) continue;
        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
        min_image(box, dx, dy, dz);
        cf y3[7];
        chill_y3m(K, dx, dy, dz, y3);
        for (int k = 0; k < 7; ++k) {
            acc[k].re += y3[k].re;
            acc[k].im += y3[k].im;
        }
    }
    for (int k = 0; k < 7; ++k) q[(size_t)i * 7 + k] = make_float2(acc[k].re, acc[k].im);
}

__global__ void __launch_bounds__(128) k_chill_classify(int rows, const int *__restrict__ verlet,
                                                        const double *__restrict__ dist, const int *__restrict__ nn,
                                                        int M, double rc, const float2 *__restrict__ q,
                                                        int *__restrict__ pattern)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int n = min(nn[i], M);
    float2 qi[7];
    float qi_norm = 0.0f;
    for (int k = 0; k < 7; ++k) {
        qi[k] = q[(size_t)i * 7 + k];
        qi_norm += qi[k].x * qi[k].x + qi[k].y * qi[k].y;
    }
    int ecl = 0, stag = 0, coord = 0;
    for (int jj = 0; jj < n; ++jj) {
        if (dist[(size_t)i * M + jj] > rc) continue;
        const int j = verlet[(size_t)i * M + jj];
        if (j < 0) continue;
        float c_re_acc = 0.0f, c_im_acc = 0.0f, qj_norm = 0.0f;
        for (int k = 0; k < 7; ++k) {
            const float2 b = q[(size_t)j * 7 + k];
            // qi * conj(qj) = (a.re b.re + a.im b.im) + i (a.im b.re - a.re b.im), evaluated as the product with
            // the conjugate (re*re - im*(-im)), like std::complex's operator*
            const float pr = qi[k].x * b.x - qi[k].y * (-b.y);
            const float pi = qi[k].x * (-b.y) + qi[k].y * b.x;
            c_re_acc += pr;
            c_im_acc += pi;
            qj_norm += b.x * b.x + b.y * b.y;
        }
        (void)c_im_acc;
        const float denom = sqrtf(qi_norm) * sqrtf(qj_norm);
        const float c_re = denom > 0.0f ? c_re_acc / denom : 0.0f;
        if (c_re > -0.35f && c_re < 0.25f) ++ecl;
        if (c_re < -0.8f) ++stag;
        ++coord;
    }
    int code = 0;
    if (coord == 4) {
        if (ecl == 4) code = 4;
        else if (ecl == 3) code = 5;
        else if (stag == 4) code = 2;
        else if (stag == 3 && ecl == 1) code = 1;
        else if (stag == 3 && ecl == 0) code = 3;
        else if (stag == 2) code = 3;
    }
    pattern[i] = code;
}

// ---- build_bond: pairs (i, j), j > i, listed distance <= cutoff[type_i][type_j]; rows in (i, slot) order
__global__ void __launch_bounds__(256) k_bond_count(int rows, const int *__restrict__ verlet, const double *__restrict__ dist,
                                                    const int *__restrict__ nn, int M, const int *__restrict__ types,
                                                    const double *__restrict__ cutoff, int ntype, int *__restrict__ count,
                                                    const int *__restrict__ offs, int *__restrict__ bonds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int it = types[i];
    const int n = min(nn[i], M);
    int c = 0;
    int *out = bonds ? bonds + 2 * (size_t)offs[i] : nullptr;
    for (int jj = 0; jj < n; ++jj) {
        const int j = verlet[(size_t)i * M + jj];
        if (j <= i) continue;
        const int jt = types[j];
        if (it < 0 || it >= ntype || jt < 0 || jt >= ntype) continue;
        if (dist[(size_t)i * M + jj] <= cutoff[it * ntype + jt]) {
            if (out) {
                out[2 * c] = i;
                out[2 * c + 1] = j;
            }
            ++c;
        }
    }
    if (!bonds) count[i] = c;
}

}  // namespace

void launch_chill_plus(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, double rc, int *pattern)
{
    const int R = s.n_rows;
    if (R <= 0) return;
    constexpr float PI = 3.14159265358979323846f;
    ChillConst K;
    K.N0 = 0.25f * std::sqrt(7.0f / PI);
    K.N1 = 0.125f * std::sqrt(21.0f / PI);
    K.N2 = 0.25f * std::sqrt(105.0f / (2.0f * PI));
    K.N3 = 0.125f * std::sqrt(35.0f / PI);
    float2 *q = reinterpret_cast<float2 *>(s.scratch.ensure<float>((size_t)s.N * 14));
    if (s.N > R) CUDA_TRY(cudaMemsetAsync(q + (size_t)R * 7, 0, sizeof(float2) * 7 * (size_t)(s.N - R), s.stream));
    MDB_LAUNCH(k_chill_q, (R + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, R, s.box, verlet, dist, nn, M, rc, K, q);
    MDB_LAUNCH(k_chill_classify, (R + 127) / 128, 128, 0, s.stream, R, verlet, dist, nn, M, rc, q, pattern);
    CUDA_TRY(cudaGetLastError());
}

// returns the number of bonds; *bonds_dev points at 2 * nbond ints (valid until the next call on this handle)
int launch_build_bond(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, const int *types_dev,
                      const double *cutoff_dev, int ntype, int **bonds_dev)
{
    const int R = s.n_rows;
    *bonds_dev = nullptr;
    if (R <= 0) return 0;
    int *count = s.scratch.ensure<int>((size_t)R + 1);
    int *offs = s.scratch2.ensure<int>((size_t)R + 1);
    MDB_LAUNCH(k_bond_count, (R + 255) / 256, 256, 0, s.stream, R, verlet, dist, nn, M, types_dev, cutoff_dev, ntype, count,
               nullptr, nullptr);
    CUDA_TRY(cudaMemsetAsync(count + R, 0, sizeof(int), s.stream));
    device_exclusive_scan(s, count, offs, R + 1);
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, offs + R, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (total <= 0) return 0;
    int *bonds = s.verlet_tmp.ensure<int>((size_t)total * 2);
    MDB_LAUNCH(k_bond_count, (R + 255) / 256, 256, 0, s.stream, R, verlet, dist, nn, M, types_dev, cutoff_dev, ntype, count,
               offs, bonds);
    CUDA_TRY(cudaGetLastError());
    *bonds_dev = bonds;
    return total;
}
