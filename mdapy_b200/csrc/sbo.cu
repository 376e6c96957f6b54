// mdapy_b200/csrc/sbo.cu
//
// Steinhardt bond-orientational order on the device.  Replaces
// src/steinhardt_bond_orientation.cpp:288-576 (_compute_ql), 188-224 (Clebsch-Gordan table),
// 243-286 (Legendre / prefactor) and 578-675 (identifySolidLiquid).
//
// q_lm(i) = (1/W) sum_j w_ij Y_lm(r_ij) over listed neighbours with 1e-15 < r <= rc, r taken
// from the STORED distance list and the direction from min-image(x[j]-x[i]) (Appendix A).
// Sums run in list order with the reference's operation order (no FMA), so q_l, w_l agree with
// the reference to the last bit for a given list.  Everything that depends only on (l, m) --
// sqrt((2l+1)/(4 pi prod)), sqrt(4 pi/(2l+1)), the Clebsch-Gordan coefficients -- is evaluated
// on the host with the reference's expressions and passed in.
#include <mutex>
#include "internal.cuh"
#include <vector>

namespace {

constexpr int SBO_MAX_L = 24;
constexpr int SBO_LOCAL = 64;  // accumulators kept in registers/local memory when ndeg*(2lmax+1) <= this

struct SboParams {
    int ndeg, lmax, nz, nnn, use_voronoi, use_weight;
    double rc;
    int l[8];
    int off[9];  // compact accumulator slots: degree il owns [off[il], off[il] + 2 l + 1), off[ndeg] = total
};

__constant__ double c_norm[8][SBO_MAX_L + 1];  // sqrt((2l+1)/(4 pi prod_{i=l-m+1}^{l+m} i)) per (degree slot, m)

// a / d for a small positive integer d, correctly rounded (== the IEEE quotient the reference computes)
// without the ~35-instruction division sequence: with y = RN(1/d), q0 = RN(a y) is within an ulp of a/d and
// each residual step q <- RN(q + RN(a - d q) y) (both FMAs exact in the residual) lands on the correctly
// rounded quotient (Markstein); two steps leave no doubt for the rare q0 that is off by more than half an
// ulp.  tests/test_gpu_descriptors.py::test_small_integer_division compares 2^26 quotients per divisor.
__constant__ double c_rcp[SBO_MAX_L + 2];

__device__ __forceinline__ double div_small(double a, int d)
{
    const double y = c_rcp[d], dd = (double)d;
    const double q0 = a * y;
    double q = fma(fma(-dd, q0, a), y, q0);
    q = fma(fma(-dd, q, a), y, q);
    // +-0 (sign!), infinities and NaN: a * y is already the IEEE quotient and the residual steps would turn
    // it into NaN / +0.  (Subnormal operands, 1e-308 and below, are not refined either; Legendre values of
    // degree <= 24 never get there.)  Branch-free: biased exponent of a in [1, 2046] <=> finite normal.
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    return (e - 1u < 2046u) ? q : q0;
}

// _associated_legendre, cpp:243-268
__device__ __forceinline__ double assoc_legendre(int l, int m, double x)
{
    double p = 1.0, pm1 = 0.0, pm2 = 0.0;
    if (m != 0) {
        const double sqx = sqrt(1.0 - x * x);
        for (int i = 1; i < m + 1; ++i) p *= (2 * i - 1) * sqx;
    }
    for (int i = m + 1; i < l + 1; ++i) {
        pm2 = pm1;
        pm1 = p;
        p = div_small((2 * i - 1) * x * pm1 - (i + m - 1) * pm2, i - m);
    }
    return p;
}

// Two independent P_l^m chains in one loop (same l, m; arguments xa, xb): the serial FP64 recurrence of one
// neighbour leaves the pipe mostly idle, two interleaved chains hide each other's latency.  Every value is
// computed by exactly the operations of assoc_legendre().
__device__ __forceinline__ void assoc_legendre2(int l, int m, double xa, double xb, double &pa_out, double &pb_out)
{
    double pa = 1.0, pa1 = 0.0, pa2 = 0.0, pb = 1.0, pb1 = 0.0, pb2 = 0.0;
    if (m != 0) {
        const double sa = sqrt(1.0 - xa * xa), sb = sqrt(1.0 - xb * xb);
        for (int i = 1; i < m + 1; ++i) {
            pa *= (2 * i - 1) * sa;
            pb *= (2 * i - 1) * sb;
        }
    }
    for (int i = m + 1; i < l + 1; ++i) {
        pa2 = pa1;
        pa1 = pa;
        pb2 = pb1;
        pb1 = pb;
        pa = div_small((2 * i - 1) * xa * pa1 - (i + m - 1) * pa2, i - m);
        pb = div_small((2 * i - 1) * xb * pb1 - (i + m - 1) * pb2, i - m);
    }
    pa_out = pa;
    pb_out = pb;
}

// Accumulators: MODE 1 = per-thread local arrays (dense [il][m] indexing), MODE 2 = shared memory, compact slots
// interleaved per thread (acc[slot * blockDim + tid]: conflict-free, no local-memory traffic), MODE 0 = in place
// in global memory (very high degrees).  The operation order per accumulator is the same in all three.
template <int MODE>
__global__ void __launch_bounds__(128) k_qlm(const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, int N, DBox box,
                                             const int *__restrict__ verlet, const double *__restrict__ dist,
                                             const int *__restrict__ nn, const double *__restrict__ weight, int M,
                                             SboParams P, double *__restrict__ qr, double *__restrict__ qi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double EPS = 1e-15;
    const int stride = P.ndeg * P.nz;
    constexpr bool LOCAL = MODE == 1;
    extern __shared__ double sh_acc[];
    double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
    double ar[LOCAL ? SBO_LOCAL : 1], ai[LOCAL ? SBO_LOCAL : 1];
    if (LOCAL) {
        for (int t = 0; t < stride; ++t) {
            ar[t] = Qr[t];  // caller-zeroed, but accumulate like the reference (inout arrays)
            ai[t] = Qi[t];
        }
    }
    // element stride and per-degree base of the accumulators
    const int es = MODE == 2 ? (int)blockDim.x : 1;
    double *R = LOCAL ? ar : (MODE == 2 ? sh_acc + threadIdx.x : Qr);
    double *I = LOCAL ? ai : (MODE == 2 ? sh_acc + (size_t)P.off[P.ndeg] * blockDim.x + threadIdx.x : Qi);
    if (MODE == 2)
        for (int il = 0; il < P.ndeg; ++il)
            for (int m = 0; m < 2 * P.l[il] + 1; ++m) {
                R[(P.off[il] + m) * es] = Qr[il * P.nz + m];
                I[(P.off[il] + m) * es] = Qi[il * P.nz + m];
            }
    const double x1 = x[i], y1 = y[i], z1 = z[i];
    int cnt = nn[i];
    if (!P.use_voronoi && P.nnn > 0) cnt = P.nnn;
    double wsum = 0.0;
    // Neighbours are taken two at a time (A, then B): their Legendre chains run interleaved, and every
    // accumulator still receives A's term before B's, i.e. the reference's neighbour order.
    struct Prepared {
        double w, ct, er, ei;
    };
    int jj = 0;
    auto next_valid = [&](Prepared &nb) -> bool {
        for (; jj < cnt; ++jj) {
            const size_t at = (size_t)i * M + jj;
            const int j = verlet[at];
            if (j < 0) continue;
            double dx = x[j] - x1, dy = y[j] - y1, dz = z[j] - z1;
            min_image(box, dx, dy, dz);
            const double rmag = dist[at];
            if (!((rmag > EPS) && (rmag <= P.rc))) continue;
            nb.w = P.use_weight ? weight[at] : 1.0;
            const double rinv = 1.0 / rmag;
            nb.ct = dz * rinv;
            double er = dx, ei = dy;
            const double rxy2 = er * er + ei * ei;
            if (rxy2 < EPS * EPS) {
                er = 1.0;
                ei = 0.0;
            } else {
                const double inv = 1.0 / sqrt(rxy2);
                er *= inv;
                ei *= inv;
            }
            nb.er = er;
            nb.ei = ei;
            ++jj;
            return true;
        }
        return false;
    };
    for (;;) {
        Prepared A, B;
        if (!next_valid(A)) break;
        const bool hasB = next_valid(B);
        if (!hasB) B = Prepared{0.0, 0.0, 1.0, 0.0};
        wsum += A.w;
        if (hasB) wsum += B.w;
        for (int il = 0; il < P.ndeg; ++il) {
            const int l = P.l[il];
            double *Rl = R + (MODE == 2 ? P.off[il] * es : il * P.nz), *Il = I + (MODE == 2 ? P.off[il] * es : il * P.nz);
            double pa, pb;
            assoc_legendre2(l, 0, A.ct, B.ct, pa, pb);
            Rl[l * es] += A.w * (c_norm[il][0] * pa);
            if (hasB) Rl[l * es] += B.w * (c_norm[il][0] * pb);
            double pra = A.er, pia = A.ei, prb = B.er, pib = B.ei;
            for (int m = 1; m < l + 1; ++m) {
                assoc_legendre2(l, m, A.ct, B.ct, pa, pb);
                const double pfa = c_norm[il][m] * pa, pfb = c_norm[il][m] * pb;
                const double wra = A.w * (pfa * pra), wia = A.w * (pfa * pia);
                const double wrb = B.w * (pfb * prb), wib = B.w * (pfb * pib);
                const bool sgn_is_odd = (m & 1) != 0;
                // slot l + m
                Rl[(l + m) * es] += wra;
                Il[(l + m) * es] += wia;
                if (hasB) {
                    Rl[(l + m) * es] += wrb;
                    Il[(l + m) * es] += wib;
                }
                // slot l - m: (-1)^m conj
                if (sgn_is_odd) {
                    Rl[(l - m) * es] -= wra;
                    Il[(l - m) * es] += wia;
                    if (hasB) {
                        Rl[(l - m) * es] -= wrb;
                        Il[(l - m) * es] += wib;
                    }
                } else {
                    Rl[(l - m) * es] += wra;
                    Il[(l - m) * es] -= wia;
                    if (hasB) {
                        Rl[(l - m) * es] += wrb;
                        Il[(l - m) * es] -= wib;
                    }
                }
                const double tra = pra * A.er - pia * A.ei, tia = pra * A.ei + pia * A.er;
                const double trb = prb * B.er - pib * B.ei, tib = prb * B.ei + pib * B.er;
                pra = tra;
                pia = tia;
                prb = trb;
                pib = tib;
            }
        }
    }
    const double fac = 1.0 / wsum;
    for (int il = 0; il < P.ndeg; ++il) {
        const int mm = 2 * P.l[il] + 1;
        const int base = MODE == 2 ? P.off[il] : il * P.nz;
        for (int m = 0; m < mm; ++m) {
            Qr[il * P.nz + m] = R[(base + m) * es] * fac;
            Qi[il * P.nz + m] = I[(base + m) * es] * fac;
        }
    }
}

// neighbour averaging (Lechner-Dellago), cpp:439-503: every listed neighbour counts, no rc filter.
// Same summation order as the reference (own q_lm first, then neighbours in list order).
template <bool LOCAL>
__global__ void __launch_bounds__(128) k_qlm_average(int N, const int *__restrict__ verlet,
                                                     const int *__restrict__ nn, int M, SboParams P,
                                                     const double *__restrict__ aqr, const double *__restrict__ aqi,
                                                     double *__restrict__ qr, double *__restrict__ qi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int stride = P.ndeg * P.nz;
    int cnt = nn[i];
    if (!P.use_voronoi && P.nnn > 0) cnt = P.nnn;
    int used = 1;
    double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
    double ar[LOCAL ? SBO_LOCAL : 1], ai[LOCAL ? SBO_LOCAL : 1];
    if (LOCAL)
        for (int t = 0; t < stride; ++t) {
            ar[t] = Qr[t];
            ai[t] = Qi[t];
        }
    double *R = LOCAL ? ar : Qr, *I = LOCAL ? ai : Qi;
    for (int jj = 0; jj < cnt; ++jj) {
        const int j = verlet[(size_t)i * M + jj];
        if (j < 0) continue;
        const double *Ar = aqr + (size_t)j * stride, *Ai = aqi + (size_t)j * stride;
        for (int il = 0; il < P.ndeg; ++il) {
            const int mm = 2 * P.l[il] + 1;
            for (int m = 0; m < mm; ++m) {
                R[il * P.nz + m] += __ldg(Ar + il * P.nz + m);
                I[il * P.nz + m] += __ldg(Ai + il * P.nz + m);
            }
        }
        ++used;
    }
    const double inv = 1.0 / used;
    for (int il = 0; il < P.ndeg; ++il) {
        const int mm = 2 * P.l[il] + 1;
        for (int m = 0; m < mm; ++m) {
            Qr[il * P.nz + m] = R[il * P.nz + m] * inv;
            Qi[il * P.nz + m] = I[il * P.nz + m] * inv;
        }
    }
}

// Warp-cooperative form of the same averaging: one warp per atom, lanes over the stride = ndeg * nz
// components of q_lm (real and imaginary rows), neighbours in list order.  Every neighbour row is then ONE
// coalesced read instead of 32 lanes striding through 32 different rows (17 GB -> ~1 GB of DRAM reads per
// 1.5 M atoms); the per-component summation order (own value, then neighbours in list order) is unchanged,
// so the result is bit-identical to k_qlm_average.
__global__ void __launch_bounds__(256) k_qlm_average_warp(int N, const int *__restrict__ verlet,
                                                          const int *__restrict__ nn, int M, int nnn, int use_voronoi,
                                                          int stride, const double *__restrict__ aqr,
                                                          const double *__restrict__ aqi, double *__restrict__ qr,
                                                          double *__restrict__ qi)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps) {
        int cnt = nn[i];
        if (!use_voronoi && nnn > 0) cnt = nnn;
        const int *row = verlet + (size_t)i * M;
        for (int c0 = 0; c0 < stride; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < stride;
            double sr = on ? qr[(size_t)i * stride + c] : 0.0, si = on ? qi[(size_t)i * stride + c] : 0.0;
            int used = 1;
            for (int jj = 0; jj < cnt; ++jj) {
                const int j = __ldg(row + jj);  // same address in every lane: one broadcast load
                if (j < 0) continue;
                if (on) {
                    sr += __ldg(aqr + (size_t)j * stride + c);
                    si += __ldg(aqi + (size_t)j * stride + c);
                }
                ++used;
            }
            const double inv = 1.0 / used;
            if (on) {
                qr[(size_t)i * stride + c] = sr * inv;
                qi[(size_t)i * stride + c] = si * inv;
            }
        }
    }
}

// q_l, w_l, w_l-hat, cpp:506-575
__global__ void __launch_bounds__(128) k_ql_wl(int N, SboParams P, const double *__restrict__ qr,
                                               const double *__restrict__ qi, const double *__restrict__ qnormfac,
                                               const double *__restrict__ cg, const double *__restrict__ sqrt2l1,
                                               int wl, int wlhat, int ncol, double *__restrict__ qn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double EPS = 1e-15;
    const int stride = P.ndeg * P.nz;
    const double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
    double *out = qn + (size_t)i * ncol;
    for (int il = 0; il < P.ndeg; ++il) {
        const int mm = 2 * P.l[il] + 1;
        double s = 0.0;
        for (int m = 0; m < mm; ++m) s += Qr[il * P.nz + m] * Qr[il * P.nz + m] + Qi[il * P.nz + m] * Qi[il * P.nz + m];
        out[il] = qnormfac[il] * sqrt(s);
    }
    if (wl | wlhat) {
        int t = 0;
        for (int il = 0; il < P.ndeg; ++il) {
            const int l = P.l[il];
            const double *R = Qr + il * P.nz, *I = Qi + il * P.nz;
            double ws = 0.0;
            for (int m1 = 0; m1 < 2 * l + 1; ++m1) {
                const int b = max(0, l - m1), e = min(2 * l + 1, 3 * l - m1 + 1);
                for (int m2 = b; m2 < e; ++m2, ++t) {
                    const int m = m1 + m2 - l;
                    const double pr = R[m1] * R[m2] - I[m1] * I[m2];
                    const double pi = R[m1] * I[m2] + I[m1] * R[m2];
                    ws += (pr * R[m] + pi * I[m]) * cg[t];
                }
            }
            const double wf = ws / sqrt2l1[il];
            if (wl) out[il + P.ndeg] = wf;
            if (wlhat) {
                const double qv = out[il];
                if (qv > EPS) {
                    const double q = qnormfac[il] / qv;
                    out[il + (wl ? 1 : 0) * P.ndeg + P.ndeg] = wf * (q * q * q);
                }
            }
        }
    }
}

// first sweep of identifySolidLiquid, cpp:605-644
__global__ void __launch_bounds__(128) k_solid_bonds(int N, const int *__restrict__ verlet,
                                                     const double *__restrict__ dist, const int *__restrict__ nn,
                                                     int M, const double *__restrict__ qr,
                                                     const double *__restrict__ qi, int stride, int off,
                                                     const double *__restrict__ Q6, double threshold, int n_bond,
                                                     int use_voronoi, int nnn, double rc,
                                                     int *__restrict__ solid, int *__restrict__ nbond)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double PI = 3.14159265358979323846;
    int cnt = nn[i];
    if (!use_voronoi && nnn > 0) cnt = nnn;
    const double *ar = qr + (size_t)i * stride + off, *ai = qi + (size_t)i * stride + off;
    int ns = 0;
    for (int jj = 0; jj < cnt; ++jj) {
        const int j = verlet[(size_t)i * M + jj];
        if (j < 0) continue;
        if (dist[(size_t)i * M + jj] > rc) continue;
        const double *br = qr + (size_t)j * stride + off, *bi = qi + (size_t)j * stride + off;
        double s = 0.0;
        for (int m = 0; m < 13; ++m) s += ar[m] * br[m] + ai[m] * bi[m];
        s = s / Q6[i] / Q6[j] * 4 * PI / 13;
        if (s > threshold) ++ns;
    }
    if (ns >= n_bond) solid[i] = 1;
    nbond[i] = ns;
}

// second sweep, cpp:645-674, on a snapshot of the first sweep (the reference reads flags that other
// threads may be clearing; for symmetric lists the outcome is the same)
__global__ void __launch_bounds__(128) k_solid_isolated(int N, const int *__restrict__ verlet,
                                                        const int *__restrict__ nn, int M, int use_voronoi, int nnn,
                                                        const int *__restrict__ snap, int *__restrict__ solid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || snap[i] != 1) return;
    int cnt = nn[i];
    if (!use_voronoi && nnn > 0) cnt = nnn;
    for (int jj = 0; jj < cnt; ++jj) {
        const int j = verlet[(size_t)i * M + jj];
        if (j < 0) continue;
        if (snap[j] == 1) return;
    }
    solid[i] = 0;
}

// n! as the reference tabulates it (15 significant digits, cpp:12-181); exact for n <= 78
double fact15(int n)
{
    long double f = 1.0L;
    for (int i = 2; i <= n; ++i) f *= i;
    char buf[64];
    snprintf(buf, sizeof buf, "%.15Lg", f);
    return strtod(buf, nullptr);
}

__global__ void k_div_small_check(const double *__restrict__ a, int n, int d, unsigned long long *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double q = div_small(a[i], d), r = a[i] / (double)d;
    if (__double_as_longlong(q) != __double_as_longlong(r) && !(q != q && r != r)) atomicAdd(bad, 1ull);
}

void upload_rcp_table(cudaStream_t st)
{
    // constant memory is per device: once per device of this process (several devices: group.cu)
    static std::mutex mu;
    static unsigned long long done = 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done >> (dev & 63) & 1ull) return;
    double h[SBO_MAX_L + 2];
    h[0] = 0.0;
    for (int d = 1; d < SBO_MAX_L + 2; ++d) h[d] = 1.0 / d;
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_rcp, h, sizeof(h), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    done |= 1ull << (dev & 63);
}

}  // namespace

// test hook: number of inputs a[i] whose div_small(a[i], d) differs from a[i] / d (device arrays)
long long sbo_div_small_mismatches(MdbSystem &s, const double *a_dev, int n, int d)
{
    MDB_REQUIRE(d >= 1 && d <= SBO_MAX_L + 1, MDB_ERR_VALUE, "divisor out of range");
    upload_rcp_table(s.stream);
    unsigned long long *bad = s.scratch2.ensure<unsigned long long>(1);
    CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), s.stream));
    MDB_LAUNCH(k_div_small_check, (n + 255) / 256, 256, 0, s.stream, a_dev, n, d, bad);
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return (long long)h;
}

void launch_steinhardt(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M,
                       const double *weight, const int *llist, int ndeg, int nnn, int lmax, bool wl, bool wlhat,
                       bool average, bool use_voronoi, double rc, bool use_weight, double *qr, double *qi, double *qn)
{
    MDB_REQUIRE(ndeg >= 1 && ndeg <= 8, MDB_ERR_VALUE, "1 to 8 degrees supported, got %d", ndeg);
    MDB_REQUIRE(lmax >= 0 && lmax <= SBO_MAX_L, MDB_ERR_VALUE, "lmax=%d exceeds the device limit %d", lmax, SBO_MAX_L);
    const double PI = 3.14159265358979323846;
    SboParams P{};
    P.ndeg = ndeg;
    P.lmax = lmax;
    P.nz = 2 * lmax + 1;
    P.nnn = nnn;
    P.use_voronoi = use_voronoi;
    P.use_weight = use_weight;
    P.rc = rc;
    double norm[8][SBO_MAX_L + 1] = {};
    std::vector<double> qnf(ndeg), s2l1(ndeg);
    for (int il = 0; il < ndeg; ++il) {
        const int l = llist[il];
        MDB_REQUIRE(l >= 0 && l <= lmax, MDB_ERR_VALUE, "degree %d outside [0, lmax=%d]", l, lmax);
        P.l[il] = l;
        P.off[il + 1] = P.off[il] + 2 * l + 1;
        for (int m = 0; m <= l; ++m) {  // _polar_prefactor, cpp:270-286
            double pf = 1.0;
            for (int i = l - m + 1; i < l + m + 1; ++i) pf *= i;
            norm[il][m] = std::sqrt((2 * l + 1) / (4 * PI * pf));
        }
        qnf[il] = std::sqrt(4 * PI / (2 * l + 1));
        s2l1[il] = std::sqrt(2 * l + 1.0);
    }
    cudaStream_t st = s.stream;
    upload_rcp_table(st);
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_norm, norm, sizeof(norm), 0, cudaMemcpyHostToDevice, st));
    const int N = s.n_rows;
    const int nb = (N + 127) / 128;
    {
        // accumulators in shared memory when the compact slots of a 128-thread block fit (l = {4, 6}: 45 KB)
        const size_t smem = (size_t)2 * P.off[ndeg] * 128 * sizeof(double);
        // per launch: the attribute belongs to the current device (several devices per process: group.cu)
        CUDA_TRY(cudaFuncSetAttribute(k_qlm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        const char *env = getenv("MDB_SBO");
        if (smem <= 100 * 1024 && !(env && !strcmp(env, "local")))
            MDB_LAUNCH(k_qlm<2>, nb, 128, smem, st, s.x, s.y, s.z, N, s.box, verlet, dist, nn, weight, M, P, qr, qi);
        else if (P.ndeg * P.nz <= SBO_LOCAL)
            MDB_LAUNCH(k_qlm<1>, nb, 128, 0, st, s.x, s.y, s.z, N, s.box, verlet, dist, nn, weight, M, P, qr, qi);
        else
            MDB_LAUNCH(k_qlm<0>, nb, 128, 0, st, s.x, s.y, s.z, N, s.box, verlet, dist, nn, weight, M, P, qr, qi);
    }
    if (average) {
        // the snapshot is indexed by neighbour ids: it covers every local atom (qr/qi hold s.N rows, the
        // ghost rows of a decomposed frame are zero)
        const size_t tot = (size_t)s.N * P.ndeg * P.nz;
        double *ar = s.scratch.ensure<double>(tot), *ai = s.scratch2.ensure<double>(tot);
        CUDA_TRY(cudaMemcpyAsync(ar, qr, sizeof(double) * tot, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ai, qi, sizeof(double) * tot, cudaMemcpyDeviceToDevice, st));
        const char *env = getenv("MDB_SBO");
        if (env && !strcmp(env, "local")) {
            if (P.ndeg * P.nz <= SBO_LOCAL) MDB_LAUNCH(k_qlm_average<true>, nb, 128, 0, st, N, verlet, nn, M, P, ar, ai, qr, qi);
            else MDB_LAUNCH(k_qlm_average<false>, nb, 128, 0, st, N, verlet, nn, M, P, ar, ai, qr, qi);
        } else {
            const int nbw = (N + 7) / 8 < 148 * 8 * 4 ? (N + 7) / 8 : 148 * 8 * 4;
            MDB_LAUNCH(k_qlm_average_warp, nbw, 256, 0, st, N, verlet, nn, M, P.nnn, P.use_voronoi, P.ndeg * P.nz, ar, ai,
                       qr, qi);
        }
    }
    // Clebsch-Gordan table, cpp:188-224
    std::vector<double> cg(1, 0.0);
    if (wl || wlhat) {
        cg.clear();
        for (int il = 0; il < ndeg; ++il) {
            const int l = llist[il];
            for (int m1 = 0; m1 < 2 * l + 1; ++m1) {
                const int aa = m1 - l;
                for (int m2 = std::max(0, l - m1); m2 < std::min(2 * l + 1, 3 * l - m1 + 1); ++m2) {
                    const int bb = m2 - l, m = aa + bb + l;
                    double sums = 0.0;
                    for (int zz = std::max(0, std::max(-aa, bb)); zz < std::min(l, std::min(l - aa, l + bb)) + 1; ++zz) {
                        const int ifac = (zz % 2) ? -1 : 1;
                        sums += ifac / (fact15(zz) * fact15(l - zz) * fact15(l - aa - zz) * fact15(l + bb - zz) *
                                        fact15(aa + zz) * fact15(-bb + zz));
                    }
                    const int cc = m - l;
                    const double sfaccg = std::sqrt(fact15(l + aa) * fact15(l - aa) * fact15(l + bb) * fact15(l - bb) *
                                                    fact15(l + cc) * fact15(l - cc) * (2 * l + 1));
                    const double sfac1 = fact15(3 * l + 1), sfac2 = fact15(l);
                    const double dcg = std::sqrt(sfac2 * sfac2 * sfac2 / sfac1);
                    cg.push_back(sums * dcg * sfaccg);
                }
            }
        }
    }
    const size_t ncg = cg.size();
    double *dtab = s.out_f64c.ensure<double>(ncg + 2 * ndeg);
    std::vector<double> tab(cg);
    tab.insert(tab.end(), qnf.begin(), qnf.end());
    tab.insert(tab.end(), s2l1.begin(), s2l1.end());
    CUDA_TRY(cudaMemcpyAsync(dtab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // tab is a host temporary
    const int ncol = ndeg * (1 + (wl ? 1 : 0) + (wlhat ? 1 : 0));
    MDB_LAUNCH(k_ql_wl, nb, 128, 0, st, N, P, qr, qi, dtab + ncg, dtab, dtab + ncg + ndeg, wl ? 1 : 0, wlhat ? 1 : 0,
               ncol, qn);
    CUDA_TRY(cudaGetLastError());
}

void launch_solid_liquid(MdbSystem &s, const int *verlet, const double *dist, const int *nn, int M, int q6index,
                         const double *Q6, const double *qr, const double *qi, int ndeg, int nz, double threshold,
                         int n_bond, bool use_voronoi, int nnn, double rc, int *solid, int *nbond)
{
    MDB_REQUIRE(nz >= 13, MDB_ERR_VALUE, "q_6m needs 2*lmax+1 >= 13 columns, got %d", nz);
    const int N = s.n_rows;
    const int nb = (N + 127) / 128;
    cudaStream_t st = s.stream;
    MDB_LAUNCH(k_solid_bonds, nb, 128, 0, st, N, verlet, dist, nn, M, qr, qi, ndeg * nz, q6index * nz, Q6, threshold,
               n_bond, use_voronoi ? 1 : 0, nnn, rc, solid, nbond);
    int *snap = s.scratch.ensure<int>(s.N);  // indexed by neighbour ids (solid holds s.N entries)
    CUDA_TRY(cudaMemcpyAsync(snap, solid, sizeof(int) * s.N, cudaMemcpyDeviceToDevice, st));
    MDB_LAUNCH(k_solid_isolated, nb, 128, 0, st, N, verlet, nn, M, use_voronoi ? 1 : 0, nnn, snap, solid);
    CUDA_TRY(cudaGetLastError());
}
