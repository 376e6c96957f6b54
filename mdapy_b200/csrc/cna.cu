// mdapy_b200/csrc/cna.cu
//
// Common neighbour analysis on the device.  Replaces src/cna.cpp:429-506
// (FixedCNA) and 289-427 (AdaptiveCNA).  Per atom: a bond bit-matrix among
// its 12 or 14 neighbours (bond iff min-image d2 <= cutoff^2 on RAW
// coordinates, cna.cpp:149-161), then for every neighbour the triplet
// (#common neighbours, #bonds among them, longest bond chain) is classified
// into 421/422/555/444/666 counts.
//
// "Longest chain" in the reference (cna.cpp:97-147) is the number of bonds in
// the largest connected component of the common-neighbour bond graph; it is
// order independent, so here it is a bit-parallel flood over adjacency masks.
#include "internal.cuh"
#include "cna_core.cuh"

namespace {

// nb[v]: bit u set iff neighbours v and u are bonded; nn <= 14
__device__ __forceinline__ CnaCounts cna_signatures(const unsigned *nb, int nn)
{
    CnaCounts c{0, 0, 0, 0, 0};
    for (int ni = 0; ni < nn; ++ni) {
        const unsigned common = nb[ni];
        const int ncommon = __popc(common);
        // bonds among the common neighbours
        int twice_bonds = 0;
        for (unsigned m = common; m; m &= m - 1) {
            const int v = __ffs(m) - 1;
            twice_bonds += __popc(nb[v] & common);
        }
        const int nbonds = twice_bonds >> 1;
        // largest connected component, measured in bonds
        int longest = 0;
        unsigned remaining = common;
        while (remaining) {
            const int v0 = __ffs(remaining) - 1;
            unsigned comp = 1u << v0, frontier = comp;
            while (frontier) {
                unsigned next = 0;
                for (unsigned m = frontier; m; m &= m - 1) next |= nb[__ffs(m) - 1] & common;
                next &= ~comp;
                comp |= next;
                frontier = next;
            }
            int e2 = 0;
            for (unsigned m = comp; m; m &= m - 1) e2 += __popc(nb[__ffs(m) - 1] & common);
            longest = max(longest, e2 >> 1);
            remaining &= ~comp;
        }
        if (ncommon == 4 && nbonds == 2) {
            if (longest == 1)
                ++c.n421;
            else if (longest == 2)
                ++c.n422;
        } else if (ncommon == 5 && nbonds == 5 && longest == 5)
            ++c.n555;
        else if (ncommon == 4 && nbonds == 4 && longest == 4)
            ++c.n444;
        else if (ncommon == 6 && nbonds == 6 && longest == 6)
            ++c.n666;
    }
    return c;
}

__device__ __forceinline__ void bond_matrix(const DBox &box, const double *px, const double *py, const double *pz,
                                            int nn, double cutsq, unsigned *nb)
{
    for (int a = 0; a < nn; ++a) nb[a] = 0;
    for (int a = 0; a < nn; ++a)
        for (int b = a + 1; b < nn; ++b) {
            const double d2 = pbc_dist_sq(box, px[a], py[a], pz[a], px[b], py[b], pz[b]);
            if (d2 <= cutsq) {
                nb[a] |= 1u << b;
                nb[b] |= 1u << a;
            }
        }
}

// cna.cpp:429-506.  pattern must be pre-zeroed by the caller (only non-zero labels are written).
__global__ void __launch_bounds__(128) k_fcna(const double *__restrict__ x, const double *__restrict__ y,
                                              const double *__restrict__ z, int N, DBox box,
                                              const int *__restrict__ verlet, const int *__restrict__ nnum, int M,
                                              double cutsq, int *__restrict__ pattern, int first)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;  // rows [first, N)
    if (i >= N) return;
    const int nn = nnum[i];
    if ((nn != 12 && nn != 14) || nn > M) return;
    double px[14], py[14], pz[14];
    const int *row = verlet + (size_t)i * M;
    for (int a = 0; a < nn; ++a) {
        const int j = row[a];
        px[a] = x[j];
        py[a] = y[j];
        pz[a] = z[j];
    }
    unsigned nb[14];
    bond_matrix(box, px, py, pz, nn, cutsq, nb);
    const CnaCounts c = cna_signatures(nb, nn);
    int p = 0;
    if (c.n421 == 12)
        p = 1;
    else if (c.n421 == 6 && c.n422 == 6)
        p = 2;
    else if (c.n555 == 12)
        p = 4;
    else if (c.n666 == 8 && c.n444 == 6)
        p = 3;
    if (p) pattern[i] = p;
}

// cna.cpp:289-427.  verlet rows hold >= 14 neighbours sorted by distance.
__global__ void __launch_bounds__(128) k_acna(const double *__restrict__ x, const double *__restrict__ y,
                                              const double *__restrict__ z, int N, DBox box,
                                              const int *__restrict__ verlet, int M, double one_plus_sqrt2,
                                              int *__restrict__ pattern)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double xi = x[i], yi = y[i], zi = z[i];
    double px[14], py[14], pz[14];
    const int *row = verlet + (size_t)i * M;
    for (int a = 0; a < 14; ++a) {
        const int j = row[a];
        px[a] = x[j];
        py[a] = y[j];
        pz[a] = z[j];
    }
    unsigned nb[14];
    // 12-neighbour pass (cna.cpp:312-370)
    double rsum = 0.0;
    for (int m = 0; m < 12; ++m) rsum += sqrt(pbc_dist_sq(box, xi, yi, zi, px[m], py[m], pz[m]));
    double cut = rsum / 12 * one_plus_sqrt2 * 0.5;
    bond_matrix(box, px, py, pz, 12, cut * cut, nb);
    CnaCounts c = cna_signatures(nb, 12);
    int p = 0;
    if (c.n421 == 12)
        p = 1;
    else if (c.n421 == 6 && c.n422 == 6)
        p = 2;
    else if (c.n555 == 12)
        p = 4;
    if (p == 0) {
        // 14-neighbour BCC pass (cna.cpp:372-425)
        rsum = 0.0;
        for (int m = 0; m < 8; ++m) rsum += sqrt(pbc_dist_sq(box, xi, yi, zi, px[m], py[m], pz[m]) / (3.0 / 4.0));
        for (int m = 8; m < 14; ++m) rsum += sqrt(pbc_dist_sq(box, xi, yi, zi, px[m], py[m], pz[m]));
        cut = rsum / 14 * one_plus_sqrt2 * 0.5;
        bond_matrix(box, px, py, pz, 14, cut * cut, nb);
        c = cna_signatures(nb, 14);
        if (c.n666 == 8 && c.n444 == 6) p = 3;
    }
    if (p) pattern[i] = p;
}


// ---------------------------------------------------------------------------------------------
// Fast fixed-cutoff CNA for orthogonal frames whose periodic edges exceed 4.2 * rc.
// Neighbour vectors r_a = min_image(x_a - x_i) are formed exactly in f64, rounded to fp32, and the
// 66 / 91 bond tests run on |r_a - r_b|^2 in fp32.  With every periodic edge > 4 rc the difference of
// two nearest-image vectors IS the minimum image of x_b - x_a, so the fp32 value only differs from the
// reference's f64 d2 by rounding (< 1e-5 rc^2); a pair whose fp32 d2 lies within 1e-4 rc^2 of the
// threshold is re-evaluated with the reference's exact expression on raw coordinates
// (cna.cpp:149-161), so the bond matrix -- and therefore every label -- is identical.
// Bond rows live in shared memory (interleaved per thread) because the signature code indexes them
// dynamically.
constexpr int CNA_THREADS = 128;

// slow path of the fast kernel: every bond from the reference's exact expression (cna.cpp:149-161)
__device__ __noinline__ int fcna_exact_atom(const double *__restrict__ x, const double *__restrict__ y,
                                            const double *__restrict__ z, const DBox &box,
                                            const int *__restrict__ row, int nn, double cutsq, unsigned short *nb)
{
    for (int a = 0; a < nn; ++a) nb[a * CNA_THREADS] = 0;
    for (int a = 0; a < nn; ++a) {
        const int ja = row[a];
        const double xa = x[ja], ya = y[ja], za = z[ja];
        for (int b = a + 1; b < nn; ++b) {
            const int jb = row[b];
            if (pbc_dist_sq(box, xa, ya, za, x[jb], y[jb], z[jb]) <= cutsq) {
                nb[a * CNA_THREADS] |= (unsigned short)(1u << b);
                nb[b * CNA_THREADS] |= (unsigned short)(1u << a);
            }
        }
    }
    return cna_label(cna_signatures_smem<CNA_THREADS>(nb, nn));
}

template <int NN>
__device__ __forceinline__ int fcna_fast_body(const double *__restrict__ x, const double *__restrict__ y,
                                              const double *__restrict__ z, const DBox &box, int i,
                                              const int *__restrict__ row, double cutsq, float cut_lo, float cut_hi,
                                              unsigned short *nb)
{
    const double xi = x[i], yi = y[i], zi = z[i];
    const double Lx = box.h[0], Ly = box.h[4], Lz = box.h[8];
    const double iLx = box.hinv[0], iLy = box.hinv[4], iLz = box.hinv[8];
    float rx[NN], ry[NN], rz[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const int j = __ldg(row + a);
        double dx = __ldg(x + j) - xi, dy = __ldg(y + j) - yi, dz = __ldg(z + j) - zi;
        // listed neighbours lie within rc << L/2 of atom i: nearest integer of d/L is the image count
        // (only fp32 bond screening uses these vectors; anything near the threshold is redone exactly)
        if (box.pbc[0]) dx -= Lx * rint(dx * iLx);
        if (box.pbc[1]) dy -= Ly * rint(dy * iLy);
        if (box.pbc[2]) dz -= Lz * rint(dz * iLz);
        rx[a] = (float)dx;
        ry[a] = (float)dy;
        rz[a] = (float)dz;
    }
    unsigned rows[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) rows[a] = 0;
    bool ambiguous = false;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
#pragma unroll
        for (int b = a + 1; b < NN; ++b) {
            const float dx = rx[b] - rx[a], dy = ry[b] - ry[a], dz = rz[b] - rz[a];
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            if (d2 < cut_lo) {
                rows[a] |= 1u << b;
                rows[b] |= 1u << a;
            } else if (d2 <= cut_hi)
                ambiguous = true;
        }
    }
    if (ambiguous) return fcna_exact_atom(x, y, z, box, row, NN, cutsq, nb);
    return cna_label(cna_signatures_regs<NN, CNA_THREADS>(rows, nb));
}

__global__ void __launch_bounds__(CNA_THREADS) k_fcna_fast(const double *__restrict__ x, const double *__restrict__ y,
                                                           const double *__restrict__ z, int N,
                                                           const __grid_constant__ DBox box,
                                                           const int *__restrict__ verlet,
                                                           const int *__restrict__ nnum, int M, double cutsq,
                                                           float cut_lo, float cut_hi, int *__restrict__ pattern,
                                                           int first)
{
    __shared__ unsigned short nb_s[14 * CNA_THREADS];
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;  // rows [first, N)
    if (i >= N) return;
    const int nn = nnum[i];
    if ((nn != 12 && nn != 14) || nn > M) return;
    const int *row = verlet + (size_t)i * M;
    unsigned short *nb = nb_s + threadIdx.x;
    const int p = nn == 12 ? fcna_fast_body<12>(x, y, z, box, i, row, cutsq, cut_lo, cut_hi, nb)
                           : fcna_fast_body<14>(x, y, z, box, i, row, cutsq, cut_lo, cut_hi, nb);
    if (p) pattern[i] = p;
}

// ---------------------------------------------------------------------------------------------
// Diamond structure identification, cna.cpp:163-287.  verlet rows: >= 4 neighbours, ascending distance.
// k_ids_classify = the reference's parallel loop (second-shell list, CNA over 12 with the local
// cut-off).  The two serial sweeps that follow in the reference (cna.cpp:251-286) visit atoms in
// ascending index and label still-unlabelled first neighbours, so the label an atom receives is the
// one of the LOWEST-index labelled atom that lists it: k_ids_claim records that index with atomicMin,
// k_ids_assign applies it -- deterministic and identical to the serial result.
__global__ void __launch_bounds__(128) k_ids_classify(const double *__restrict__ x, const double *__restrict__ y,
                                                      const double *__restrict__ z, int N, int NA, DBox box,
                                                      const int *__restrict__ verlet, int M,
                                                      int *__restrict__ second_out, int *__restrict__ pattern)
{
    // N rows are classified; neighbour ids are valid below NA (NA > N: ghosts of a decomposed frame,
    // whose k-nearest rows exist as well)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int second[12];
    int count = 0;
    const int *row = verlet + (size_t)i * M;
    bool ok = true;
    for (int m = 0; m < 4; ++m) {
        const int j = row[m];
        if (j < 0 || j >= NA) {
            ok = false;
            break;
        }
        const int *rj = verlet + (size_t)j * M;
        int taken = 0;
        for (int kk = 0; kk < 4; ++kk) {
            const int k = rj[kk];
            if (k != i && taken < 3) {
                second[count++] = k;
                ++taken;
            }
        }
    }
    for (int m = 0; m < 12 && ok; ++m)
        if (m >= count || second[m] < 0 || second[m] >= NA) ok = false;
    if (second_out)
        for (int m = 0; m < 12; ++m) second_out[(size_t)i * 12 + m] = m < count ? second[m] : 0;
    if (!ok) return;  // short / padded rows: the reference would read out of bounds; leave 0
    const double xi = x[i], yi = y[i], zi = z[i];
    double px[12], py[12], pz[12];
    double rsum = 0.0;
    for (int m = 0; m < 12; ++m) {
        const int j = second[m];
        px[m] = x[j];
        py[m] = y[j];
        pz[m] = z[j];
        rsum += sqrt(pbc_dist_sq(box, xi, yi, zi, px[m], py[m], pz[m]));
    }
    rsum /= 12.0;
    const double cut = rsum * 1.2071068;
    unsigned nb[12];
    bond_matrix(box, px, py, pz, 12, cut * cut, nb);
    const CnaCounts c = cna_signatures(nb, 12);
    if (c.n421 == 12) pattern[i] = 1;
    else if (c.n421 == 6 && c.n422 == 6) pattern[i] = 4;
}

// owner[j] = (order key of the claiming atom << 32) | its local index; the order key is the GLOBAL id in
// a decomposed frame (the reference's sweeps run in ascending original index), the index itself otherwise
__global__ void k_ids_claim(int N, const int *__restrict__ verlet, int M, const int *__restrict__ pattern, int ta, int tb,
                            const int *__restrict__ gid, unsigned long long *__restrict__ owner)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int t = pattern[i];
    if (t != ta && t != tb) return;
    const unsigned long long key = ((unsigned long long)(unsigned)(gid ? gid[i] : i) << 32) | (unsigned)i;
    for (int jj = 0; jj < 4; ++jj) {
        const int j = verlet[(size_t)i * M + jj];
        if (j >= 0 && j < N && pattern[j] == 0) atomicMin(owner + j, key);
    }
}

__global__ void k_ids_assign(int N, int *__restrict__ pattern, unsigned long long *__restrict__ owner)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const unsigned long long o = owner[j];
    if (o != ~0ull) {
        pattern[j] = pattern[(int)(o & 0xffffffffu)] + 1;  // owners carry 1/4 (or 2/5), receivers were 0: no read/write overlap
        owner[j] = ~0ull;
    }
}

__global__ void k_fill_u64(int n, unsigned long long v, unsigned long long *__restrict__ p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_fill_int(int n, int v, int *__restrict__ p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// rows [first, first + count) (count < 0: all rows)
void launch_fcna(MdbSystem &s, const int *verlet, const int *nn, int M, double rc, int *pattern, int first, int count)
{
    const int N = count < 0 ? s.n_rows : first + count;
    const int span = N - first;
    if (span <= 0) return;
    // fast path: orthogonal frame, the list's own cut-off bounds the neighbour distances (cut-off list
    // built with list_rc >= every stored distance) and every periodic edge exceeds 4.2 * that bound
    bool fast = !s.box.triclinic && s.list_kind == LIST_CUTOFF && s.list_rc > 0;
    const double rb = s.list_rc > rc ? s.list_rc : rc;
    for (int d = 0; d < 3 && fast; ++d)
        if (s.box.pbc[d] && !(s.box.h[4 * d] > 4.2 * rb)) fast = false;
    const char *env = getenv("MDB_CNA");
    if (env && !strcmp(env, "exact")) fast = false;
    if (fast) {
        const double c2 = rc * rc;
        MDB_LAUNCH(k_fcna_fast, (span + CNA_THREADS - 1) / CNA_THREADS, CNA_THREADS, 0, s.stream, s.x, s.y, s.z, N, s.box,
                   verlet, nn, M, c2, (float)(c2 * (1.0 - 1e-4)), (float)(c2 * (1.0 + 1e-4)), pattern, first);
        CUDA_TRY(cudaGetLastError());
        return;
    }
    MDB_LAUNCH(k_fcna, (span + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, N, s.box, verlet, nn, M, rc * rc, pattern,
               first);
    CUDA_TRY(cudaGetLastError());
}

void launch_acna(MdbSystem &s, const int *verlet, int M, int *pattern)
{
    const int N = s.n_rows;
    MDB_REQUIRE(M >= 14, MDB_ERR_VALUE, "adaptive CNA needs >= 14 sorted neighbours per atom, row width is %d", M);
    const double f = 1.0 + std::sqrt(2.0);
    MDB_LAUNCH(k_acna, (N + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, N, s.box, verlet, M, f, pattern);
    CUDA_TRY(cudaGetLastError());
}

// pattern must be pre-zeroed; second_out (N x 12, may be null) receives the reference's new_verlet_list
void launch_ids(MdbSystem &s, const int *verlet, int M, int *second_out, int *pattern)
{
    const int N = s.n_rows;
    MDB_REQUIRE(M >= 4, MDB_ERR_VALUE, "diamond identification needs >= 4 sorted neighbours per atom, row width is %d", M);
    const int nb = (N + 127) / 128;
    MDB_LAUNCH(k_ids_classify, nb, 128, 0, s.stream, s.x, s.y, s.z, N, s.N, s.box, verlet, M, second_out, pattern);
    unsigned long long *owner = s.scratch.ensure<unsigned long long>(N);
    MDB_LAUNCH(k_fill_u64, (N + 255) / 256, 256, 0, s.stream, N, ~0ull, owner);
    for (int pass = 0; pass < 2; ++pass) {
        MDB_LAUNCH(k_ids_claim, (N + 255) / 256, 256, 0, s.stream, N, verlet, M, pattern, pass ? 2 : 1, pass ? 5 : 4, s.gid,
                   owner);
        MDB_LAUNCH(k_ids_assign, (N + 255) / 256, 256, 0, s.stream, N, pattern, owner);
    }
    CUDA_TRY(cudaGetLastError());
}
