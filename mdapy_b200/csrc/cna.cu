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

namespace {

struct CnaCounts {
    int n421, n422, n555, n444, n666;
};

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
                                              double cutsq, int *__restrict__ pattern)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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

}  // namespace

void launch_fcna(MdbSystem &s, const int *verlet, const int *nn, int M, double rc, int *pattern)
{
    const int N = s.n_rows;
    MDB_LAUNCH(k_fcna, (N + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, N, s.box, verlet, nn, M, rc * rc, pattern);
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
