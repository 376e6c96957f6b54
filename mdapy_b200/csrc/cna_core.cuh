// mdapy_b200/csrc/cna_core.cuh -- CNA signature counting shared by the list-based kernels (cna.cu) and the
// fused neighbour + CNA kernel (neighbor_tiled.cu).  src/cna.cpp:16-161, 429-506.
#pragma once

struct CnaCounts {
    int n421, n422, n555, n444, n666;
};

__device__ __forceinline__ int cna_label(const CnaCounts &c)
{
    if (c.n421 == 12) return 1;
    if (c.n421 == 6 && c.n422 == 6) return 2;
    if (c.n555 == 12) return 4;
    if (c.n666 == 8 && c.n444 == 6) return 3;
    return 0;
}

// Signature counts from bond rows in shared memory (nb[v * STRIDE] = row of neighbour v).
// The reference's "longest chain" (cna.cpp:97-147) is the bond count of the largest connected
// component of the common-neighbour bond graph.  For the signatures that are counted it reduces to:
//   (4 common, 2 bonds): chain 2 iff some common neighbour has degree 2, else 1 (two disjoint bonds);
//   (5,5) and (4,4):     the bonds are necessarily connected (two bond-carrying components on 5 / 4
//                        vertices hold at most 4 / 2 bonds), so the chain is 5 / 4;
//   (6,6):               connected unless the bonds split 3+3, 4+2 or 5+1: one flood decides.
template <int STRIDE>
__device__ __forceinline__ CnaCounts cna_signatures_smem(const unsigned short *nb, int nn)
{
    CnaCounts c{0, 0, 0, 0, 0};
#pragma unroll 1
    for (int ni = 0; ni < nn; ++ni) {
        const unsigned common = nb[ni * STRIDE];
        const int ncommon = __popc(common);
        if (ncommon < 4 || ncommon > 6) continue;
        int twice = 0, maxdeg = 0;
        unsigned first_row = 0;
        int first_v = -1;
        for (unsigned m = common; m; m &= m - 1) {
            const int v = __ffs(m) - 1;
            const unsigned r = nb[v * STRIDE] & common;
            const int d = __popc(r);
            twice += d;
            maxdeg = max(maxdeg, d);
            if (first_v < 0 && d > 0) {
                first_v = v;
                first_row = r;
            }
        }
        const int nbonds = twice >> 1;
        if (ncommon == 4) {
            if (nbonds == 2) {
                if (maxdeg == 2) ++c.n422;
                else ++c.n421;
            } else if (nbonds == 4)
                ++c.n444;
        } else if (ncommon == 5) {
            if (nbonds == 5) ++c.n555;
        } else if (nbonds == 6) {
            // flood the component of the first bonded vertex; all 6 bonds must lie inside it
            unsigned comp = (1u << first_v) | first_row, frontier = first_row;
            while (frontier) {
                unsigned next = 0;
                for (unsigned m = frontier; m; m &= m - 1) next |= nb[(__ffs(m) - 1) * STRIDE] & common;
                next &= ~comp;
                comp |= next;
                frontier = next;
            }
            int e2 = 0;
            for (unsigned m = comp; m; m &= m - 1) e2 += __popc(nb[(__ffs(m) - 1) * STRIDE] & common);
            if (e2 == 12) ++c.n666;
        }
    }
    return c;
}

// Register form of the same signature counts for the fast kernel (rows[] in registers, NN static).
// For a bonded pair (a, b) the number of common neighbours d = popc(rows[a] & rows[b]) is at the same time the
// degree of b inside the common-neighbour graph of a AND of a inside that of b, so one pass over the
// NN(NN-1)/2 pairs with static indices yields, for every neighbour ni, twice the bond count and the largest
// degree among its common neighbours -- half the popcounts of the per-neighbour loop and no shared-memory
// traffic.  Only the (6 common, 6 bonds) case still needs the flood (BCC), done on the shared-memory copy.
template <int NN, int STRIDE>
__device__ __forceinline__ CnaCounts cna_signatures_regs(const unsigned (&rows)[NN], unsigned short *nb)
{
    int tw[NN], mx[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        tw[a] = 0;
        mx[a] = 0;
    }
#pragma unroll
    for (int a = 0; a < NN; ++a) {
#pragma unroll
        for (int b = a + 1; b < NN; ++b) {
            const int d = (rows[a] >> b) & 1u ? __popc(rows[a] & rows[b]) : 0;
            tw[a] += d;
            tw[b] += d;
            mx[a] = max(mx[a], d);
            mx[b] = max(mx[b], d);
        }
    }
    CnaCounts c{0, 0, 0, 0, 0};
    bool need_flood = false;
#pragma unroll
    for (int ni = 0; ni < NN; ++ni) {
        const int ncommon = __popc(rows[ni]);
        const int nbonds = tw[ni] >> 1;
        if (ncommon == 4) {
            if (nbonds == 2) {
                if (mx[ni] == 2) ++c.n422;
                else ++c.n421;
            } else if (nbonds == 4)
                ++c.n444;
        } else if (ncommon == 5) {
            if (nbonds == 5) ++c.n555;
        } else if (ncommon == 6 && nbonds == 6)
            need_flood = true;
    }
    if (need_flood) {
#pragma unroll
        for (int a = 0; a < NN; ++a) nb[a * STRIDE] = (unsigned short)rows[a];
#pragma unroll 1
        for (int ni = 0; ni < NN; ++ni) {
            const unsigned common = nb[ni * STRIDE];
            if (__popc(common) != 6) continue;
            int twice = 0, first_v = -1;
            unsigned first_row = 0;
            for (unsigned m = common; m; m &= m - 1) {
                const int v = __ffs(m) - 1;
                const unsigned r = nb[v * STRIDE] & common;
                const int d = __popc(r);
                twice += d;
                if (first_v < 0 && d > 0) {
                    first_v = v;
                    first_row = r;
                }
            }
            if (twice != 12) continue;
            // flood the component of the first bonded vertex; all 6 bonds must lie inside it
            unsigned comp = (1u << first_v) | first_row, frontier = first_row;
            while (frontier) {
                unsigned next = 0;
                for (unsigned m = frontier; m; m &= m - 1) next |= nb[(__ffs(m) - 1) * STRIDE] & common;
                next &= ~comp;
                comp |= next;
                frontier = next;
            }
            int e2 = 0;
            for (unsigned m = comp; m; m &= m - 1) e2 += __popc(nb[(__ffs(m) - 1) * STRIDE] & common);
            if (e2 == 12) ++c.n666;
        }
    }
    return c;
}


