// mdapy_b200/csrc/neighbor_tiled.cu
//
// Cell-tile neighbour build: the production kernel for orthogonal boxes (src/neighbor.cpp:102-187
// semantics, rows bit-identical to k_neighbor_direct and to the reference, including row order).
//
// One CTA owns a T x T x T block of cells.  The sorted records of the (T+2)^3 surrounding cells are
// staged in shared memory pencil by pencil -- every z-pencil of cells is one contiguous run of
// 32-byte SortedAtom records, copied with cp.async.bulk (TMA 1-D, mbarrier completion).  Each staged
// atom also gets an fp32 copy of its position relative to the tile centre (nearest periodic image).
// One thread per owned atom then walks its 27 cells in the reference's order:
//   phase 1  fp32 test |rj|^2 - 2 ri.rj <= rc^2 (1 + guard) - |ri|^2 (three FMAs per candidate): a
//            conservative pre-filter (no false negatives: the guard is > 5x the fp32 error bound for
//            staged positions inside the radius limit; a tile holding an atom beyond the limit takes
//            the exact direct path), survivors are queued per thread;
//   phase 2  the queued survivors take the exact f64 test of the reference (xi wrapped, x[j] raw,
//            division-free min-image, left-to-right sum, <= rc^2) and are written in queue order.
// Only ~20 % of the 27-cell candidates survive phase 1, so the f64 pipe sees ~13 pairs per atom
// instead of ~67.  Rows are padded by the writing thread (-1 / rc+1), so no prefill pass is needed.
#include "internal.cuh"
#include "cna_core.cuh"
#include <type_traits>

namespace {

constexpr int TILE_THREADS = 256;
constexpr int SURV_CAP = 20;      // per-thread survivor queue (uint16 staged indices)
constexpr int SURV_RESERVE = 10;  // warp drains when any lane holds more than this at a pencil boundary

struct TileArgs {
    const SortedAtom *sorted;
    const int *cell_start;
    DBox box;
    CellGrid g;
    double rcsq;
    double pad;     // rc + 1.0
    float rcsq_hi;  // fp32 acceptance bound of the pre-filter
    float w_limit;  // largest |r|^2 (relative to the tile centre) the fp32 error bound was derived for
    int M;
    int n_rows;
    int *verlet;
    double *dist;
    int *nn;
    int *max_count;  // atomicMax of the true neighbour count
    int *min_count;  // atomicMin of the true neighbour count (tells uniform frames from disordered ones)
    int cap;         // staged-atom capacity of the shared buffers
    int p_lo, p_hi;  // owned range of stored x planes
    int wrap_x;      // 1: planes wrap around (single GPU), 0: slab window with ghost planes
    int tiles_y, tiles_z;
    int tile_stride, tile_offset, n_tiles;  // sampling of the tile list (count-only estimate)
    int count_only;
    int use_tma;
    int tiles_x;     // tiles along x (warp-per-cell kernel: strip order of the tile list)
    int ocap;        // owned-atom capacity of the per-tile wrapped-position table
    int strip;       // floor measurement: staging + row stores only (MDB_STRIP=1)
    int *pattern;    // fused neighbour + CNA kernel: labels out
    float rcsq_lo;   // fp32 bound below which a candidate is a neighbour without the exact test
    float cut_lo, cut_hi;   // fp32 band of the bond test among neighbours
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA 1-D); bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_rec(unsigned addr, double &x, double &y, double &z, int &idx)
{
    double w;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(z), "=d"(w) : "r"(addr + 16));
    idx = __double2loint(w);
}

// Exact min-image for a pair that already passed the fp32 pre-filter on a frame with >= 7 cells per
// periodic axis: the nearest-image displacement is < 0.15 L, so n = floor(x/L + 0.5) of the reference
// (src/box.h:119-124) is simply the integer nearest to x/L -- far from every rounding boundary, hence
// rint(x * (1/L)) yields the same integer -- and x - L*n is evaluated with the reference's two roundings.
__device__ __forceinline__ double near_image(double x, double L, double invL)
{
    const double n = rint(x * invL);
    return x - L * n;
}

// Tile centre and centre-relative nearest-image position of a staged atom, orthogonal and triclinic boxes.
// Cells are (cx, cy, cz) in CELL units along the three lattice directions; a cell is rc / thickness wide in
// fractional coordinates (cell_of, src/neighbor.cpp:30-62).  Only the fp32 pre-filter uses these positions: two
// atoms reduced around the same centre differ by the reference's minimum-image vector whenever that vector is
// short (the block spans well under half a box length), and everything that decides an output is re-evaluated
// with the reference's exact expression.
struct TileCentre {
    double c[3];
};
__device__ __forceinline__ TileCentre tile_centre(const DBox &box, const CellGrid &g, double cx, double cy, double cz)
{
    const double rcw = 1.0 / g.rc_inv;
    TileCentre t;
    if (!box.triclinic) {
        t.c[0] = box.origin[0] + cx * rcw;
        t.c[1] = box.origin[1] + cy * rcw;
        t.c[2] = box.origin[2] + cz * rcw;
    } else {
        const double f0 = cx * rcw / box.thick[0], f1 = cy * rcw / box.thick[1], f2 = cz * rcw / box.thick[2];
        t.c[0] = box.origin[0] + f0 * box.h[0] + f1 * box.h[3] + f2 * box.h[6];
        t.c[1] = box.origin[1] + f0 * box.h[1] + f1 * box.h[4] + f2 * box.h[7];
        t.c[2] = box.origin[2] + f0 * box.h[2] + f1 * box.h[5] + f2 * box.h[8];
    }
    return t;
}
// returns true when a periodic shift was applied
__device__ __forceinline__ bool rel_image(const DBox &box, const TileCentre &t, double x, double y, double z, double &d0,
                                          double &d1, double &d2)
{
    d0 = x - t.c[0];
    d1 = y - t.c[1];
    d2 = z - t.c[2];
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;
    if (!box.triclinic) {
        if (box.pbc[0]) n0 = rint(d0 * box.hinv[0]);
        if (box.pbc[1]) n1 = rint(d1 * box.hinv[4]);
        if (box.pbc[2]) n2 = rint(d2 * box.hinv[8]);
        d0 -= box.h[0] * n0;
        d1 -= box.h[4] * n1;
        d2 -= box.h[8] * n2;
    } else {
        double f0 = d0 * box.hinv[0] + d1 * box.hinv[3] + d2 * box.hinv[6];
        double f1 = d0 * box.hinv[1] + d1 * box.hinv[4] + d2 * box.hinv[7];
        double f2 = d0 * box.hinv[2] + d1 * box.hinv[5] + d2 * box.hinv[8];
        if (box.pbc[0]) n0 = rint(f0);
        if (box.pbc[1]) n1 = rint(f1);
        if (box.pbc[2]) n2 = rint(f2);
        f0 -= n0;
        f1 -= n1;
        f2 -= n2;
        d0 = f0 * box.h[0] + f1 * box.h[3] + f2 * box.h[6];
        d1 = f0 * box.h[1] + f1 * box.h[4] + f2 * box.h[7];
        d2 = f0 * box.h[2] + f1 * box.h[5] + f2 * box.h[8];
    }
    return (n0 != 0.0) | (n1 != 0.0) | (n2 != 0.0);
}

__device__ __forceinline__ void load_rec(const SortedAtom *p, double &x, double &y, double &z, int &idx, int &cell)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 lo = q[0], hi = q[1];
    x = lo.x;
    y = lo.y;
    z = hi.x;
    idx = __double2loint(hi.y);
    cell = __double2hiint(hi.y);
}

// exclusive scan of v[0..n) in shared memory by warp 0 (n <= 1024); total written to v[n]
__device__ __forceinline__ void warp0_exclusive_scan(int *v, int n)
{
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const int x = i < n ? v[i] : 0;
            int incl = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (i < n) v[i] = carry + incl - x;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) v[n] = carry;
    }
}

// unwrapped cell coordinate u along an axis with n cells -> stored cell (or -1 when not needed/absent)
__device__ __forceinline__ int map_axis(int u, int n)
{
    if (u >= 0 && u < n) return u;
    if (u == -1) return n - 1;
    if (u == n) return 0;
    return -1;
}

// Overflow tile (more atoms than the shared buffers hold): the owned atom at sorted position sg walks
// global memory directly, like k_neighbor_direct.  Out of line: it is cold and large.
template <bool COUNT_ONLY>
__device__ __noinline__ int direct_atom(const TileArgs &A, int sg)
{
    const CellGrid &g = A.g;
    const DBox &box = A.box;
    double xi, yi, zi;
    int my_idx, my_cell, cnt = 0;
    load_rec(A.sorted + sg, xi, yi, zi, my_idx, my_cell);
    if (my_idx >= A.n_rows) return -1;
    wrap_into_box(box, xi, yi, zi);
    int *vrow = A.verlet + (size_t)my_idx * A.M;
    double *drow = A.dist + (size_t)my_idx * A.M;
    int ic, jc, kc;
    cell_decode(g, my_cell, ic, jc, kc);
#pragma unroll 1
    for (int st = 0; st < 27; ++st) {
        const int di = st / 9 - 1, dj = (st / 3) % 3 - 1, dk = st % 3 - 1;
        const int c = cell_linear(g, wrap_cell(ic + di, g.n[0]), wrap_cell(jc + dj, g.n[1]), wrap_cell(kc + dk, g.n[2]));
        if (c < 0) continue;
        const int cb = __ldg(A.cell_start + c), ce = __ldg(A.cell_start + c + 1);
        for (int q = ce - 1; q >= cb; --q) {
            if (q == sg) continue;
            double xj, yj, zj;
            int jdx, jcell;
            load_rec(A.sorted + q, xj, yj, zj, jdx, jcell);
            double dx = xj - xi, dy = yj - yi, dz = zj - zi;
            min_image(box, dx, dy, dz);
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 <= A.rcsq) {
                if (!COUNT_ONLY && cnt < A.M) {
                    vrow[cnt] = jdx;
                    drow[cnt] = sqrt(d2);
                }
                ++cnt;
            }
        }
    }
    A.nn[my_idx] = cnt;
    if (!COUNT_ONLY)
        for (int u = cnt; u < A.M; ++u) {
            vrow[u] = -1;
            drow[u] = A.pad;
        }
    return cnt;
}

template <int T, int TZ, bool COUNT_ONLY>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_neighbor_tiled(const __grid_constant__ TileArgs A)
{
    constexpr int P = T + 2;     // block edge in x, y
    constexpr int PZ = TZ + 2;   // block edge in z (the contiguous direction of the sorted copy)
    constexpr int NPEN = P * P;
    constexpr int NCELL = NPEN * PZ;
    constexpr int CSW = PZ + 1;  // row width of the per-pencil offset table

    extern __shared__ __align__(128) unsigned char smem[];
    SortedAtom *raw = reinterpret_cast<SortedAtom *>(smem);
    float4 *f4 = reinterpret_cast<float4 *>(smem + (size_t)A.cap * sizeof(SortedAtom));
    unsigned short *surv = reinterpret_cast<unsigned short *>(f4 + A.cap);  // [SURV_CAP][TILE_THREADS], interleaved
    int *cs = reinterpret_cast<int *>(surv + TILE_THREADS * SURV_CAP);      // [NPEN][PZ+1] staged offsets
    int *gstart = cs + NPEN * CSW;                                           // [NPEN][PZ] global start per cell (-1 absent)
    int *ptot = gstart + NCELL;                                              // [NPEN+1]
    int *opref = ptot + NPEN + 1;                                            // [T*T+1]
    int *far_flag = opref + T * T + 1;                                       // [2]: beyond the fp32 radius / periodic shift used
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(opref + T * T + 3) + 7) & ~static_cast<uintptr_t>(7));

    const int tid = threadIdx.x;
    const CellGrid &g = A.g;
    int tx, ty, tz;
    if (A.tile_stride != 1) {   // sampled estimate pass (stride > 1) or a grid too large for a 3-D launch (stride 0)
        int tl = blockIdx.x * max(A.tile_stride, 1) + A.tile_offset;
        if (tl >= A.n_tiles) return;
        tz = tl % A.tiles_z;
        tl /= A.tiles_z;
        ty = tl % A.tiles_y;
        tx = tl / A.tiles_y;
    } else {
        // strip order (round 2): grid = (tiles_z, 8 rows of tiles in y, strips * tiles_x) -- x- and y-neighbouring
        // tiles run within a few hundred CTAs of each other, so shared halo planes are re-staged from L2
        tz = blockIdx.x;
        const int strip = blockIdx.z / A.tiles_x;
        tx = blockIdx.z - strip * A.tiles_x;
        ty = strip * 8 + blockIdx.y;
        if (ty >= A.tiles_y) return;
    }
    const int u0x = A.p_lo + tx * T - 1, u0y = ty * T - 1, u0z = tz * TZ - 1;  // unwrapped coords of position 0
    // owned positions inside the block: 1..amax x 1..bmax x 1..kmax (edge tiles are cut at the grid end;
    // positions beyond it only serve as wrapped candidates)
    const int amax = min(T, A.p_hi - (u0x + 1)), bmax = min(T, g.n[1] - (u0y + 1)), kmax = min(TZ, g.n[2] - (u0z + 1));

    if (tid == 0) {
        mbar_init(bar, 1);
        far_flag[0] = 0;
        far_flag[1] = 0;
    }

    // ---- A. population and global start of every cell of the block
    // z slots are in MEMORY order: slot ks holds z position kk = PZ-1-ks (cells are stored in descending z)
    for (int c = tid; c < NCELL; c += TILE_THREADS) {
        const int ks = c % PZ, b = (c / PZ) % P, a = c / (PZ * P);
        const int kk = PZ - 1 - ks;
        int px;
        const int ux = u0x + a;
        if (A.wrap_x) px = (ux >= A.p_lo - 1 && ux <= A.p_hi) ? map_axis(ux, g.n[0]) : -1;
        else px = (ux >= A.p_lo - 1 && ux <= A.p_hi && ux >= 0 && ux < g.nxl) ? ux : -1;
        const int py = map_axis(u0y + b, g.n[1]);
        const int pz = map_axis(u0z + kk, g.n[2]);
        int beg = -1, cnt = 0;
        if (px >= 0 && py >= 0 && pz >= 0) {
            const int cell = (px * g.n[1] + py) * g.n[2] + (g.n[2] - 1 - pz);
            beg = __ldg(A.cell_start + cell);
            cnt = __ldg(A.cell_start + cell + 1) - beg;
        }
        gstart[c] = beg;
        cs[(a * P + b) * CSW + ks + 1] = cnt;  // counts for now
    }
    __syncthreads();
    for (int p = tid; p < NPEN; p += TILE_THREADS) {
        int s = 0;
        for (int kk = 0; kk < PZ; ++kk) s += cs[p * CSW + kk + 1];
        ptot[p] = s;
    }
    __syncthreads();
    warp0_exclusive_scan(ptot, NPEN);
    __syncthreads();
    for (int p = tid; p < NPEN; p += TILE_THREADS) {
        int off = ptot[p];
        cs[p * CSW] = off;
        for (int kk = 0; kk < PZ; ++kk) {
            off += cs[p * CSW + kk + 1];
            cs[p * CSW + kk + 1] = off;
        }
    }
    __syncthreads();
    for (int i = tid; i < T * T; i += TILE_THREADS) {
        const int a = i / T + 1, b = i % T + 1;
        const int p = a * P + b;
        // owned z positions 1..kmax are memory slots PZ-1-kmax .. PZ-2
        opref[i] = (a <= amax && b <= bmax) ? cs[p * CSW + PZ - 1] - cs[p * CSW + PZ - 1 - kmax] : 0;
    }
    __syncthreads();
    warp0_exclusive_scan(opref, T * T);
    __syncthreads();
    const int n_staged = ptot[NPEN];
    const int n_owned = opref[T * T];
    if (n_owned == 0) return;
    const bool fits = n_staged <= A.cap;

    const double rcsq = A.rcsq;
    const DBox &box = A.box;
    const int warp = tid >> 5, lane = tid & 31;
    (void)box;

    if (fits) {
        // ---- B. stage the records: one bulk copy per run of consecutive global cells of a pencil
        if (A.use_tma) {
            if (tid == 0) mbar_expect_tx(bar, (unsigned)n_staged * (unsigned)sizeof(SortedAtom));
            __syncthreads();
            for (int p = tid; p < NPEN; p += TILE_THREADS) {
                int kk = 0;
                while (kk < PZ) {
                    const int beg = gstart[p * PZ + kk];
                    const int dst0 = cs[p * CSW + kk];
                    int total = cs[p * CSW + kk + 1] - dst0;
                    int k2 = kk + 1;
                    if (beg >= 0) {
                        while (k2 < PZ && gstart[p * PZ + k2] == beg + total) {
                            total += cs[p * CSW + k2 + 1] - cs[p * CSW + k2];
                            ++k2;
                        }
                        if (total > 0)
                            bulk_g2s(raw + dst0, A.sorted + beg, (unsigned)total * (unsigned)sizeof(SortedAtom), bar);
                    }
                    kk = k2;
                }
            }
            if (warp == 0) mbar_wait(bar, 0);   // one warp polls; the others sleep on the CTA barrier
            __syncthreads();
        } else {
            for (int c = warp; c < NCELL; c += TILE_THREADS / 32) {
                const int beg = gstart[c];
                if (beg < 0) continue;
                const int p = c / PZ, kk = c % PZ;
                const int dst0 = cs[p * CSW + kk];
                const int chunks = (cs[p * CSW + kk + 1] - dst0) * 2;  // 16-byte chunks
                const double2 *src = reinterpret_cast<const double2 *>(A.sorted + beg);
                double2 *dst = reinterpret_cast<double2 *>(raw + dst0);
                for (int t = lane; t < chunks; t += 32) dst[t] = __ldg(src + t);
            }
            __syncthreads();
        }

        // ---- C. fp32 positions relative to the tile centre (nearest periodic image) + cell tag.
        // One warp per pencil, lanes over the pencil's atoms.
        const int gx0 = A.wrap_x ? (u0x + 1) : (u0x + 1 + g.x0) % g.n[0];  // global x cell of the first owned plane
        const TileCentre ctr = tile_centre(box, g, gx0 + 0.5 * T, u0y + 1 + 0.5 * T, u0z + 1 + 0.5 * TZ);
        for (int s = tid; s < n_staged; s += TILE_THREADS) {   // flat over the staged atoms (all lanes busy)
            const double2 lo = reinterpret_cast<const double2 *>(raw + s)[0];
            const double zr = reinterpret_cast<const double *>(raw + s)[2];
            double d0, d1, d2;
            if (rel_image(box, ctr, lo.x, lo.y, zr, d0, d1, d2)) far_flag[1] = 1;  // some staged atom is a periodic image
            const float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2;
            const float w = __fmaf_rn(f2, f2, __fmaf_rn(f1, f1, f0 * f0));
            if (!(w <= A.w_limit)) far_flag[0] = 1;  // outside the radius the fp32 bound covers (or NaN)
            f4[s] = make_float4(f0, f1, f2, w);
        }
        __syncthreads();
    }
    const bool staged_ok = fits && far_flag[0] == 0;
    const bool tile_shifted = far_flag[1] != 0;

    // ---- D. one thread per owned atom
    int local_max = 0, local_min = INT_MAX;
    if (A.strip) {
        // FLOOR MEASUREMENT (MDB_STRIP=1, tools/neigh_probe.py --strip): tile tables + staging + fp32 copy as
        // above, then every owned atom writes a full row of dummy entries with 16-byte stores and its count --
        // everything the kernel does EXCEPT the search.  Not a product path (results are garbage).
        for (int t = tid; t < n_owned; t += TILE_THREADS) {
            int lo = 0, hi = T * T;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (opref[mid] <= t) lo = mid;
                else hi = mid;
            }
            const int p = (lo / T + 1) * P + (lo % T + 1);
            const int s_i = cs[p * CSW + PZ - 1 - kmax] + (t - opref[lo]);
            const int idx = raw[s_i].idx;
            if (idx >= A.n_rows) continue;
            const float4 me = f4[s_i];
            int *vrow = A.verlet + (size_t)idx * A.M;
            double *drow = A.dist + (size_t)idx * A.M;
            for (int u = 0; u + 3 < A.M; u += 4) {
                *reinterpret_cast<int4 *>(vrow + u) = make_int4(s_i, u, t, idx);
                *reinterpret_cast<double2 *>(drow + u) = make_double2(me.x, me.y);
                *reinterpret_cast<double2 *>(drow + u + 2) = make_double2(me.z, me.w);
            }
            A.nn[idx] = A.M;
        }
        if (tid == 0) {
            atomicMax(A.max_count, A.M);
            atomicMin(A.min_count, A.M);
        }
        return;
    }
    if (!staged_ok) {
#pragma unroll 1
        for (int t = tid; t < n_owned; t += TILE_THREADS) {
            int lo = 0, hi = T * T;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (opref[mid] <= t) lo = mid;
                else hi = mid;
            }
            const int p = (lo / T + 1) * P + (lo % T + 1);
            const int sg = gstart[p * PZ + PZ - 1 - kmax] + (t - opref[lo]);  // owned cells of a pencil: one global run
            const int c = direct_atom<COUNT_ONLY>(A, sg);
            local_max = max(local_max, c);
            if (c >= 0) local_min = min(local_min, c);
        }
    } else {
        const unsigned f4_base = smem_u32(f4), raw_base = smem_u32(raw);
        const unsigned q_base = smem_u32(surv) + 2u * tid;  // entry u at q_base + u * 2 * TILE_THREADS
        const float rc2hi = A.rcsq_hi;
        const double Lx = box.h[0], Ly = box.h[4], Lz = box.h[8];
        const double iLx = box.hinv[0], iLy = box.hinv[4], iLz = box.hinv[8];
        const bool px = box.pbc[0] != 0, py = box.pbc[1] != 0, pz = box.pbc[2] != 0;
        const int M = A.M;
        const bool vec4 = !COUNT_ONLY && (M & 3) == 0;  // rows are 16-byte aligned: int4 / double2 stores
        // odd lanes read the two halves of a 32-byte record in swapped order, which spreads the
        // gathered records over all shared-memory banks (halves the conflicts of the two LDS.128)
        const unsigned sw = (unsigned)(lane & 1) << 4;
#pragma unroll 1
        for (int base = 0; base < n_owned; base += TILE_THREADS) {
            const int t = base + tid;
            const bool active = t < n_owned;
            int pi = 0;
            if (active) {
                int lo = 0, hi = T * T;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (opref[mid] <= t) lo = mid;
                    else hi = mid;
                }
                pi = lo;
            }
            const int a = pi / T + 1, b = pi % T + 1;
            const int p = a * P + b;
            double xi = 0, yi = 0, zi = 0;
            int my_idx = 0, my_cell = 0, cnt = 0, s_i = 0, kk = 1;
            int *vrow = nullptr;
            double *drow = nullptr;
            float fx = 0, fy = 0, fz = 0, thr = 0;
            bool live = false;
            // img: the exact test needs the minimum-image step.  When no staged atom of the tile is a
            // periodic image and this atom lies inside the box, every pre-filter survivor has
            // |dx| <= ~rc << L/2, so n = floor(dx/L + 0.5) = 0 and dx - L*0 == dx: the step is skipped.
            bool img = tile_shifted || box.triclinic;   // triclinic: the fractional round trip of box.h:98-114 always runs
            int i0 = 0, i1 = 0, i2 = 0, i3 = 0;  // last four accepted indices (shift register)
            double e0 = 0.0, e1 = 0.0;            // last two accepted distances
            if (active) {
                s_i = cs[p * CSW + PZ - 1 - kmax] + (t - opref[pi]);
                load_rec(raw + s_i, xi, yi, zi, my_idx, my_cell);
                live = my_idx < A.n_rows;
                if (!box.triclinic) {
                    if (px) { const double d = xi - box.origin[0]; img |= !(d >= box.wrap_t[0][1] && d < box.wrap_t[0][2]); }
                    if (py) { const double d = yi - box.origin[1]; img |= !(d >= box.wrap_t[1][1] && d < box.wrap_t[1][2]); }
                    if (pz) { const double d = zi - box.origin[2]; img |= !(d >= box.wrap_t[2][1] && d < box.wrap_t[2][2]); }
                }
                wrap_into_box(box, xi, yi, zi);
                const float4 me = f4[s_i];
                fx = -2.0f * me.x;
                fy = -2.0f * me.y;
                fz = -2.0f * me.z;
                thr = rc2hi - me.w;
                // my memory slot along z: owned slots are PZ-1-kmax .. PZ-2
                kk = PZ - 1 - kmax;
                const int *prow = cs + p * CSW;
                while (kk < PZ - 2 && prow[kk + 1] <= s_i) ++kk;
                vrow = A.verlet + (size_t)my_idx * M;
                drow = A.dist + (size_t)my_idx * M;
            }
            unsigned q_top = q_base;                            // queue write pointer
            const unsigned q_full = q_base + 2u * TILE_THREADS * SURV_CAP;
            const unsigned q_warn = q_base + 2u * TILE_THREADS * SURV_RESERVE;
            const unsigned self_q = (unsigned)s_i;

            // phase 2: exact test of the queued survivors, in queue (= reference) order
            auto drain_t = [&](auto IMG) {
#pragma unroll 1
                for (unsigned qa = q_base; qa < q_top; qa += 2u * TILE_THREADS) {
                    const unsigned q = lds_u16(qa);
                    if (q == self_q) continue;
                    const unsigned ra = raw_base + 32u * q;
                    double a0, a1, b0, b1;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a0), "=d"(a1) : "r"(ra + sw));
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b0), "=d"(b1) : "r"(ra + (sw ^ 16u)));
                    const double xj = sw ? b0 : a0, yj = sw ? b1 : a1, zj = sw ? a0 : b0;
                    const int jdx = __double2loint(sw ? a1 : b1);
                    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
                    if (decltype(IMG)::value) {
                        if (box.triclinic) min_image(box, dx, dy, dz);
                        else {
                            if (px) dx = near_image(dx, Lx, iLx);
                            if (py) dy = near_image(dy, Ly, iLy);
                            if (pz) dz = near_image(dz, Lz, iLz);
                        }
                    }
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 <= rcsq) {
                        if (!COUNT_ONLY && cnt < M) {
                            const double dd = sqrt(d2);
                            if (vec4) {
                                i0 = i1;
                                i1 = i2;
                                i2 = i3;
                                i3 = jdx;
                                e0 = e1;
                                e1 = dd;
                                if (cnt & 1) *reinterpret_cast<double2 *>(drow + cnt - 1) = make_double2(e0, e1);
                                if ((cnt & 3) == 3) *reinterpret_cast<int4 *>(vrow + cnt - 3) = make_int4(i0, i1, i2, i3);
                            } else {
                                vrow[cnt] = jdx;
                                drow[cnt] = dd;
                            }
                        }
                        ++cnt;
                    }
                }
                q_top = q_base;
            };
            auto drain = [&]() {
                if (img) drain_t(std::true_type{});
                else drain_t(std::false_type{});
            };

            // phase 1: the 9 pencils of the stencil in (x, y) order.  The three z cells of a pencil are
            // one contiguous staged run stored in descending z, so a single backward walk yields
            // cell z-1, z, z+1, each in descending original index: the reference's order.
            int pen_off = -P - 1, pen_y = 0;  // pencil offsets of the 3 x 3 stencil in (x, y) order
#pragma unroll 1
            for (int pen = 0; pen < 9; ++pen) {
                if (live) {
                    const int *row = cs + (p + pen_off) * CSW + kk;
                    const unsigned abeg = f4_base + 16u * (unsigned)row[-1];
                    unsigned aq = f4_base + 16u * (unsigned)row[2];  // one past the last candidate
#pragma unroll 1
                    while (aq > abeg) {
                        // scan at most as many candidates as the queue has room for, then (rarely) drain
                        const unsigned room = (q_full - q_top) / (2u * TILE_THREADS);
                        const unsigned astop = (aq - abeg > 16u * room) ? aq - 16u * room : abeg;
#pragma unroll 2
                        while (aq > astop) {
                            aq -= 16u;
                            const float4 o = lds_f4(aq);
                            // |rj|^2 - 2 ri.rj  <=  rc^2 (1 + guard) - |ri|^2   (the atom itself passes; phase 2 drops it)
                            const float t2 = __fmaf_rn(fx, o.x, __fmaf_rn(fy, o.y, __fmaf_rn(fz, o.z, o.w)));
                            if (t2 <= thr) {
                                sts_u16(q_top, (aq - f4_base) >> 4);
                                q_top += 2u * TILE_THREADS;
                            }
                        }
                        if (aq > abeg) drain();
                    }
                }
                if (++pen_y == 3) {
                    pen_y = 0;
                    pen_off += P - 2;
                } else ++pen_off;
                if (__any_sync(0xffffffffu, q_top > q_warn) || pen == 8) drain();
            }
            if (active && live) {
                A.nn[my_idx] = cnt;
                local_max = max(local_max, cnt);
                local_min = min(local_min, cnt);
                if (!COUNT_ONLY) {
                    if (vec4 && cnt < M) {
                        // still in the shift registers: the last (cnt & 3) indices and (cnt & 1) distance
                        const int r = cnt & 3;
                        if (r >= 3) vrow[cnt - 3] = i1;
                        if (r >= 2) vrow[cnt - 2] = i2;
                        if (r >= 1) vrow[cnt - 1] = i3;
                        if (cnt & 1) drow[cnt - 1] = e1;
                    }
                    for (int u = cnt; u < M; ++u) {
                        vrow[u] = -1;
                        drow[u] = A.pad;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, d));
        local_min = min(local_min, __shfl_xor_sync(0xffffffffu, local_min, d));
    }
    if (lane == 0 && local_max > 0) atomicMax(A.max_count, local_max);
    if (lane == 0 && local_min != INT_MAX) atomicMin(A.min_count, local_min);
}


// =====================================================================================================
// Cooperative kernel (round 2).  Same tile staging as k_neighbor_tiled; the two phases are mapped the way
// each of them is cheap (measured: profiles/r2_*):
//
//   phase 1  one THREAD per owned atom walks its 27-cell stencil (9 pencils x one contiguous 3-cell run,
//            backwards = the reference's order) with the fp32 pre-filter and pushes survivors onto its
//            own queue in shared memory -- compaction is free when every thread owns its queue;
//   flatten  a warp scan of the queue lengths turns the 32 queues into one atom-major ring of
//            (candidate | owner | row slot) words;
//   phase 2  the ring is consumed 32 survivors at a time, one per LANE, all lanes busy: exact f64 test of
//            the reference (xi wrapped, x[j] raw, left-to-right sum, <= rc^2), sqrt, and the rows leave as
//            contiguous 4-byte / 8-byte segments of consecutive lanes.
//            The row slot of a survivor is its rank among the ACCEPTED survivors of its atom: entries of one
//            atom are consecutive in the ring, so a ballot of the accept flags, a ballot of the "first
//            entry of an atom" flags and one popc give it (plus a carry for an atom that straddles two
//            calls).  The exact count of every atom is left in shared memory for its thread, which writes
//            neighbor_number and the row tail.
// Row order, distances and counts are bit-identical to k_neighbor_tiled / k_neighbor_direct / the reference.
struct __align__(32) OwnAtom {
    double x, y, z;  // wrapped position (box.h:131-176), what the reference uses as atom i
    int idx;         // original index (row)
    int flags;       // bit 0: outside the box -> the minimum-image step is needed
};

constexpr int COOP_RCAP = 256;   // survivor ring per warp (entries, power of two)
constexpr int COOP_QCAP = 26;    // per-thread queue (uint16 entries)
constexpr int COOP_CHUNK = 12;   // a thread enters a run with at least this much room in its queue

__device__ __forceinline__ unsigned lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <int T, int TZ, int NT, bool COUNT_ONLY>
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 2) k_neighbor_coop(const __grid_constant__ TileArgs A)
{
    constexpr int P = T + 2;
    constexpr int PZ = TZ + 2;
    constexpr int NPEN = P * P;
    constexpr int NCELL = NPEN * PZ;
    constexpr int CSW = PZ + 1;
    constexpr int NW = NT / 32;

    extern __shared__ __align__(128) unsigned char smem[];
    // every array is addressed as smem + offset (no pointer -> integer -> pointer round trip, so the compiler
    // keeps all accesses in the shared window: LDS / STS instead of generic loads)
    const unsigned o_f4 = (unsigned)A.cap * 32u;                       // [cap] 16 B
    const unsigned o_own = o_f4 + (unsigned)A.cap * 16u;               // [ocap] 32 B
    const unsigned o_ring = o_own + (unsigned)A.ocap * 32u;            // [NW][COOP_RCAP] u32
    const unsigned o_queue = o_ring + NW * COOP_RCAP * 4u;             // [COOP_QCAP][NT] u16, interleaved
    const unsigned o_bar = o_queue + COOP_QCAP * NT * 2u;              // mbarrier (8 B, 16-byte slot)
    const unsigned o_cs = o_bar + 16u;                                 // [NPEN][CSW] int
    const unsigned o_gstart = o_cs + NPEN * CSW * 4u;                  // [NCELL]
    const unsigned o_ptot = o_gstart + NCELL * 4u;                     // [NPEN + 1]
    const unsigned o_opref = o_ptot + (NPEN + 1) * 4u;                 // [T*T + 1]
    const unsigned o_flag = o_opref + (T * T + 1) * 4u;                // [4]
    const unsigned o_acnt = o_flag + 16u;                              // [ocap] u16 accepted neighbours per owned atom
    SortedAtom *raw = reinterpret_cast<SortedAtom *>(smem);
    float4 *f4 = reinterpret_cast<float4 *>(smem + o_f4);
    OwnAtom *own = reinterpret_cast<OwnAtom *>(smem + o_own);
    unsigned *ring = reinterpret_cast<unsigned *>(smem + o_ring);
    unsigned short *queue = reinterpret_cast<unsigned short *>(smem + o_queue);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + o_bar);
    int *cs = reinterpret_cast<int *>(smem + o_cs);
    int *gstart = reinterpret_cast<int *>(smem + o_gstart);
    int *ptot = reinterpret_cast<int *>(smem + o_ptot);
    int *opref = reinterpret_cast<int *>(smem + o_opref);
    int *far_flag = reinterpret_cast<int *>(smem + o_flag);   // [0] beyond the fp32 radius, [1] periodic shift used, [2] atoms to redo
    unsigned short *acnt = reinterpret_cast<unsigned short *>(smem + o_acnt);

    const int tid = threadIdx.x;
    const CellGrid &g = A.g;
    int tx, ty, tz;
    if (A.tile_stride != 1) {   // sampled estimate pass (stride > 1) or a grid too large for a 3-D launch (stride 0): linear list of tiles
        int tl = blockIdx.x * max(A.tile_stride, 1) + A.tile_offset;
        if (tl >= A.n_tiles) return;
        tz = tl % A.tiles_z;
        tl /= A.tiles_z;
        ty = tl % A.tiles_y;
        tx = tl / A.tiles_y;
    } else {
        // strip order: grid = (tiles_z, 8 rows of tiles in y, strips * tiles_x).  x- and y-neighbouring tiles
        // run within a few hundred CTAs of each other, so their shared halo planes are still in L2 when
        // they are staged again
        tz = blockIdx.x;
        const int strip = blockIdx.z / A.tiles_x;
        tx = blockIdx.z - strip * A.tiles_x;
        ty = strip * 8 + blockIdx.y;
        if (ty >= A.tiles_y) return;
    }
    const int u0x = A.p_lo + tx * T - 1, u0y = ty * T - 1, u0z = tz * TZ - 1;
    const int amax = min(T, A.p_hi - (u0x + 1)), bmax = min(T, g.n[1] - (u0y + 1)), kmax = min(TZ, g.n[2] - (u0z + 1));

    if (tid == 0) {
        mbar_init(bar, 1);
        far_flag[0] = 0;
        far_flag[1] = 0;
        far_flag[2] = 0;
    }

    // ---- A. population and global start of every cell of the block (z slots in memory order)
    for (int c = tid; c < NCELL; c += NT) {
        const int ks = c % PZ, b = (c / PZ) % P, a = c / (PZ * P);
        const int kk = PZ - 1 - ks;
        int px;
        const int ux = u0x + a;
        if (A.wrap_x) px = (ux >= A.p_lo - 1 && ux <= A.p_hi) ? map_axis(ux, g.n[0]) : -1;
        else px = (ux >= A.p_lo - 1 && ux <= A.p_hi && ux >= 0 && ux < g.nxl) ? ux : -1;
        const int py = map_axis(u0y + b, g.n[1]);
        const int pz = map_axis(u0z + kk, g.n[2]);
        int beg = -1, cnt = 0;
        if (px >= 0 && py >= 0 && pz >= 0) {
            const int cell = (px * g.n[1] + py) * g.n[2] + (g.n[2] - 1 - pz);
            beg = __ldg(A.cell_start + cell);
            cnt = __ldg(A.cell_start + cell + 1) - beg;
        }
        gstart[c] = beg;
        cs[(a * P + b) * CSW + ks + 1] = cnt;
    }
    __syncthreads();
    for (int p = tid; p < NPEN; p += NT) {
        int s = 0;
        for (int kk = 0; kk < PZ; ++kk) s += cs[p * CSW + kk + 1];
        ptot[p] = s;
    }
    __syncthreads();
    warp0_exclusive_scan(ptot, NPEN);
    __syncthreads();
    for (int p = tid; p < NPEN; p += NT) {
        int off = ptot[p];
        cs[p * CSW] = off;
        for (int kk = 0; kk < PZ; ++kk) {
            off += cs[p * CSW + kk + 1];
            cs[p * CSW + kk + 1] = off;
        }
    }
    __syncthreads();
    for (int i = tid; i < T * T; i += NT) {
        const int a = i / T + 1, b = i % T + 1;
        const int p = a * P + b;
        opref[i] = (a <= amax && b <= bmax) ? cs[p * CSW + PZ - 1] - cs[p * CSW + PZ - 1 - kmax] : 0;
    }
    __syncthreads();
    warp0_exclusive_scan(opref, T * T);
    __syncthreads();
    const int n_staged = ptot[NPEN];
    const int n_owned = opref[T * T];
    if (n_owned == 0) return;
    const bool fits = n_staged <= A.cap && n_owned <= A.ocap;

    const DBox &box = A.box;
    const int warp = tid >> 5, lane = tid & 31;

    // owned atom t of the tile -> its pencil (index into the T x T owned pencils) by bisection over opref
    auto pencil_of = [&](int t) {
        int lo = 0, hi = T * T;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (opref[mid] <= t) lo = mid;
            else hi = mid;
        }
        return lo;
    };

    if (fits) {
        // ---- B. stage the records: one bulk copy per run of consecutive global cells of a pencil
        if (A.use_tma) {
            if (tid == 0) mbar_expect_tx(bar, (unsigned)n_staged * (unsigned)sizeof(SortedAtom));
            __syncthreads();
            for (int p = tid; p < NPEN; p += NT) {
                int kk = 0;
                while (kk < PZ) {
                    const int beg = gstart[p * PZ + kk];
                    const int dst0 = cs[p * CSW + kk];
                    int total = cs[p * CSW + kk + 1] - dst0;
                    int k2 = kk + 1;
                    if (beg >= 0) {
                        while (k2 < PZ && gstart[p * PZ + k2] == beg + total) {
                            total += cs[p * CSW + k2 + 1] - cs[p * CSW + k2];
                            ++k2;
                        }
                        if (total > 0)
                            bulk_g2s(raw + dst0, A.sorted + beg, (unsigned)total * (unsigned)sizeof(SortedAtom), bar);
                    }
                    kk = k2;
                }
            }
            if (warp == 0) mbar_wait(bar, 0);   // one warp polls; the others sleep on the CTA barrier
            __syncthreads();
        } else {
            for (int c = warp; c < NCELL; c += NW) {
                const int beg = gstart[c];
                if (beg < 0) continue;
                const int p = c / PZ, kk = c % PZ;
                const int dst0 = cs[p * CSW + kk];
                const int chunks = (cs[p * CSW + kk + 1] - dst0) * 2;
                const double2 *src = reinterpret_cast<const double2 *>(A.sorted + beg);
                double2 *dst = reinterpret_cast<double2 *>(raw + dst0);
                for (int t = lane; t < chunks; t += 32) dst[t] = __ldg(src + t);
            }
            __syncthreads();
        }

        // ---- C. fp32 positions relative to the tile centre (nearest periodic image), flat over the staged atoms
        const int gx0 = A.wrap_x ? (u0x + 1) : (u0x + 1 + g.x0) % g.n[0];
        const TileCentre ctr = tile_centre(box, g, gx0 + 0.5 * T, u0y + 1 + 0.5 * T, u0z + 1 + 0.5 * TZ);
        for (int s = tid; s < n_staged; s += NT) {
            const double2 lo = reinterpret_cast<const double2 *>(raw + s)[0];
            const double z = reinterpret_cast<const double *>(raw + s)[2];
            double d0, d1, d2;
            if (rel_image(box, ctr, lo.x, lo.y, z, d0, d1, d2)) far_flag[1] = 1;   // some staged atom is a periodic image
            const float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2;
            const float w = __fmaf_rn(f2, f2, __fmaf_rn(f1, f1, f0 * f0));
            if (!(w <= A.w_limit)) far_flag[0] = 1;   // outside the radius the fp32 bound covers (or NaN)
            f4[s] = make_float4(f0, f1, f2, w);
        }
        __syncthreads();
    }
    const bool staged_ok = fits && far_flag[0] == 0;
    const bool tile_shifted = far_flag[1] != 0;

    int local_max = 0, local_min = INT_MAX;
    if (!staged_ok) {
        // overflow tile / atoms beyond the fp32 radius: exact per-thread walk through global memory
#pragma unroll 1
        for (int t = tid; t < n_owned; t += NT) {
            const int lo = pencil_of(t);
            const int p = (lo / T + 1) * P + (lo % T + 1);
            const int sg = gstart[p * PZ + PZ - 1 - kmax] + (t - opref[lo]);
            const int c = direct_atom<COUNT_ONLY>(A, sg);
            local_max = max(local_max, c);
            if (c >= 0) local_min = min(local_min, c);
        }
    } else {
        const unsigned f4_base = smem_u32(f4), raw_base = smem_u32(raw), own_base = smem_u32(own);
        const unsigned ring_base = smem_u32(ring + warp * COOP_RCAP);
        const unsigned q_base = smem_u32(queue) + 2u * tid;   // entry u of this thread at q_base + u * 2 * NT
        constexpr unsigned QS = 2u * NT;
        const float rc2hi = A.rcsq_hi;
        const double rcsq = A.rcsq;
        const double Lx = box.h[0], Ly = box.h[4], Lz = box.h[8];
        const double iLx = box.hinv[0], iLy = box.hinv[4], iLz = box.hinv[8];
        const bool px = box.pbc[0] != 0, py = box.pbc[1] != 0, pz = box.pbc[2] != 0;
        const int M = A.M;
        unsigned head = 0, tail = 0;   // ring positions (monotonic, warp-uniform)

        // ---- phase 2: 32 survivors per call, one per lane
        const unsigned lt = lanemask_lt();
        const unsigned acnt_base = smem_u32(acnt);
        auto consume = [&](int nvalid) {
            const bool valid = lane < nvalid;
            unsigned e = lds_u32(ring_base + (((head + (unsigned)lane) & (COOP_RCAP - 1)) << 2));
            if (!valid) e = 0xffffu << 12;   // owner no real atom has: a segment of its own
            const unsigned cand = e & 0xfffu, t = e >> 12;
            const unsigned tt = valid ? t : 0u;
            // (idle lanes read a staged record instead of own[0], which another warp may be writing)
            const unsigned ra = raw_base + 32u * cand, oa = valid ? own_base + 32u * tt : raw_base;
            double xj, yj, zj, wj, xi, yi, zi, wi;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj), "=d"(yj) : "r"(ra));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(zj), "=d"(wj) : "r"(ra + 16u));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi), "=d"(yi) : "r"(oa));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(zi), "=d"(wi) : "r"(oa + 16u));
            const unsigned have = valid ? lds_u16(acnt_base + 2u * tt) : 0u;   // accepted in earlier calls
            const int jdx = __double2loint(wj);
            const int idx = __double2loint(wi), fl = __double2hiint(wi);
            double dx = xj - xi, dy = yj - yi, dz = zj - zi;
            if (box.triclinic) min_image(box, dx, dy, dz);   // the fractional round trip of box.h:98-114 always runs
            else if (tile_shifted || __any_sync(0xffffffffu, valid && (fl & 1))) {
                if (px) dx = near_image(dx, Lx, iLx);
                if (py) dy = near_image(dy, Ly, iLy);
                if (pz) dz = near_image(dz, Lz, iLz);
            }
            const double d2 = dx * dx + dy * dy + dz * dz;
            const bool acc = valid && d2 <= rcsq;
            // entries of one atom are consecutive: a new atom starts where the owner changes
            const unsigned tprev = __shfl_up_sync(0xffffffffu, t, 1);
            const bool first = lane == 0 || t != tprev;
            const unsigned mfirst = __ballot_sync(0xffffffffu, first), macc = __ballot_sync(0xffffffffu, acc);
            const int start = 31 - __clz(mfirst & (lt | (1u << lane)));   // lane 0 always starts a run
            const unsigned slot = have + __popc(macc & lt & ~((1u << start) - 1u));
            if (acc && (int)slot < M) {
                const size_t off = (size_t)idx * M + slot;
                A.verlet[off] = jdx;
                A.dist[off] = sqrt(d2);
            }
            // the last entry of every run leaves the atom's running count behind
            if (valid && (lane == 31 || ((mfirst >> (lane + 1)) & 1u)))
                sts_u16(acnt_base + 2u * tt, min(slot + (acc ? 1u : 0u), 65535u));
            head += (unsigned)nvalid;
            __syncwarp();
        };

#pragma unroll 1
        for (int base = warp * 32; base < n_owned; base += NT) {
            const int t = base + lane;
            const bool active = t < n_owned;
            int idx = 0, s_i = 0, kk = 1, p = P + 1;
            bool live = false;
            float fx = 0.f, fy = 0.f, fz = 0.f, thr = 0.f;
            if (active) {
                const int pi = pencil_of(t);
                p = (pi / T + 1) * P + (pi % T + 1);
                const int *prow = cs + p * CSW;
                s_i = prow[PZ - 1 - kmax] + (t - opref[pi]);
                kk = PZ - 1 - kmax;   // my memory slot along z: owned slots are PZ-1-kmax .. PZ-2
                while (kk < PZ - 2 && prow[kk + 1] <= s_i) ++kk;
                const double2 lo = reinterpret_cast<const double2 *>(raw + s_i)[0];
                const double2 hi = reinterpret_cast<const double2 *>(raw + s_i)[1];
                idx = __double2loint(hi.y);
                live = idx < A.n_rows;
                double xi = lo.x, yi = lo.y, zi = hi.x;
                int fl = 0;
                if (!box.triclinic) {
                    if (box.pbc[0]) { const double d = xi - box.origin[0]; fl |= !(d >= box.wrap_t[0][1] && d < box.wrap_t[0][2]); }
                    if (box.pbc[1]) { const double d = yi - box.origin[1]; fl |= !(d >= box.wrap_t[1][1] && d < box.wrap_t[1][2]); }
                    if (box.pbc[2]) { const double d = zi - box.origin[2]; fl |= !(d >= box.wrap_t[2][1] && d < box.wrap_t[2][2]); }
                }
                wrap_into_box(box, xi, yi, zi);
                OwnAtom o;
                o.x = xi;
                o.y = yi;
                o.z = zi;
                o.idx = idx;
                o.flags = fl;
                own[t] = o;
                acnt[t] = 0;
                const float4 me = f4[s_i];
                fx = -2.0f * me.x;
                fy = -2.0f * me.y;
                fz = -2.0f * me.z;
                thr = rc2hi - me.w;
            }
            unsigned q_top = q_base;   // queue write pointer
            const unsigned q_full = q_base + QS * COOP_QCAP, q_warn = q_base + QS * (COOP_QCAP - COOP_CHUNK);
            unsigned cnt = 0;          // pre-filter count so far (the estimate pass reports it)
            const unsigned tbits = (unsigned)t << 12;

            // queues -> ring (atom-major), phase 2 as the ring fills
            auto flush = [&]() {
                const int n = (int)((q_top - q_base) / QS);
                bool pending = n > 0;
#pragma unroll 1
                while (__any_sync(0xffffffffu, pending)) {
                    int incl = pending ? n : 0;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    const int space = COOP_RCAP - (int)(tail - head);
                    const bool go = pending && incl <= space;   // a prefix of the pending lanes
                    const unsigned gom = __ballot_sync(0xffffffffu, go);
                    const int total = __shfl_sync(0xffffffffu, incl, 31 - __clz(gom | 1u));
                    const int maxn = __reduce_max_sync(0xffffffffu, go ? n : 0);
                    unsigned ra = ring_base + (((tail + (unsigned)(incl - n)) & (COOP_RCAP - 1)) << 2);
                    unsigned qa = q_base;
#pragma unroll 1
                    for (int i = 0; i < maxn; ++i) {
                        if (go && i < n) {
                            const unsigned k = ((lds_u16(qa) - f4_base) & 0xffffu) >> 4;
                            sts_u32(ra, k | tbits);
                            qa += QS;
                            ra += 4u;
                            if (ra == ring_base + COOP_RCAP * 4u) ra = ring_base;
                        }
                    }
                    if (go) {
                        pending = false;
                        cnt += (unsigned)n;
                        q_top = q_base;
                    }
                    tail += gom ? (unsigned)total : 0u;
                    __syncwarp();
                    // everything that was copied is consumed before the next copy: a call never sees two
                    // separate runs of one atom (its running count is read once per call)
                    while (tail != head) consume(min(32, (int)(tail - head)));
                }
            };

            // phase 1: the 9 pencils of the stencil in (x, y) order.  The three z cells of a pencil are one
            // contiguous staged run stored in descending z, so a single backward walk yields cell z-1, z,
            // z+1, each in descending original index: the reference's order.  The centre pencil is walked in
            // two pieces that leave the atom itself out.
            int pen_off = -P - 1, pen_y = 0;
            const unsigned self = f4_base + 16u * (unsigned)s_i;
#pragma unroll 1
            for (int pen = 0; pen < 9; ++pen) {
                unsigned abeg = f4_base, aend = f4_base;
                if (live) {
                    const int *row = cs + (p + pen_off) * CSW + kk;
                    abeg = f4_base + 16u * (unsigned)row[-1];
                    aend = f4_base + 16u * (unsigned)row[2];   // one past the last candidate
                }
#pragma unroll 1
                for (int piece = 0; piece < (pen == 4 ? 2 : 1); ++piece) {
                    unsigned lo_a = abeg, aq = aend;
                    if (pen == 4 && live) {
                        if (piece == 0) lo_a = self + 16u;   // candidates above me
                        else aq = self;                      // candidates below me
                    }
#pragma unroll 1
                    while (__any_sync(0xffffffffu, aq > lo_a)) {
                        // scan at most as many candidates as the queue has room for (all of the run unless the
                        // frame is dense); the next candidate is loaded while the current one is tested
                        const unsigned room = (q_full - q_top) / QS;
                        const unsigned astop = (aq - lo_a > 16u * room && aq > lo_a) ? aq - 16u * room : lo_a;
                        float4 o = lds_f4(aq - 16u);   // (a read below the run is harmless: it stays inside the staged arrays)
#pragma unroll 2
                        while (aq > astop) {
                            aq -= 16u;
                            const float4 c = o;
                            o = lds_f4(aq - 16u);
                            // |rj|^2 - 2 ri.rj  <=  rc^2 (1 + guard) - |ri|^2
                            if (__fmaf_rn(fx, c.x, __fmaf_rn(fy, c.y, __fmaf_rn(fz, c.z, c.w))) <= thr) {
                                if (COUNT_ONLY) ++cnt;
                                else {
                                    sts_u16(q_top, aq);
                                    q_top += QS;
                                }
                            }
                        }
                        if (!COUNT_ONLY && __any_sync(0xffffffffu, q_top > q_warn || aq > lo_a)) flush();
                    }
                }
                if (++pen_y == 3) {
                    pen_y = 0;
                    pen_off += P - 2;
                } else ++pen_off;
            }
            if (!COUNT_ONLY) {
                flush();
                while (tail != head) consume(min(32, (int)(tail - head)));   // the batch's last (partial) call
                cnt = active ? acnt[t] : 0u;                                  // exact count
            }
            if (live) {
                A.nn[idx] = (int)cnt;
                local_max = max(local_max, (int)cnt);
                local_min = min(local_min, (int)cnt);
                if (!COUNT_ONLY && (int)cnt < M) {   // row tail (-1 / rc+1)
                    int *vrow = A.verlet + (size_t)idx * M;
                    double *drow = A.dist + (size_t)idx * M;
                    for (int u = (int)cnt; u < M; ++u) {
                        vrow[u] = -1;
                        drow[u] = A.pad;
                    }
                }
            }
        }
    }
    {
        // (estimate pass: pre-filter counts, i.e. upper bounds; fill pass: exact counts)
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, d));
            local_min = min(local_min, __shfl_xor_sync(0xffffffffu, local_min, d));
        }
        if (lane == 0 && local_max > 0) atomicMax(A.max_count, local_max);
        if (lane == 0 && local_min != INT_MAX) atomicMin(A.min_count, local_min);
    }
}

template <int T, int TZ, int NT> size_t coop_smem_bytes(int cap, int ocap)
{
    constexpr int P = T + 2, PZ = TZ + 2, NPEN = P * P, NCELL = NPEN * PZ, NW = NT / 32;
    return (size_t)cap * (sizeof(SortedAtom) + sizeof(float4)) + (size_t)ocap * (sizeof(OwnAtom) + 2) +
           sizeof(unsigned) * NW * COOP_RCAP + sizeof(unsigned short) * COOP_QCAP * NT +
           sizeof(int) * (NPEN * (PZ + 1) + NCELL + NPEN + 1 + T * T + 1 + 4) + 16 + 64;
}

template <int T, int TZ, int NT> void launch_cells_T(const TileArgs &A, int nblocks, cudaStream_t st)
{
    const size_t smem = coop_smem_bytes<T, TZ, NT>(A.cap, A.ocap);
    // per launch: the attribute belongs to the current device (several devices per process: group.cu)
    CUDA_TRY(cudaFuncSetAttribute(k_neighbor_coop<T, TZ, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_neighbor_coop<T, TZ, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    dim3 grid(nblocks, 1, 1);
    if (A.tile_stride == 1) {   // strip order (see the kernel)
        const int strips = (A.tiles_y + 7) / 8;
        grid = dim3(A.tiles_z, A.tiles_y < 8 ? A.tiles_y : 8, strips * A.tiles_x);
    }
    if (A.count_only) MDB_LAUNCH((k_neighbor_coop<T, TZ, NT, true>), grid, NT, smem, st, A);
    else MDB_LAUNCH((k_neighbor_coop<T, TZ, NT, false>), grid, NT, smem, st, A);
}

template <int T, int TZ> size_t tile_smem_bytes(int cap)
{
    constexpr int P = T + 2, PZ = TZ + 2, NPEN = P * P, NCELL = NPEN * PZ;
    return (size_t)cap * (sizeof(SortedAtom) + sizeof(float4)) + sizeof(unsigned short) * TILE_THREADS * SURV_CAP +
           sizeof(int) * (NPEN * (PZ + 1) + NCELL + NPEN + 1 + T * T + 3) + 16;
}

template <int T, int TZ> void launch_T(const TileArgs &A, int nblocks, cudaStream_t st)
{
    const size_t smem = tile_smem_bytes<T, TZ>(A.cap);
    // per launch: the attribute belongs to the current device (several devices per process: group.cu)
    CUDA_TRY(cudaFuncSetAttribute(k_neighbor_tiled<T, TZ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_neighbor_tiled<T, TZ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    dim3 grid(nblocks, 1, 1);
    if (A.tile_stride == 1) {   // strip order (see the kernel)
        const int strips = (A.tiles_y + 7) / 8;
        grid = dim3(A.tiles_z, A.tiles_y < 8 ? A.tiles_y : 8, strips * A.tiles_x);
    }
    if (A.count_only) MDB_LAUNCH((k_neighbor_tiled<T, TZ, true>), grid, TILE_THREADS, smem, st, A);
    else MDB_LAUNCH((k_neighbor_tiled<T, TZ, false>), grid, TILE_THREADS, smem, st, A);
}

// =====================================================================================================
// Fused neighbour search + fixed-cutoff CNA (SURVEY.md 8d "fused neighbour+CNA (no list)", 28 B/atom).
// Replaces build_neighbor (src/neighbor.cpp:130-186) followed by FixedCNA (src/cna.cpp:429-506) for callers
// that only read the labels: the neighbour list never goes to HBM.  Same tile staging as the list kernels;
// one thread per owned atom:
//   1. the 27-cell stencil with the fp32 form of the distance test.  A candidate well inside the cut-off
//      (fp32 d^2 <= rc^2 (1 - 4e-4)) is a neighbour, one well outside is not -- the band is > 5x the fp32
//      error bound -- and a candidate inside the band takes the reference's exact f64 test (xi wrapped,
//      x[j] raw, minimum image, <= rc^2).  The neighbour SET therefore equals the reference's list row;
//      CNA does not depend on its order.
//   2. atoms with 12 or 14 neighbours: the 66 / 91 bond tests among the neighbours on the staged fp32
//      positions (all nearest images of one tile centre, so differences are minimum-image vectors), a pair
//      within 1e-4 rc^2 of the threshold sends the atom to the exact f64 bond matrix (cna.cpp:149-161 on
//      raw coordinates); signatures from the bond rows in registers (cna_core.cuh).
// Labels are identical to FixedCNA on the reference's list (tests/test_gpu_fused.py, at-scale parity tool).
constexpr int FUSED_QCAP = 24;   // candidate slots per thread (14 neighbours + guard-band candidates)

// exact membership test of the list build for one pair (cold)
__device__ __noinline__ bool fused_exact_neighbor(const DBox &box, const SortedAtom *raw, int s_i, int k, double rcsq)
{
    double xi = raw[s_i].x, yi = raw[s_i].y, zi = raw[s_i].z;
    wrap_into_box(box, xi, yi, zi);
    double dx = raw[k].x - xi, dy = raw[k].y - yi, dz = raw[k].z - zi;
    min_image(box, dx, dy, dz);
    return dx * dx + dy * dy + dz * dz <= rcsq;
}

// exact bond matrix + signatures for one atom (cold): cna.cpp:149-161 on the raw coordinates of the neighbours
template <int STRIDE>
__device__ __noinline__ int fused_exact_cna(const DBox &box, const SortedAtom *raw, const unsigned short *q, int nn,
                                            double cutsq, unsigned short *nb)
{
    for (int a = 0; a < nn; ++a) nb[a * STRIDE] = 0;
    for (int a = 0; a < nn; ++a) {
        const SortedAtom &A_ = raw[q[a * STRIDE]];
        for (int b = a + 1; b < nn; ++b) {
            const SortedAtom &B_ = raw[q[b * STRIDE]];
            double dx = B_.x - A_.x, dy = B_.y - A_.y, dz = B_.z - A_.z;
            min_image(box, dx, dy, dz);
            if (dx * dx + dy * dy + dz * dz <= cutsq) {
                nb[a * STRIDE] |= (unsigned short)(1u << b);
                nb[b * STRIDE] |= (unsigned short)(1u << a);
            }
        }
    }
    return cna_label(cna_signatures_smem<STRIDE>(nb, nn));
}

// Exact fallback for one atom through global memory (cold): atoms of overflow tiles and atoms far outside
// the box (clamped into an edge cell of an open axis).  Neighbours by the list builder's exact test,
// bonds by cna.cpp:149-161.
template <int STRIDE>
__device__ __noinline__ int fused_direct_atom(const TileArgs &A, int sg, unsigned short *nb)
{
    const CellGrid &g = A.g;
    const DBox &box = A.box;
    double xi, yi, zi;
    int my_idx, my_cell;
    load_rec(A.sorted + sg, xi, yi, zi, my_idx, my_cell);
    wrap_into_box(box, xi, yi, zi);
    int ic, jc, kc;
    cell_decode(g, my_cell, ic, jc, kc);
    int nbr[14];
    int n = 0;
#pragma unroll 1
    for (int st = 0; st < 27; ++st) {
        const int di = st / 9 - 1, dj = (st / 3) % 3 - 1, dk = st % 3 - 1;
        const int c = cell_linear(g, wrap_cell(ic + di, g.n[0]), wrap_cell(jc + dj, g.n[1]), wrap_cell(kc + dk, g.n[2]));
        if (c < 0) continue;
        const int cb = __ldg(A.cell_start + c), ce = __ldg(A.cell_start + c + 1);
        for (int q = ce - 1; q >= cb; --q) {
            if (q == sg) continue;
            double xj, yj, zj;
            int jdx, jcell;
            load_rec(A.sorted + q, xj, yj, zj, jdx, jcell);
            double dx = xj - xi, dy = yj - yi, dz = zj - zi;
            min_image(box, dx, dy, dz);
            if (dx * dx + dy * dy + dz * dz <= A.rcsq) {
                if (n < 14) nbr[n] = q;
                ++n;
            }
        }
    }
    if (n != 12 && n != 14) return 0;
    for (int a = 0; a < n; ++a) nb[a * STRIDE] = 0;
    for (int a = 0; a < n; ++a) {
        double xa, ya, za;
        int t0, t1;
        load_rec(A.sorted + nbr[a], xa, ya, za, t0, t1);
        for (int b = a + 1; b < n; ++b) {
            double xb, yb, zb;
            load_rec(A.sorted + nbr[b], xb, yb, zb, t0, t1);
            double dx = xb - xa, dy = yb - ya, dz = zb - za;
            min_image(box, dx, dy, dz);
            if (dx * dx + dy * dy + dz * dz <= A.rcsq) {
                nb[a * STRIDE] |= (unsigned short)(1u << b);
                nb[b * STRIDE] |= (unsigned short)(1u << a);
            }
        }
    }
    return cna_label(cna_signatures_smem<STRIDE>(nb, n));
}

template <int NN, int STRIDE>
__device__ __forceinline__ int fused_cna_body(const DBox &box, const SortedAtom *raw, const float4 *f4,
                                              const unsigned short *q, double cutsq, float cut_lo, float cut_hi,
                                              unsigned short *nb)
{
    float rx[NN], ry[NN], rz[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const float4 o = f4[q[a * STRIDE]];
        rx[a] = o.x;
        ry[a] = o.y;
        rz[a] = o.z;
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
    if (ambiguous) return fused_exact_cna<STRIDE>(box, raw, q, NN, cutsq, nb);
    return cna_label(cna_signatures_regs<NN, STRIDE>(rows, nb));
}

template <int T, int TZ, int NT>
__global__ void __launch_bounds__(NT, 3) k_fused_cna(const __grid_constant__ TileArgs A)
{
    constexpr int P = T + 2;
    constexpr int PZ = TZ + 2;
    constexpr int NPEN = P * P;
    constexpr int NCELL = NPEN * PZ;
    constexpr int CSW = PZ + 1;
    constexpr int NW = NT / 32;

    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned o_f4 = (unsigned)A.cap * 32u;                       // [cap] 16 B
    const unsigned o_queue = o_f4 + (unsigned)A.cap * 16u;             // [FUSED_QCAP][NT] u16, interleaved
    const unsigned o_nb = o_queue + FUSED_QCAP * NT * 2u;              // [14][NT] u16 bond rows (flood / exact path)
    const unsigned o_bar = o_nb + 14 * NT * 2u;                        // mbarrier (8 B, 16-byte slot)
    const unsigned o_cs = o_bar + 16u;                                 // [NPEN][CSW] int
    const unsigned o_gstart = o_cs + NPEN * CSW * 4u;                  // [NCELL]
    const unsigned o_ptot = o_gstart + NCELL * 4u;                     // [NPEN + 1]
    const unsigned o_opref = o_ptot + (NPEN + 1) * 4u;                 // [T*T + 1]
    const unsigned o_flag = o_opref + (T * T + 1) * 4u;                // [4]
    SortedAtom *raw = reinterpret_cast<SortedAtom *>(smem);
    float4 *f4 = reinterpret_cast<float4 *>(smem + o_f4);
    unsigned short *queue = reinterpret_cast<unsigned short *>(smem + o_queue);
    unsigned short *nbs = reinterpret_cast<unsigned short *>(smem + o_nb);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + o_bar);
    int *cs = reinterpret_cast<int *>(smem + o_cs);
    int *gstart = reinterpret_cast<int *>(smem + o_gstart);
    int *ptot = reinterpret_cast<int *>(smem + o_ptot);
    int *opref = reinterpret_cast<int *>(smem + o_opref);

    const int tid = threadIdx.x;
    const CellGrid &g = A.g;
    int tx, ty, tz;
    if (A.tile_stride != 1) {   // grid too large for a 3-D launch: linear list of tiles
        int tl = blockIdx.x * max(A.tile_stride, 1) + A.tile_offset;
        if (tl >= A.n_tiles) return;
        tz = tl % A.tiles_z;
        tl /= A.tiles_z;
        ty = tl % A.tiles_y;
        tx = tl / A.tiles_y;
    } else {   // strip order, see k_neighbor_coop
        tz = blockIdx.x;
        const int strip = blockIdx.z / A.tiles_x;
        tx = blockIdx.z - strip * A.tiles_x;
        ty = strip * 8 + blockIdx.y;
        if (ty >= A.tiles_y) return;
    }
    const int u0x = A.p_lo + tx * T - 1, u0y = ty * T - 1, u0z = tz * TZ - 1;
    const int amax = min(T, A.p_hi - (u0x + 1)), bmax = min(T, g.n[1] - (u0y + 1)), kmax = min(TZ, g.n[2] - (u0z + 1));

    if (tid == 0) mbar_init(bar, 1);

    // ---- A. population and global start of every cell of the block (z slots in memory order)
    for (int c = tid; c < NCELL; c += NT) {
        const int ks = c % PZ, b = (c / PZ) % P, a = c / (PZ * P);
        const int kk = PZ - 1 - ks;
        int px;
        const int ux = u0x + a;
        if (A.wrap_x) px = (ux >= A.p_lo - 1 && ux <= A.p_hi) ? map_axis(ux, g.n[0]) : -1;
        else px = (ux >= A.p_lo - 1 && ux <= A.p_hi && ux >= 0 && ux < g.nxl) ? ux : -1;
        const int py = map_axis(u0y + b, g.n[1]);
        const int pz = map_axis(u0z + kk, g.n[2]);
        int beg = -1, cnt = 0;
        if (px >= 0 && py >= 0 && pz >= 0) {
            const int cell = (px * g.n[1] + py) * g.n[2] + (g.n[2] - 1 - pz);
            beg = __ldg(A.cell_start + cell);
            cnt = __ldg(A.cell_start + cell + 1) - beg;
        }
        gstart[c] = beg;
        cs[(a * P + b) * CSW + ks + 1] = cnt;
    }
    __syncthreads();
    for (int p = tid; p < NPEN; p += NT) {
        int s = 0;
        for (int kk = 0; kk < PZ; ++kk) s += cs[p * CSW + kk + 1];
        ptot[p] = s;
    }
    __syncthreads();
    warp0_exclusive_scan(ptot, NPEN);
    __syncthreads();
    for (int p = tid; p < NPEN; p += NT) {
        int off = ptot[p];
        cs[p * CSW] = off;
        for (int kk = 0; kk < PZ; ++kk) {
            off += cs[p * CSW + kk + 1];
            cs[p * CSW + kk + 1] = off;
        }
    }
    __syncthreads();
    for (int i = tid; i < T * T; i += NT) {
        const int a = i / T + 1, b = i % T + 1;
        const int p = a * P + b;
        opref[i] = (a <= amax && b <= bmax) ? cs[p * CSW + PZ - 1] - cs[p * CSW + PZ - 1 - kmax] : 0;
    }
    __syncthreads();
    warp0_exclusive_scan(opref, T * T);
    __syncthreads();
    const int n_staged = ptot[NPEN];
    const int n_owned = opref[T * T];
    if (n_owned == 0) return;
    const bool fits = n_staged <= A.cap;

    const DBox &box = A.box;
    const int warp = tid >> 5, lane = tid & 31;
    auto pencil_of = [&](int t) {
        int lo = 0, hi = T * T;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (opref[mid] <= t) lo = mid;
            else hi = mid;
        }
        return lo;
    };

    if (fits) {
        // ---- B. stage the records (TMA bulk copies, one per run of consecutive global cells of a pencil)
        if (tid == 0) mbar_expect_tx(bar, (unsigned)n_staged * (unsigned)sizeof(SortedAtom));
        __syncthreads();
        for (int p = tid; p < NPEN; p += NT) {
            int kk = 0;
            while (kk < PZ) {
                const int beg = gstart[p * PZ + kk];
                const int dst0 = cs[p * CSW + kk];
                int total = cs[p * CSW + kk + 1] - dst0;
                int k2 = kk + 1;
                if (beg >= 0) {
                    while (k2 < PZ && gstart[p * PZ + k2] == beg + total) {
                        total += cs[p * CSW + k2 + 1] - cs[p * CSW + k2];
                        ++k2;
                    }
                    if (total > 0)
                        bulk_g2s(raw + dst0, A.sorted + beg, (unsigned)total * (unsigned)sizeof(SortedAtom), bar);
                }
                kk = k2;
            }
        }
        if (warp == 0) mbar_wait(bar, 0);   // one warp polls; the others sleep on the CTA barrier
        __syncthreads();

        // ---- C. fp32 positions relative to the tile centre (nearest periodic image), flat over the staged atoms
        const int gx0 = A.wrap_x ? (u0x + 1) : (u0x + 1 + g.x0) % g.n[0];
        const TileCentre ctr = tile_centre(box, g, gx0 + 0.5 * T, u0y + 1 + 0.5 * T, u0z + 1 + 0.5 * TZ);
        for (int s = tid; s < n_staged; s += NT) {
            const double2 lo = reinterpret_cast<const double2 *>(raw + s)[0];
            const double z = reinterpret_cast<const double *>(raw + s)[2];
            double d0, d1, d2;
            rel_image(box, ctr, lo.x, lo.y, z, d0, d1, d2);
            const float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2;
            const float w = __fmaf_rn(f2, f2, __fmaf_rn(f1, f1, f0 * f0));
            // An atom beyond the radius the fp32 bound covers (12.6 rc from the tile centre: an atom far outside
            // the box clamped into an edge cell, the far side of an open axis, NaN) cannot be within rc of an
            // owned atom that is inside it: as a CANDIDATE it never passes; as an OWNED atom it takes the exact
            // fallback below.
            f4[s] = (w <= A.w_limit) ? make_float4(f0, f1, f2, w) : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
        }
        __syncthreads();
    }
    if (!fits) {
        // overflow tile (more atoms than the shared buffers hold): exact walk through global memory
        for (int t = tid; t < n_owned; t += NT) {
            const int lo = pencil_of(t);
            const int p = (lo / T + 1) * P + (lo % T + 1);
            const int sg = gstart[p * PZ + PZ - 1 - kmax] + (t - opref[lo]);
            const int idx = A.sorted[sg].idx;
            if (idx < A.n_rows) A.pattern[idx] = fused_direct_atom<NT>(A, sg, nbs + tid);
        }
        return;
    }

    const unsigned f4_base = smem_u32(f4);
    const unsigned q_base = smem_u32(queue) + 2u * tid;
    constexpr unsigned QS = 2u * NT;
    const unsigned q_full = q_base + QS * FUSED_QCAP;
    const float rc2hi = A.rcsq_hi, rc2lo = A.rcsq_lo;
#pragma unroll 1
    for (int base = warp * 32; base < n_owned; base += NT) {
        const int t = base + lane;
        if (t >= n_owned) continue;
        const int pi = pencil_of(t);
        const int p = (pi / T + 1) * P + (pi % T + 1);
        const int *prow = cs + p * CSW;
        const int s_i = prow[PZ - 1 - kmax] + (t - opref[pi]);
        int kk = PZ - 1 - kmax;   // my memory slot along z: owned slots are PZ-1-kmax .. PZ-2
        while (kk < PZ - 2 && prow[kk + 1] <= s_i) ++kk;
        const int idx = raw[s_i].idx;
        if (idx >= A.n_rows) continue;   // ghost atom of a decomposed frame: neighbour only
        const float4 me = f4[s_i];
        if (!(me.w <= A.w_limit)) {   // far outside the box: exact fallback (see phase C)
            A.pattern[idx] = fused_direct_atom<NT>(A, gstart[p * PZ + PZ - 1 - kmax] + (t - opref[pi]), nbs + tid);
            continue;
        }
        const float fx = -2.0f * me.x, fy = -2.0f * me.y, fz = -2.0f * me.z;
        const float thr_hi = rc2hi - me.w, thr_lo = rc2lo - me.w;
        unsigned q_top = q_base;
        const unsigned self = f4_base + 16u * (unsigned)s_i;

        // the 9 pencils of the stencil; the centre pencil in two pieces that leave the atom itself out.
        // The loop only collects the candidates whose fp32 d^2 is not clearly outside (<= rc^2 (1 + 4e-4));
        // they are classified afterwards, so a candidate costs a load, three FMAs, a compare and a
        // predicated store.  The queue pointer saturates at the last slot: a full queue means more than 14
        // neighbours or a pathological frame, and sends the atom to the exact fallback.
        int pen_off = -P - 1, pen_y = 0;
#pragma unroll 1
        for (int pen = 0; pen < 9; ++pen) {
            const int *row = cs + (p + pen_off) * CSW + kk;
            const unsigned abeg = f4_base + 16u * (unsigned)row[-1];
            const unsigned aend = f4_base + 16u * (unsigned)row[2];
#pragma unroll 1
            for (int piece = 0; piece < (pen == 4 ? 2 : 1); ++piece) {
                unsigned lo_a = abeg, aq = aend;
                if (pen == 4) {
                    if (piece == 0) lo_a = self + 16u;
                    else aq = self;
                }
                float4 o = lds_f4(aq - 16u);   // (a read below the run is harmless: it stays inside the staged arrays)
#pragma unroll 2
                while (aq > lo_a) {
                    aq -= 16u;
                    const float4 c = o;
                    o = lds_f4(aq - 16u);
                    if (__fmaf_rn(fx, c.x, __fmaf_rn(fy, c.y, __fmaf_rn(fz, c.z, c.w))) <= thr_hi) {
                        sts_u16(q_top, aq);
                        q_top = min(q_top + QS, q_full);
                    }
                }
            }
            if (++pen_y == 3) {
                pen_y = 0;
                pen_off += P - 2;
            } else ++pen_off;
        }
        // classify the collected candidates: well inside -> neighbour; inside the band -> exact f64 test
        const int n_hi = (int)((q_top - q_base) / QS);
        if (n_hi >= FUSED_QCAP) {   // queue saturated: exact fallback decides
            A.pattern[idx] = fused_direct_atom<NT>(A, gstart[p * PZ + PZ - 1 - kmax] + (t - opref[pi]), nbs + tid);
            continue;
        }
        int n = 0;
        {
            unsigned short *qw = queue + tid;
            for (int a = 0; a < n_hi; ++a) {
                const int k = (int)(((unsigned)qw[a * NT] - f4_base) & 0xffffu) >> 4;
                const float4 c = f4[k];
                const float t2 = __fmaf_rn(fx, c.x, __fmaf_rn(fy, c.y, __fmaf_rn(fz, c.z, c.w)));
                bool in = t2 <= thr_lo;
                if (!in) in = fused_exact_neighbor(box, raw, s_i, k, A.rcsq);
                if (in) qw[(n++) * NT] = (unsigned short)k;   // compacted in place (n <= a)
            }
        }
        int label = 0;
        const unsigned short *q = queue + tid;
        unsigned short *nb = nbs + tid;
        if (n == 12) label = fused_cna_body<12, NT>(box, raw, f4, q, A.rcsq, A.cut_lo, A.cut_hi, nb);
        else if (n == 14) label = fused_cna_body<14, NT>(box, raw, f4, q, A.rcsq, A.cut_lo, A.cut_hi, nb);
        A.pattern[idx] = label;
    }
}

template <int T, int TZ, int NT> size_t fused_smem_bytes(int cap)
{
    constexpr int P = T + 2, PZ = TZ + 2, NPEN = P * P, NCELL = NPEN * PZ;
    return (size_t)cap * (sizeof(SortedAtom) + sizeof(float4)) + sizeof(unsigned short) * (FUSED_QCAP + 14) * NT +
           sizeof(int) * (NPEN * (PZ + 1) + NCELL + NPEN + 1 + T * T + 1 + 4) + 16 + 64;
}

template <int T, int TZ, int NT> void launch_fused_T(const TileArgs &A, int nblocks, cudaStream_t st)
{
    const size_t smem = fused_smem_bytes<T, TZ, NT>(A.cap);
    // per launch: the attribute belongs to the current device (several devices per process: group.cu)
    CUDA_TRY(cudaFuncSetAttribute(k_fused_cna<T, TZ, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid(nblocks, 1, 1);
    if (A.tile_stride == 1) {
        const int strips = (A.tiles_y + 7) / 8;
        grid = dim3(A.tiles_z, A.tiles_y < 8 ? A.tiles_y : 8, strips * A.tiles_x);
    }
    MDB_LAUNCH((k_fused_cna<T, TZ, NT>), grid, NT, smem, st, A);
}

}  // namespace

// Returns false when the frame is not eligible (triclinic, or too few cells along a periodic axis for
// an unambiguous nearest image); the caller then uses the direct kernel.
bool tiled_neighbor_plan(const MdbSystem &s, int &T)
{
    const CellGrid &g = s.grid;
    {   // triclinic boxes take the tile kernels as well; MDB_TRICLINIC=direct restores the old path
        const char *tenv = getenv("MDB_TRICLINIC");
        if (s.box.triclinic && tenv && !strcmp(tenv, "direct")) return false;
    }
    const double rho = (double)s.N / ((double)g.nxl * g.n[1] * g.n[2]);  // atoms per cell
    // tile shape code: T*16 + TZ.  ~240 owned atoms per 256-thread CTA, <= ~900 staged atoms
    int code = rho <= 0.45 ? (8 * 16 + 8) : (rho <= 1.6 ? (4 * 16 + 8) : (rho <= 3.6 ? (4 * 16 + 6) : (rho <= 8.0 ? (2 * 16 + 4)
                                                                                       : (rho <= 14.0 ? (2 * 16 + 2) : (1 * 16 + 1)))));
    if (rho > 24.0) return false;  // a single cell may exceed the survivor queue: direct kernel
    for (int d = 0; d < 3; ++d) {
        if (!s.box.pbc[d]) continue;
        for (;;) {  // shrink the tile until the nearest-image rule holds along every periodic axis
            const int t = d == 2 ? (code & 15) : (code >> 4);
            if (g.n[d] >= t + 6) break;
            if (code == 1 * 16 + 1) return false;
            code = code == 8 * 16 + 8 ? 4 * 16 + 8 : (code == 4 * 16 + 8 ? 4 * 16 + 6 : (code == 4 * 16 + 6 ? 2 * 16 + 4
                   : (code == 2 * 16 + 4 ? 2 * 16 + 2 : 1 * 16 + 1)));
        }
    }
    const char *env = getenv("MDB_NEIGHBOR");
    if (env && !strcmp(env, "direct")) return false;
    if (const char *tenv = getenv("MDB_TILE")) {   // experiments: force a tile shape (T*16 + TZ)
        const int want = atoi(tenv);
        if (want == 8 * 16 + 8 || want == 4 * 16 + 8 || want == 4 * 16 + 6 || want == 2 * 16 + 4 || want == 2 * 16 + 2 || want == 17) code = want;
    }
    T = code;
    return true;
}

// count_only + sample_stride > 1: estimate of the maximum count from every sample_stride-th tile.
void launch_neighbor_tiled(MdbSystem &s, double rc, int M, int T, bool count_only, int sample_stride)
{
    TileArgs A{};
    const CellGrid &g = s.grid;
    A.sorted = s.sorted.as<SortedAtom>();
    A.cell_start = s.cell_start.as<int>();
    A.box = s.box;
    A.g = g;
    A.rcsq = rc * rc;
    A.pad = rc + 1.0;
    // fp32 error of the dot-product form: <= ~8 ulp of the largest partial sum, i.e. 8 * 6e-8 * w_limit
    // = 7.7e-5 rc^2 at w_limit = 160 rc^2 (a (T+2)-cell block with cells <= 1.34 rc stays below 135 rc^2);
    // the guard of 4e-4 rc^2 leaves a factor 5
    A.rcsq_hi = (float)(rc * rc * (1.0 + 4e-4)) * (1.0f + 1e-6f);
    A.w_limit = (float)(160.0 * rc * rc);
    A.M = M;
    A.n_rows = s.n_rows;
    A.nn = s.nn.ensure<int>(s.n_rows);
    if (!count_only) {
        A.verlet = s.verlet.ensure<int>((size_t)s.n_rows * M);
        A.dist = s.dist.ensure<double>((size_t)s.n_rows * M);
    }
    int *counters = s.counters.ensure<int>(8);
    A.max_count = counters + 6;
    A.min_count = counters + 7;
    const int init[2] = {0, INT_MAX};
    CUDA_TRY(cudaMemcpyAsync(A.max_count, init, sizeof(init), cudaMemcpyHostToDevice, s.stream));
    const bool slab = s.slab_nx > 0;
    A.wrap_x = slab ? 0 : 1;
    A.p_lo = slab ? 1 : 0;
    A.p_hi = slab ? g.nxl - 1 : g.n[0];
    const int TT = T >> 4, TZ = T & 15;
    const int tiles_x = (A.p_hi - A.p_lo + TT - 1) / TT;
    A.tiles_y = (g.n[1] + TT - 1) / TT;
    A.tiles_z = (g.n[2] + TZ - 1) / TZ;
    A.n_tiles = tiles_x * A.tiles_y * A.tiles_z;
    A.tile_stride = sample_stride > 1 ? sample_stride : 1;
    A.tile_offset = 0;
    A.count_only = count_only ? 1 : 0;
    const char *env = getenv("MDB_STAGE");
    A.use_tma = !(env && !strcmp(env, "ldg"));
    A.strip = getenv("MDB_STRIP") != nullptr && !count_only && (M & 3) == 0;
    {   // staged-atom capacity from the mean cell population (tiles above it take the in-kernel direct path)
        const double rho = (double)s.N / ((double)g.nxl * g.n[1] * g.n[2]);
        int cap = (int)(rho * (TT + 2) * (TT + 2) * (TZ + 2) * 1.12) + 32;
        cap = (cap + 15) / 16 * 16;
        A.cap = cap < 256 ? 256 : (cap > 2048 ? 2048 : cap);
    }
    const int nblocks = (A.n_tiles + A.tile_stride - 1) / A.tile_stride;
    if (nblocks <= 0) return;
    A.tiles_x = tiles_x;
    {   // owned-atom capacity of the warp-per-cell kernel's wrapped-position table (slot field: 10 bits)
        const double rho = (double)s.N / ((double)g.nxl * g.n[1] * g.n[2]);
        int ocap = (int)(rho * TT * TT * TZ * 1.15) + 32;
        ocap = (ocap + 31) / 32 * 32;
        A.ocap = ocap > 1024 ? 1024 : ocap;
    }
    // Kernel choice (measured, profiles/r2_neighbor_kernel_study.md): the thread-per-atom kernel above is the
    // faster one while rows are short (FCC / BCC first shells: 20.5 vs 21.9 ms per 99.6 M atoms), the
    // cooperative kernel wins as rows grow (M = 64: 20.9 vs 31.0 ms per 16.4 M atoms) -- dense frames use the
    // small tile shapes.  MDB_NEIGHBOR=tiled_v1 / coop forces one of them.
    const char *kenv = getenv("MDB_NEIGHBOR");
    const bool v1 = A.strip || (kenv ? !strcmp(kenv, "tiled_v1") : TT >= 4);
    if (A.tile_stride == 1 && ((long long)((A.tiles_y + 7) / 8) * tiles_x > 65535 || A.tiles_z > 65535))
        A.tile_stride = 0;   // grid too large for the 3-D strip launch: linear tile list
    if (!v1) {
        static const int nt = getenv("MDB_CELLS_NT") ? atoi(getenv("MDB_CELLS_NT")) : 256;
        switch (T) {
            case 8 * 16 + 8: launch_cells_T<8, 8, 256>(A, nblocks, s.stream); break;
            case 4 * 16 + 8:
                if (nt == 384) launch_cells_T<4, 8, 384>(A, nblocks, s.stream);
                else launch_cells_T<4, 8, 256>(A, nblocks, s.stream);
                break;
            case 4 * 16 + 6:
                if (nt == 384) launch_cells_T<4, 6, 384>(A, nblocks, s.stream);
                else launch_cells_T<4, 6, 256>(A, nblocks, s.stream);
                break;
            case 2 * 16 + 4: launch_cells_T<2, 4, 256>(A, nblocks, s.stream); break;
            case 2 * 16 + 2: launch_cells_T<2, 2, 256>(A, nblocks, s.stream); break;
            default: launch_cells_T<1, 1, 256>(A, nblocks, s.stream); break;
        }
        CUDA_TRY(cudaGetLastError());
        return;
    }
    switch (T) {
        case 8 * 16 + 8: launch_T<8, 8>(A, nblocks, s.stream); break;
        case 4 * 16 + 8: launch_T<4, 8>(A, nblocks, s.stream); break;
        case 4 * 16 + 6: launch_T<4, 6>(A, nblocks, s.stream); break;
        case 2 * 16 + 4: launch_T<2, 4>(A, nblocks, s.stream); break;
        case 2 * 16 + 2: launch_T<2, 2>(A, nblocks, s.stream); break;
        default: launch_T<1, 1>(A, nblocks, s.stream); break;
    }
    CUDA_TRY(cudaGetLastError());
}

int neighbor_tiled_max(MdbSystem &s, int *min_count)
{
    int v[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(v, s.counters.as<int>() + 6, sizeof(v), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (min_count) *min_count = v[1];
    return v[0];
}

// Fused neighbour search + fixed-cutoff CNA on the binned frame; labels into `pattern` (device).  Returns
// false when the frame is not eligible for the tile kernels (the caller then takes the list path).
// *n_fallback receives the number of atoms the kernel could not classify (label -1: overflow tiles).
bool launch_fused_cna(MdbSystem &s, double rc, int *pattern, int *n_fallback)
{
    int T = 0;
    if (!tiled_neighbor_plan(s, T)) return false;
    TileArgs A{};
    const CellGrid &g = s.grid;
    A.sorted = s.sorted.as<SortedAtom>();
    A.cell_start = s.cell_start.as<int>();
    A.box = s.box;
    A.g = g;
    A.rcsq = rc * rc;
    A.rcsq_hi = (float)(rc * rc * (1.0 + 4e-4)) * (1.0f + 1e-6f);
    A.rcsq_lo = (float)(rc * rc * (1.0 - 4e-4)) * (1.0f - 1e-6f);
    A.cut_lo = (float)(rc * rc * (1.0 - 1e-4));
    A.cut_hi = (float)(rc * rc * (1.0 + 1e-4));
    A.w_limit = (float)(160.0 * rc * rc);
    A.n_rows = s.n_rows;
    A.pattern = pattern;
    int *counters = s.counters.ensure<int>(8);
    A.max_count = counters + 6;
    CUDA_TRY(cudaMemsetAsync(A.max_count, 0, sizeof(int), s.stream));
    const bool slab = s.slab_nx > 0;
    A.wrap_x = slab ? 0 : 1;
    A.p_lo = slab ? 1 : 0;
    A.p_hi = slab ? g.nxl - 1 : g.n[0];
    const int TT = T >> 4, TZ = T & 15;
    const int tiles_x = (A.p_hi - A.p_lo + TT - 1) / TT;
    A.tiles_x = tiles_x;
    A.tiles_y = (g.n[1] + TT - 1) / TT;
    A.tiles_z = (g.n[2] + TZ - 1) / TZ;
    A.n_tiles = tiles_x * A.tiles_y * A.tiles_z;
    A.tile_stride = 1;
    if ((long long)((A.tiles_y + 7) / 8) * tiles_x > 65535 || A.tiles_z > 65535) A.tile_stride = 0;
    {
        const double rho = (double)s.N / ((double)g.nxl * g.n[1] * g.n[2]);
        int cap = (int)(rho * (TT + 2) * (TT + 2) * (TZ + 2) * 1.15) + 48;
        cap = (cap + 15) / 16 * 16;
        A.cap = cap < 256 ? 256 : (cap > 2048 ? 2048 : cap);
    }
    if (A.n_tiles <= 0) return true;
    switch (T) {
        case 8 * 16 + 8: launch_fused_T<8, 8, 256>(A, A.n_tiles, s.stream); break;
        case 4 * 16 + 8: launch_fused_T<4, 8, 256>(A, A.n_tiles, s.stream); break;
        case 4 * 16 + 6: launch_fused_T<4, 6, 256>(A, A.n_tiles, s.stream); break;
        case 2 * 16 + 4: launch_fused_T<2, 4, 256>(A, A.n_tiles, s.stream); break;
        case 2 * 16 + 2: launch_fused_T<2, 2, 256>(A, A.n_tiles, s.stream); break;
        default: launch_fused_T<1, 1, 256>(A, A.n_tiles, s.stream); break;
    }
    CUDA_TRY(cudaGetLastError());
    if (n_fallback) {
        CUDA_TRY(cudaMemcpyAsync(n_fallback, A.max_count, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    return true;
}
