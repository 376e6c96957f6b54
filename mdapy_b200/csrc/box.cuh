// mdapy_b200/csrc/box.cuh
//
// Simulation cell shared by every kernel on the path.  Replaces the reference's
// `struct Box` / `get_box` (src/box.h:8-245).  The arithmetic of min_image()
// and wrap() follows the reference operation by operation (src/box.h:94-176)
// because integer outputs downstream (neighbour membership, CNA bonds, RDF
// bins) flip on the last ulp; the whole library is compiled with -fmad=false
// so no multiply-add is contracted behind our back.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MDB_HD __host__ __device__ __forceinline__
#define MDB_COLD __host__ __device__ __noinline__
#else
#define MDB_HD inline
#define MDB_COLD inline
#endif

struct DBox {
    double h[9];      // rows a, b, c                      (Box::data[0..8])
    double hinv[9];   // inverse, same layout as reference (Box::data[9..17])
    double origin[3];
    double thick[3];  // perpendicular thickness per axis
    int pbc[3];
    int triclinic;
    int any_pbc;
    // steps of n(x) = floor(x/L + 0.5) per orthogonal periodic axis, see ortho_image_axis() below
    double img_t[3][4];
    // steps of floor(dx/L) for the wrap, see ortho_wrap_axis()
    double wrap_t[3][4];
};


// ---------------------------------------------------------------------------------------------
// Division-free orthogonal minimum image.
//
// The reference evaluates xij -= L * floor(xij / L + 0.5) (src/box.h:119-124).  n(x) =
// floor(fl(fl(x/L) + 0.5)) is a monotone step function of x, so it is fully described by the
// doubles at which it steps.  The host finds those steps EXACTLY (bisection over the ordered bit
// patterns, evaluating the very same expression), and the device picks n in {-1, 0, +1} with
// comparisons; L*n is then exact and xij - L*n is the same single rounding as the reference's.
// Anything outside [t0, t3) (a neighbour more than 1.5 box lengths away in raw coordinates) takes
// the literal expression.  Bit-identical by construction (verified on 16 M samples per box length
// including +-4 ulp around every step); tests/test_gpu_neighbor.py exercises both branches.
static inline long long ord_bits(double v)
{
    long long b;
    memcpy(&b, &v, 8);
    return b < 0 ? (long long)0x8000000000000000ull - b : b;
}
static inline double ord_double(long long k)
{
    long long b = k < 0 ? (long long)0x8000000000000000ull - k : k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

// smallest double x with floor(x / L + 0.5) >= k
static inline double image_step(double L, int k)
{
    long long lo = ord_bits((k - 1.0) * L), hi = ord_bits((k + 0.5) * L);  // n(lo) < k <= n(hi)
    while ((__int128)hi - lo > 1) {
        const long long mid = (long long)(((__int128)lo + hi) >> 1);  // hi - lo can exceed 2^63
        if (std::floor(ord_double(mid) / L + 0.5) >= k) hi = mid;
        else lo = mid;
    }
    return ord_double(hi);
}

// t[0]: x >= t0 <=> n >= -1,  t[1]: n >= 0,  t[2]: n >= 1,  t[3]: n >= 2
// rarely taken literal forms, kept out of line so the IEEE division sequences are not replicated
// into every inlined call site
static MDB_COLD double image_far(double x, double L) { return x - L * floor(x / L + 0.5); }
static MDB_COLD double wrap_far(double origin, double dx, double L) { return origin + dx - L * floor(dx / L); }

MDB_HD double ortho_image_axis(double x, const double *t, double L)
{
    if (x < t[0] || !(x < t[3])) return image_far(x, L);  // far image (or NaN): literal expression
    // n = -1: x - (L * -1.0) == x + L exactly;  n = 0: x - L*0 == x;  n = +1: x - L
    const double shifted = x + (x < t[1] ? L : -L);
    return (x >= t[1] && x < t[2]) ? x : shifted;
}

// Division-free wrap of dx = x - origin: the reference computes origin + dx - L * floor(dx / L)
// (src/box.h:160-174).  floor(fl(dx / L)) is again a monotone step function of dx; w[k+1] is the
// smallest dx with floor(dx / L) >= k for k = -1, 0, 1, 2 (exact, found by bisection on the host).
static inline double wrap_step(double L, int k)
{
    long long lo = ord_bits((k - 1.5) * L), hi = ord_bits((k + 0.5) * L);  // floor(lo/L) < k <= floor(hi/L)
    while ((__int128)hi - lo > 1) {
        const long long mid = (long long)(((__int128)lo + hi) >> 1);
        if (std::floor(ord_double(mid) / L) >= k) hi = mid;
        else lo = mid;
    }
    return ord_double(hi);
}

MDB_HD double ortho_wrap_axis(double x, double origin, const double *w, double L)
{
    const double dx = x - origin;
    if (dx < w[0] || !(dx < w[3])) return wrap_far(origin, dx, L);
    // n = 0 -> L*0 = 0: origin + dx - 0;  n = -1 -> origin + dx - (-L);  n = 1 -> origin + dx - L
    const double s = origin + dx;
    const double shifted = dx < w[1] ? s + L : s - L;
    return (dx >= w[1] && dx < w[2]) ? s : shifted;
}

// ---- host construction: src/box.h:208-245 (get_box), 182-203, 54-89 --------
static inline double dbox_volume(const DBox &b)
{
    const double *d = b.h;
    if (b.triclinic)
        return d[0] * (d[4] * d[8] - d[5] * d[7]) - d[1] * (d[3] * d[8] - d[5] * d[6]) +
               d[2] * (d[3] * d[7] - d[4] * d[6]);
    return d[0] * d[4] * d[8];
}

// returns 0 on success, 1 if the cell volume is zero (reference throws, box.h:185)
static inline int dbox_make(DBox &b, const double *box9, const double *origin3, const int *boundary3)
{
    b.triclinic = 0;
    for (int i = 0; i < 9; ++i) {
        b.h[i] = box9[i];
        b.hinv[i] = 0.0;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (i != j && std::fabs(b.h[i * 3 + j]) > 1e-10) b.triclinic = 1;
    if (b.h[0] < 0 || b.h[4] < 0 || b.h[8] < 0) b.triclinic = 1;
    if (b.triclinic) {
        const double det = dbox_volume(b);
        if (std::fabs(det) < 1e-12) return 1;
        const double inv_det = 1.0 / det;
        const double *m = b.h;
        b.hinv[0] = (m[4] * m[8] - m[5] * m[7]) * inv_det;
        b.hinv[1] = -(m[1] * m[8] - m[2] * m[7]) * inv_det;
        b.hinv[2] = (m[1] * m[5] - m[2] * m[4]) * inv_det;
        b.hinv[3] = -(m[3] * m[8] - m[5] * m[6]) * inv_det;
        b.hinv[4] = (m[0] * m[8] - m[2] * m[6]) * inv_det;
        b.hinv[5] = -(m[0] * m[5] - m[2] * m[3]) * inv_det;
        b.hinv[6] = (m[3] * m[7] - m[4] * m[6]) * inv_det;
        b.hinv[7] = -(m[0] * m[7] - m[1] * m[6]) * inv_det;
        b.hinv[8] = (m[0] * m[4] - m[1] * m[3]) * inv_det;
    } else {
        b.hinv[0] = 1.0 / b.h[0];
        b.hinv[4] = 1.0 / b.h[4];
        b.hinv[8] = 1.0 / b.h[8];
    }
    b.any_pbc = 0;
    for (int i = 0; i < 3; ++i) {
        b.origin[i] = origin3[i];
        b.pbc[i] = boundary3[i];
        if (boundary3[i]) b.any_pbc = 1;
    }
    for (int dir = 0; dir < 3; ++dir) {
        if (!b.triclinic) {
            b.thick[dir] = b.h[dir * 4];
            continue;
        }
        const double V = dbox_volume(b);
        const double *a = b.h, *bb = b.h + 3, *c = b.h + 6;
        double m, n, k;
        if (dir == 0) {
            m = bb[1] * c[2] - bb[2] * c[1];
            n = bb[2] * c[0] - bb[0] * c[2];
            k = bb[0] * c[1] - bb[1] * c[0];
        } else if (dir == 1) {
            m = a[1] * c[2] - a[2] * c[1];
            n = a[2] * c[0] - a[0] * c[2];
            k = a[0] * c[1] - a[1] * c[0];
        } else {
            m = a[1] * bb[2] - a[2] * bb[1];
            n = a[2] * bb[0] - a[0] * bb[2];
            k = a[0] * bb[1] - a[1] * bb[0];
        }
        b.thick[dir] = V / std::sqrt(m * m + n * n + k * k);
    }
    for (int d = 0; d < 3; ++d)
        for (int k = -1; k <= 2; ++k)
        {
            const bool on = b.pbc[d] && !b.triclinic && b.h[4 * d] > 0;
            b.img_t[d][k + 1] = on ? image_step(b.h[4 * d], k) : 0.0;
            b.wrap_t[d][k + 1] = on ? wrap_step(b.h[4 * d], k) : 0.0;
        }
    return 0;
}

// ---- device arithmetic ------------------------------------------------------
// Minimum image, src/box.h:94-126.
MDB_HD void min_image(const DBox &b, double &xij, double &yij, double &zij)
{
    if (b.triclinic) {
        double x = xij * b.hinv[0] + yij * b.hinv[3] + zij * b.hinv[6];
        double y = xij * b.hinv[1] + yij * b.hinv[4] + zij * b.hinv[7];
        double z = xij * b.hinv[2] + yij * b.hinv[5] + zij * b.hinv[8];
        if (b.pbc[0]) x -= floor(x + 0.5);
        if (b.pbc[1]) y -= floor(y + 0.5);
        if (b.pbc[2]) z -= floor(z + 0.5);
        xij = x * b.h[0] + y * b.h[3] + z * b.h[6];
        yij = x * b.h[1] + y * b.h[4] + z * b.h[7];
        zij = x * b.h[2] + y * b.h[5] + z * b.h[8];
    } else {
        if (b.pbc[0]) xij = ortho_image_axis(xij, b.img_t[0], b.h[0]);
        if (b.pbc[1]) yij = ortho_image_axis(yij, b.img_t[1], b.h[4]);
        if (b.pbc[2]) zij = ortho_image_axis(zij, b.img_t[2], b.h[8]);
    }
}

// Orthogonal-only forms (no triclinic branch in the instruction stream) for kernels that are
// dispatched on orthogonal frames only.
MDB_HD void min_image_ortho(const DBox &b, double &xij, double &yij, double &zij)
{
    if (b.pbc[0]) xij = ortho_image_axis(xij, b.img_t[0], b.h[0]);
    if (b.pbc[1]) yij = ortho_image_axis(yij, b.img_t[1], b.h[4]);
    if (b.pbc[2]) zij = ortho_image_axis(zij, b.img_t[2], b.h[8]);
}

MDB_HD void wrap_ortho(const DBox &b, double &x, double &y, double &z)
{
    if (b.pbc[0]) x = ortho_wrap_axis(x, b.origin[0], b.wrap_t[0], b.h[0]);
    if (b.pbc[1]) y = ortho_wrap_axis(y, b.origin[1], b.wrap_t[1], b.h[4]);
    if (b.pbc[2]) z = ortho_wrap_axis(z, b.origin[2], b.wrap_t[2], b.h[8]);
}

// Wrap into the primary cell, src/box.h:131-176.
MDB_HD void wrap_into_box(const DBox &b, double &x, double &y, double &z)
{
    if (b.triclinic) {
        const double dx = x - b.origin[0];
        const double dy = y - b.origin[1];
        const double dz = z - b.origin[2];
        double nx = dx * b.hinv[0] + dy * b.hinv[3] + dz * b.hinv[6];
        double ny = dx * b.hinv[1] + dy * b.hinv[4] + dz * b.hinv[7];
        double nz = dx * b.hinv[2] + dy * b.hinv[5] + dz * b.hinv[8];
        if (b.pbc[0]) nx -= floor(nx);
        if (b.pbc[1]) ny -= floor(ny);
        if (b.pbc[2]) nz -= floor(nz);
        x = b.origin[0] + nx * b.h[0] + ny * b.h[3] + nz * b.h[6];
        y = b.origin[1] + nx * b.h[1] + ny * b.h[4] + nz * b.h[7];
        z = b.origin[2] + nx * b.h[2] + ny * b.h[5] + nz * b.h[8];
    } else {
        if (b.pbc[0]) x = ortho_wrap_axis(x, b.origin[0], b.wrap_t[0], b.h[0]);
        if (b.pbc[1]) y = ortho_wrap_axis(y, b.origin[1], b.wrap_t[1], b.h[4]);
        if (b.pbc[2]) z = ortho_wrap_axis(z, b.origin[2], b.wrap_t[2], b.h[8]);
    }
}

// Squared min-image distance between raw positions, src/cna.cpp:149-161.
MDB_HD double pbc_dist_sq(const DBox &b, double xi, double yi, double zi, double xj, double yj, double zj)
{
    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    min_image(b, dx, dy, dz);
    return dx * dx + dy * dy + dz * dz;
}

// Cell grid of the cut-off search, src/neighbor.cpp:30-62 and 367-370.
// x0/nxl describe the window of x-planes that is actually stored: the whole grid on one
// GPU (x0 = 0, nxl = n[0]); an owned slab plus one ghost plane on each side when the frame
// is decomposed across GPUs (mdapy_b200/distributed.py).  Cell geometry is always GLOBAL.
struct CellGrid {
    int n[3];
    int total;   // stored cells = nxl * n[1] * n[2]
    double rc_inv;
    int x0;
    int nxl;
};

static inline CellGrid cellgrid_make(const DBox &b, double rc)
{
    CellGrid g;
    for (int i = 0; i < 3; ++i) {
        int c = static_cast<int>(std::floor(b.thick[i] / rc));
        g.n[i] = c > 3 ? c : 3;
    }
    g.total = g.n[0] * g.n[1] * g.n[2];
    g.rc_inv = 1.0 / rc;
    g.x0 = 0;
    g.nxl = g.n[0];
    return g;
}

MDB_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// position must already be wrapped (neighbor.cpp:87-94)
MDB_HD void cell_of(const DBox &b, const CellGrid &g, double x, double y, double z, int &ic, int &jc, int &kc)
{
    if (b.triclinic) {
        const double dx = x - b.origin[0];
        const double dy = y - b.origin[1];
        const double dz = z - b.origin[2];
        const double nx = dx * b.hinv[0] + dy * b.hinv[3] + dz * b.hinv[6];
        const double ny = dx * b.hinv[1] + dy * b.hinv[4] + dz * b.hinv[7];
        const double nz = dx * b.hinv[2] + dy * b.hinv[5] + dz * b.hinv[8];
        ic = static_cast<int>(floor(nx * b.thick[0] * g.rc_inv));
        jc = static_cast<int>(floor(ny * b.thick[1] * g.rc_inv));
        kc = static_cast<int>(floor(nz * b.thick[2] * g.rc_inv));
    } else {
        ic = static_cast<int>(floor((x - b.origin[0]) * g.rc_inv));
        jc = static_cast<int>(floor((y - b.origin[1]) * g.rc_inv));
        kc = static_cast<int>(floor((z - b.origin[2]) * g.rc_inv));
    }
    ic = clampi(ic, 0, g.n[0] - 1);
    jc = clampi(jc, 0, g.n[1] - 1);
    kc = clampi(kc, 0, g.n[2] - 1);
}

MDB_HD int wrap_cell(int a, int n)
{
    int r = a % n;
    return r < 0 ? r + n : r;
}

// Stored linear id of global cell (ci, cj, ck), ci already wrapped into [0, n0); -1 outside the
// window.  Inside a z-pencil the cells are stored in DESCENDING ck: together with ascending original
// index inside a cell, ONE backward walk over the three consecutive cells ck-1, ck, ck+1 visits the
// atoms exactly in the reference's order (cells ck-1, ck, ck+1, each in descending index -- the
// head-insertion chains of src/neighbor.cpp:97-98 walked by the loops of 147-181).
MDB_HD int cell_linear(const CellGrid &g, int ci, int cj, int ck)
{
    int p = ci - g.x0;
    if (p < 0) p += g.n[0];
    if (p >= g.nxl) return -1;
    return (p * g.n[1] + cj) * g.n[2] + (g.n[2] - 1 - ck);
}

// inverse of cell_linear: global x plane, y and z cell of a stored cell id
MDB_HD void cell_decode(const CellGrid &g, int cell, int &ic, int &jc, int &kc)
{
    kc = g.n[2] - 1 - cell % g.n[2];
    jc = (cell / g.n[2]) % g.n[1];
    ic = wrap_cell(cell / (g.n[2] * g.n[1]) + g.x0, g.n[0]);
}

