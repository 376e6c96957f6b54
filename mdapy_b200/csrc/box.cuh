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

#if defined(__CUDACC__)
#define MDB_HD __host__ __device__ __forceinline__
#else
#define MDB_HD inline
#endif

struct DBox {
    double h[9];      // rows a, b, c                      (Box::data[0..8])
    double hinv[9];   // inverse, same layout as reference (Box::data[9..17])
    double origin[3];
    double thick[3];  // perpendicular thickness per axis
    int pbc[3];
    int triclinic;
    int any_pbc;
};

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
        if (b.pbc[0]) xij -= b.h[0] * floor(xij / b.h[0] + 0.5);
        if (b.pbc[1]) yij -= b.h[4] * floor(yij / b.h[4] + 0.5);
        if (b.pbc[2]) zij -= b.h[8] * floor(zij / b.h[8] + 0.5);
    }
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
        if (b.pbc[0]) {
            const double dx = x - b.origin[0];
            x = b.origin[0] + dx - b.h[0] * floor(dx / b.h[0]);
        }
        if (b.pbc[1]) {
            const double dy = y - b.origin[1];
            y = b.origin[1] + dy - b.h[4] * floor(dy / b.h[4]);
        }
        if (b.pbc[2]) {
            const double dz = z - b.origin[2];
            z = b.origin[2] + dz - b.h[8] * floor(dz / b.h[8]);
        }
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

// stored linear id of global cell (ci, cj, ck), ci already wrapped into [0, n0); -1 outside the window
MDB_HD int cell_linear(const CellGrid &g, int ci, int cj, int ck)
{
    int p = ci - g.x0;
    if (p < 0) p += g.n[0];
    if (p >= g.nxl) return -1;
    return (p * g.n[1] + cj) * g.n[2] + ck;
}
