/* oracle/port/mdapy_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithms on the hot path, written
 * from the reference's behaviour (file:line cited per function), NOT copied:
 * it is the checker that is always buildable (no /root/reference needed) and
 * is itself pinned against the compiled reference (oracle/_ref) and the
 * committed golden vectors by tests/test_oracle_*.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it; the
 * product (mdapy_b200) never does.
 *
 * Built with -ffp-contract=off: the reference build has no FMA contraction.
 * PTM is not restated here (9 kLoC vendored library); its checker is
 * oracle/_ref plus the golden fixtures under tests/golden/.
 */
#include "mdapy_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ cell (src/box.h:8-245) */
typedef struct {
    double h[9], hi[9], o[3], thick[3];
    int pbc[3], tri, anyp;
} cell_t;

static double cell_volume(const cell_t *c)
{
    const double *d = c->h;
    if (c->tri)
        return d[0] * (d[4] * d[8] - d[5] * d[7]) - d[1] * (d[3] * d[8] - d[5] * d[6]) +
               d[2] * (d[3] * d[7] - d[4] * d[6]);
    return d[0] * d[4] * d[8];
}

/* box.h:208-245 get_box, 182-203 inverse, 54-89 thickness */
static void cell_init(cell_t *c, const double *box9, const double *origin3, const int *boundary3)
{
    memset(c, 0, sizeof(*c));
    memcpy(c->h, box9, 9 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (i != j && fabs(c->h[3 * i + j]) > 1e-10) c->tri = 1;
    if (c->h[0] < 0 || c->h[4] < 0 || c->h[8] < 0) c->tri = 1;
    if (c->tri) {
        const double *m = c->h;
        const double inv = 1.0 / cell_volume(c);
        c->hi[0] = (m[4] * m[8] - m[5] * m[7]) * inv;
        c->hi[1] = -(m[1] * m[8] - m[2] * m[7]) * inv;
        c->hi[2] = (m[1] * m[5] - m[2] * m[4]) * inv;
        c->hi[3] = -(m[3] * m[8] - m[5] * m[6]) * inv;
        c->hi[4] = (m[0] * m[8] - m[2] * m[6]) * inv;
        c->hi[5] = -(m[0] * m[5] - m[2] * m[3]) * inv;
        c->hi[6] = (m[3] * m[7] - m[4] * m[6]) * inv;
        c->hi[7] = -(m[0] * m[7] - m[1] * m[6]) * inv;
        c->hi[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
    } else {
        c->hi[0] = 1.0 / c->h[0];
        c->hi[4] = 1.0 / c->h[4];
        c->hi[8] = 1.0 / c->h[8];
    }
    for (int i = 0; i < 3; ++i) {
        c->o[i] = origin3[i];
        c->pbc[i] = boundary3[i];
        if (boundary3[i]) c->anyp = 1;
    }
    for (int d = 0; d < 3; ++d) {
        if (!c->tri) {
            c->thick[d] = c->h[4 * d];
            continue;
        }
        const double *u = c->h + 3 * ((d + 1) % 3), *v = c->h + 3 * ((d + 2) % 3);
        /* the reference orders the cross products per direction (box.h:65-82); the sign
         * differs for dir 1 but only the squared norm enters */
        double m, n, k;
        if (d == 1) {
            const double *a = c->h, *cc = c->h + 6;
            m = a[1] * cc[2] - a[2] * cc[1];
            n = a[2] * cc[0] - a[0] * cc[2];
            k = a[0] * cc[1] - a[1] * cc[0];
        } else {
            m = u[1] * v[2] - u[2] * v[1];
            n = u[2] * v[0] - u[0] * v[2];
            k = u[0] * v[1] - u[1] * v[0];
        }
        c->thick[d] = cell_volume(c) / sqrt(m * m + n * n + k * k);
    }
}

/* box.h:94-126 */
static void min_image(const cell_t *c, double *dx, double *dy, double *dz)
{
    if (c->tri) {
        double a = *dx * c->hi[0] + *dy * c->hi[3] + *dz * c->hi[6];
        double b = *dx * c->hi[1] + *dy * c->hi[4] + *dz * c->hi[7];
        double g = *dx * c->hi[2] + *dy * c->hi[5] + *dz * c->hi[8];
        if (c->pbc[0]) a -= floor(a + 0.5);
        if (c->pbc[1]) b -= floor(b + 0.5);
        if (c->pbc[2]) g -= floor(g + 0.5);
        *dx = a * c->h[0] + b * c->h[3] + g * c->h[6];
        *dy = a * c->h[1] + b * c->h[4] + g * c->h[7];
        *dz = a * c->h[2] + b * c->h[5] + g * c->h[8];
    } else {
        if (c->pbc[0]) *dx -= c->h[0] * floor(*dx / c->h[0] + 0.5);
        if (c->pbc[1]) *dy -= c->h[4] * floor(*dy / c->h[4] + 0.5);
        if (c->pbc[2]) *dz -= c->h[8] * floor(*dz / c->h[8] + 0.5);
    }
}

/* box.h:131-176 */
static void wrap_point(const cell_t *c, double *x, double *y, double *z)
{
    if (c->tri) {
        const double dx = *x - c->o[0], dy = *y - c->o[1], dz = *z - c->o[2];
        double a = dx * c->hi[0] + dy * c->hi[3] + dz * c->hi[6];
        double b = dx * c->hi[1] + dy * c->hi[4] + dz * c->hi[7];
        double g = dx * c->hi[2] + dy * c->hi[5] + dz * c->hi[8];
        if (c->pbc[0]) a -= floor(a);
        if (c->pbc[1]) b -= floor(b);
        if (c->pbc[2]) g -= floor(g);
        *x = c->o[0] + a * c->h[0] + b * c->h[3] + g * c->h[6];
        *y = c->o[1] + a * c->h[1] + b * c->h[4] + g * c->h[7];
        *z = c->o[2] + a * c->h[2] + b * c->h[5] + g * c->h[8];
    } else {
        double *p[3] = {x, y, z};
        for (int d = 0; d < 3; ++d)
            if (c->pbc[d]) {
                const double t = *p[d] - c->o[d];
                *p[d] = c->o[d] + t - c->h[4 * d] * floor(t / c->h[4 * d]);
            }
    }
}

static double raw_dist_sq(const cell_t *c, const double *x, const double *y, const double *z, int i, int j)
{ /* cna.cpp:149-161 */
    double dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
    min_image(c, &dx, &dy, &dz);
    return dx * dx + dy * dy + dz * dz;
}

static int imod(int a, int n)
{
    int r = a % n;
    return r < 0 ? r + n : r;
}

/* ------------------------------------------------------------------ cut-off neighbours
 * neighbor.cpp:30-62 cell index, 64-100 build_cell, 102-187 build_verlet_list, 351-388.
 * The reference pushes atoms at the head of per-cell chains, so a chain is walked in
 * DESCENDING atom index; here cells are a counting sort and walked backwards. */
static void cell_index(const cell_t *c, double rinv, const int *nc, double x, double y, double z, int *out)
{
    double f[3];
    if (c->tri) {
        const double dx = x - c->o[0], dy = y - c->o[1], dz = z - c->o[2];
        f[0] = (dx * c->hi[0] + dy * c->hi[3] + dz * c->hi[6]) * c->thick[0] * rinv;
        f[1] = (dx * c->hi[1] + dy * c->hi[4] + dz * c->hi[7]) * c->thick[1] * rinv;
        f[2] = (dx * c->hi[2] + dy * c->hi[5] + dz * c->hi[8]) * c->thick[2] * rinv;
    } else {
        f[0] = (x - c->o[0]) * rinv;
        f[1] = (y - c->o[1]) * rinv;
        f[2] = (z - c->o[2]) * rinv;
    }
    for (int d = 0; d < 3; ++d) {
        int v = (int)floor(f[d]);
        if (v > nc[d] - 1) v = nc[d] - 1;
        if (v < 0) v = 0;
        out[d] = v;
    }
}

void port_build_neighbor(const double *x, const double *y, const double *z, int N, const double *box9,
                         const double *origin3, const int *boundary3, double rc, int *verlet, double *dist, int *nn,
                         int M, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    int nc[3];
    for (int d = 0; d < 3; ++d) {
        nc[d] = (int)floor(c.thick[d] / rc);
        if (nc[d] < 3) nc[d] = 3;
    }
    const int ncell = nc[0] * nc[1] * nc[2];
    const double rinv = 1.0 / rc, rcsq = rc * rc;
    int *start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    int *cell = (int *)malloc(sizeof(int) * (size_t)N);
    int *order = (int *)malloc(sizeof(int) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        double xi = x[i], yi = y[i], zi = z[i];
        if (c.anyp) wrap_point(&c, &xi, &yi, &zi);
        int ci[3];
        cell_index(&c, rinv, nc, xi, yi, zi, ci);
        cell[i] = (ci[0] * nc[1] + ci[1]) * nc[2] + ci[2];
        ++start[cell[i] + 1];
    }
    for (int k = 0; k < ncell; ++k) start[k + 1] += start[k];
    int *fill = (int *)calloc((size_t)ncell, sizeof(int));
    for (int i = 0; i < N; ++i) order[start[cell[i]] + fill[cell[i]]++] = i; /* ascending index per cell */
    free(fill);

#pragma omp parallel for num_threads(num_t) schedule(dynamic, 256)
    for (int i = 0; i < N; ++i) {
        double xi = x[i], yi = y[i], zi = z[i];
        if (c.anyp) wrap_point(&c, &xi, &yi, &zi);
        int ci[3];
        cell_index(&c, rinv, nc, xi, yi, zi, ci);
        int cnt = 0;
        for (int a = ci[0] - 1; a <= ci[0] + 1; ++a)
            for (int b = ci[1] - 1; b <= ci[1] + 1; ++b)
                for (int g = ci[2] - 1; g <= ci[2] + 1; ++g) {
                    const int cc = (imod(a, nc[0]) * nc[1] + imod(b, nc[1])) * nc[2] + imod(g, nc[2]);
                    for (int q = start[cc + 1] - 1; q >= start[cc]; --q) {
                        const int j = order[q];
                        if (j == i) continue;
                        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
                        min_image(&c, &dx, &dy, &dz);
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 <= rcsq) {
                            if (cnt < M && verlet) {
                                verlet[(size_t)i * M + cnt] = j;
                                dist[(size_t)i * M + cnt] = sqrt(d2);
                            }
                            ++cnt;
                        }
                    }
                }
        nn[i] = cnt;
    }
    free(start);
    free(cell);
    free(order);
}

/* neighbor.cpp:745-778: selection sort of the first k slots, scanning the whole row */
void port_sort_verlet_by_distance(int *verlet, double *dist, int N, int M, int k, int num_t)
{
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        int *v = verlet + (size_t)i * M;
        double *d = dist + (size_t)i * M;
        const int eff = k < M ? k : M;
        for (int a = 0; a < eff; ++a) {
            int best = a;
            for (int b = a + 1; b < M; ++b)
                if (d[b] < d[best]) best = b;
            if (best != a) {
                double td = d[a];
                d[a] = d[best];
                d[best] = td;
                int tv = v[a];
                v[a] = v[best];
                v[best] = tv;
            }
        }
    }
}

/* ------------------------------------------------------------------ k nearest neighbours
 * fast_knn.cpp:846-916 semantics by exhaustive search over (atom, image shift):
 * wrap 86-99 / 682-703, shifts 801-841, distance 600-605 / 534-537, self rule 641,
 * ascending output with sqrt, -1/-1.0 padding 879-886.  O(N^2 * shifts): small cases only. */
void port_knn(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, int k, int *indices, double *distances, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    double *w = (double *)malloc(sizeof(double) * 3 * (size_t)N);
    for (int i = 0; i < N; ++i) {
        double p[3] = {x[i], y[i], z[i]};
        if (c.tri) {
            double r[3];
            r[0] = p[0] * c.hi[0] + p[1] * c.hi[3] + p[2] * c.hi[6];
            r[1] = p[0] * c.hi[1] + p[1] * c.hi[4] + p[2] * c.hi[7];
            r[2] = p[0] * c.hi[2] + p[1] * c.hi[5] + p[2] * c.hi[8];
            for (int d = 0; d < 3; ++d)
                if (c.pbc[d]) {
                    const double s = floor(r[d]);
                    if (s != 0.0) {
                        p[0] -= s * c.h[3 * d];
                        p[1] -= s * c.h[3 * d + 1];
                        p[2] -= s * c.h[3 * d + 2];
                    }
                }
        } else {
            for (int d = 0; d < 3; ++d)
                if (c.pbc[d]) {
                    const double s = floor((p[d] - c.o[d]) * c.hi[4 * d]);
                    if (s != 0.0) p[d] -= s * c.h[4 * d];
                }
        }
        w[3 * i] = p[0];
        w[3 * i + 1] = p[1];
        w[3 * i + 2] = p[2];
    }
    int nim = 1;
    if (c.anyp) {
        long cl = N < 50 ? 50 : (N > 200 ? 200 : N);
        nim = (int)(200 / cl);
        if (nim < 1) nim = 1;
        if (nim < 2 && c.tri) nim = 2;
    }
    const int ex[3] = {c.pbc[0] ? nim : 0, c.pbc[1] ? nim : 0, c.pbc[2] ? nim : 0};
    const int ns = (2 * ex[0] + 1) * (2 * ex[1] + 1) * (2 * ex[2] + 1);
    double *sh = (double *)malloc(sizeof(double) * 3 * (size_t)ns);
    int t = 0;
    for (int iz = -ex[2]; iz <= ex[2]; ++iz)
        for (int iy = -ex[1]; iy <= ex[1]; ++iy)
            for (int ix = -ex[0]; ix <= ex[0]; ++ix, ++t) {
                if (c.tri) {
                    sh[3 * t] = ix * c.h[0] + iy * c.h[3] + iz * c.h[6];
                    sh[3 * t + 1] = ix * c.h[1] + iy * c.h[4] + iz * c.h[7];
                    sh[3 * t + 2] = ix * c.h[2] + iy * c.h[5] + iz * c.h[8];
                } else {
                    sh[3 * t] = ix * c.h[0];
                    sh[3 * t + 1] = iy * c.h[4];
                    sh[3 * t + 2] = iz * c.h[8];
                }
            }
#pragma omp parallel for num_threads(num_t) schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        double bd[64];
        int bi[64], nb = 0;
        for (int s = 0; s < ns; ++s) {
            const double qx = w[3 * i] - sh[3 * s], qy = w[3 * i + 1] - sh[3 * s + 1], qz = w[3 * i + 2] - sh[3 * s + 2];
            for (int j = 0; j < N; ++j) {
                const double dx = w[3 * j] - qx, dy = w[3 * j + 1] - qy, dz = w[3 * j + 2] - qz;
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (j == i && d2 == 0.0) continue;
                if (nb == k && !(d2 < bd[k - 1])) continue;
                int pos = nb < k ? nb : k - 1;
                while (pos > 0 && bd[pos - 1] > d2) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d2;
                bi[pos] = j;
                if (nb < k) ++nb;
            }
        }
        for (int q = 0; q < k; ++q) {
            indices[(size_t)i * k + q] = q < nb ? bi[q] : -1;
            distances[(size_t)i * k + q] = q < nb ? sqrt(bd[q]) : -1.0;
        }
    }
    free(w);
    free(sh);
}

/* ------------------------------------------------------------------ common neighbour analysis
 * cna.cpp:16-147 signature helpers, 289-427 adaptive, 429-506 fixed.  The reference's
 * "max chain length" is the bond count of the largest connected component of the
 * common-neighbour bond graph. */
typedef struct {
    int n421, n422, n555, n444, n666;
} sig_t;

static int popc(unsigned v) { return __builtin_popcount(v); }

static sig_t signatures(const unsigned *nb, int nn)
{
    sig_t s = {0, 0, 0, 0, 0};
    for (int ni = 0; ni < nn; ++ni) {
        const unsigned common = nb[ni];
        const int ncommon = popc(common);
        int twice = 0;
        for (int v = 0; v < nn; ++v)
            if (common >> v & 1u) twice += popc(nb[v] & common);
        const int nbonds = twice / 2;
        int longest = 0;
        unsigned left = common;
        while (left) {
            unsigned comp = left & (~left + 1u), front = comp;
            while (front) {
                unsigned nxt = 0;
                for (int v = 0; v < nn; ++v)
                    if (front >> v & 1u) nxt |= nb[v] & common;
                nxt &= ~comp;
                comp |= nxt;
                front = nxt;
            }
            int e2 = 0;
            for (int v = 0; v < nn; ++v)
                if (comp >> v & 1u) e2 += popc(nb[v] & common);
            if (e2 / 2 > longest) longest = e2 / 2;
            left &= ~comp;
        }
        if (ncommon == 4 && nbonds == 2) {
            if (longest == 1) ++s.n421;
            else if (longest == 2) ++s.n422;
        } else if (ncommon == 5 && nbonds == 5 && longest == 5) ++s.n555;
        else if (ncommon == 4 && nbonds == 4 && longest == 4) ++s.n444;
        else if (ncommon == 6 && nbonds == 6 && longest == 6) ++s.n666;
    }
    return s;
}

static void bonds(const cell_t *c, const double *x, const double *y, const double *z, const int *row, int nn,
                  double cutsq, unsigned *nb)
{
    for (int a = 0; a < nn; ++a) nb[a] = 0;
    for (int a = 0; a < nn; ++a)
        for (int b = a + 1; b < nn; ++b)
            if (raw_dist_sq(c, x, y, z, row[a], row[b]) <= cutsq) {
                nb[a] |= 1u << b;
                nb[b] |= 1u << a;
            }
}

void port_fcna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, const int *nn, int *pattern, double rc, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const double cutsq = rc * rc;
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        const int n = nn[i];
        if (n != 12 && n != 14) continue;
        unsigned nb[32];
        bonds(&c, x, y, z, verlet + (size_t)i * M, n, cutsq, nb);
        const sig_t s = signatures(nb, n);
        if (s.n421 == 12) pattern[i] = 1;
        else if (s.n421 == 6 && s.n422 == 6) pattern[i] = 2;
        else if (s.n555 == 12) pattern[i] = 4;
        else if (s.n666 == 8 && s.n444 == 6) pattern[i] = 3;
    }
}

void port_acna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, int *pattern, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        const int *row = verlet + (size_t)i * M;
        unsigned nb[32];
        double sum = 0.0;
        for (int m = 0; m < 12; ++m) sum += sqrt(raw_dist_sq(&c, x, y, z, i, row[m]));
        double cut = sum / 12 * (1.0 + sqrt(2.0)) * 0.5; /* cna.cpp:320 */
        bonds(&c, x, y, z, row, 12, cut * cut, nb);
        sig_t s = signatures(nb, 12);
        int p = 0;
        if (s.n421 == 12) p = 1;
        else if (s.n421 == 6 && s.n422 == 6) p = 2;
        else if (s.n555 == 12) p = 4;
        if (!p) {
            sum = 0.0;
            for (int m = 0; m < 8; ++m) sum += sqrt(raw_dist_sq(&c, x, y, z, i, row[m]) / (3.0 / 4.0));
            for (int m = 8; m < 14; ++m) sum += sqrt(raw_dist_sq(&c, x, y, z, i, row[m]));
            cut = sum / 14 * (1.0 + sqrt(2.0)) * 0.5; /* cna.cpp:387 */
            bonds(&c, x, y, z, row, 14, cut * cut, nb);
            s = signatures(nb, 14);
            if (s.n666 == 8 && s.n444 == 6) p = 3;
        }
        if (p) pattern[i] = p;
    }
}

/* ------------------------------------------------------------------ diamond structure
 * cna.cpp:163-287 IdentifyDiamond.  verlet rows hold >= 4 neighbours sorted by distance.  Second-shell
 * list = for each of the 4 first neighbours j, the first 3 entries of j's row that are not i; CNA on
 * those 12 with cutoff 1.2071068 * mean distance; then two serial sweeps spread the labels to first
 * and second neighbours (lowest atom index wins, cna.cpp:251-286). */
void port_ids(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, int *new_verlet, int *pattern, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        int *second = new_verlet + (size_t)i * 12;
        int count = 0;
        for (int m = 0; m < 4; ++m) {
            const int j = verlet[(size_t)i * M + m];
            int taken = 0;
            for (int kk = 0; kk < 4; ++kk) {
                const int k = verlet[(size_t)j * M + kk];
                if (k != i && taken < 3) {
                    second[count++] = k;
                    ++taken;
                }
            }
        }
        double sum = 0.0;
        for (int m = 0; m < 12; ++m) sum += sqrt(raw_dist_sq(&c, x, y, z, i, second[m]));
        sum /= 12.0;
        const double cut = sum * 1.2071068; /* cna.cpp:208 */
        unsigned nb[32];
        bonds(&c, x, y, z, second, 12, cut * cut, nb);
        sig_t s = signatures(nb, 12);
        if (s.n421 == 12) pattern[i] = 1;
        else if (s.n421 == 6 && s.n422 == 6) pattern[i] = 4;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int a = pass ? 2 : 1, b = pass ? 5 : 4;
        for (int i = 0; i < N; ++i) {
            const int t = pattern[i];
            if (t != a && t != b) continue;
            for (int jj = 0; jj < 4; ++jj) {
                const int j = verlet[(size_t)i * M + jj];
                if (pattern[j] == 0) pattern[j] = t + 1;
            }
        }
    }
}

/* ------------------------------------------------------------------ centro-symmetry
 * centro_symmetry_parameter.cpp:12-92: all pair sums |r_j + r_k|^2, the nnei/2 smallest
 * added in ascending order. */
static int cmp_double(const void *a, const void *b)
{
    const double u = *(const double *)a, v = *(const double *)b;
    return (u > v) - (u < v);
}

void port_csp(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, int nnei, double *csp, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const int npair = nnei * (nnei - 1) / 2;
#pragma omp parallel num_threads(num_t)
    {
        double *v = (double *)malloc(sizeof(double) * (size_t)(npair > 0 ? npair : 1));
        double *r = (double *)malloc(sizeof(double) * 3 * (size_t)nnei);
#pragma omp for
        for (int i = 0; i < N; ++i) {
            for (int a = 0; a < nnei; ++a) {
                const int j = verlet[(size_t)i * M + a];
                r[3 * a] = x[j] - x[i];
                r[3 * a + 1] = y[j] - y[i];
                r[3 * a + 2] = z[j] - z[i];
                min_image(&c, r + 3 * a, r + 3 * a + 1, r + 3 * a + 2);
            }
            int t = 0;
            for (int a = 0; a < nnei; ++a)
                for (int b = a + 1; b < nnei; ++b, ++t) {
                    const double sx = r[3 * a] + r[3 * b], sy = r[3 * a + 1] + r[3 * b + 1],
                                 sz = r[3 * a + 2] + r[3 * b + 2];
                    v[t] = sx * sx + sy * sy + sz * sz;
                }
            qsort(v, (size_t)npair, sizeof(double), cmp_double);
            double s = 0.0;
            for (int q = 0; q < nnei / 2; ++q) s += v[q];
            csp[i] = s;
        }
        free(v);
        free(r);
    }
}

/* ------------------------------------------------------------------ Ackland-Jones
 * ackland_jones_analysis.cpp:9-172 */
void port_aja(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, const double *dist, int Md, int *aja, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    static const double edge[7] = {-0.945, -0.915, -0.755, -0.195, 0.195, 0.245, 0.795};
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        const double *d = dist + (size_t)i * Md;
        const int *v = verlet + (size_t)i * M;
        double r0 = 0.0;
        for (int j = 0; j < 6; ++j) r0 += d[j] * d[j];
        r0 /= 6.0;
        const double lim0 = 1.45 * r0, lim1 = 1.55 * r0;
        int n0 = 0, n1 = 0;
        for (int j = 0; j < 14; ++j) {
            const double r2 = d[j] * d[j];
            if (r2 < lim1) {
                ++n1;
                if (r2 < lim0) ++n0;
            }
        }
        double r[14][3];
        for (int j = 0; j < n0; ++j) {
            r[j][0] = x[v[j]] - x[i];
            r[j][1] = y[v[j]] - y[i];
            r[j][2] = z[v[j]] - z[i];
            min_image(&c, &r[j][0], &r[j][1], &r[j][2]);
        }
        int al[8] = {0};
        for (int j = 0; j < n0; ++j)
            for (int k = j + 1; k < n0; ++k) {
                const double dot = r[j][0] * r[k][0] + r[j][1] * r[k][1] + r[j][2] * r[k][2];
                const double ct = dot / (d[j] * d[k]);
                int b = 0;
                while (b < 7 && !(ct < edge[b])) ++b;
                ++al[b];
            }
        const double s_cp = fabs(1.0 - al[6] / 24.0);
        const int s56m4 = al[5] + al[6] - al[4];
        double s_bcc = s_cp + 1.0;
        if (s56m4 != 0) s_bcc = 0.35 * al[4] / (double)s56m4;
        double s_fcc = 0.61 * (abs(al[0] + al[1] - 6) + al[2]) / 6.0;
        double s_hcp = (fabs(al[0] - 3.0) + abs(al[0] + al[1] + al[2] + al[3] - 9)) / 12.0;
        if (al[0] == 7) s_bcc = 0.0;
        else if (al[0] == 6) s_fcc = 0.0;
        else if (al[0] <= 3) s_hcp = 0.0;
        int t;
        if (al[7] > 0) t = 0;
        else if (al[4] < 3) t = (n1 > 13 || n1 < 11) ? 0 : 4;
        else if (s_bcc <= s_cp) t = n1 < 11 ? 0 : 3;
        else if (n1 > 12 || n1 < 11) t = 0;
        else t = s_fcc < s_hcp ? 1 : 2;
        aja[i] = t;
    }
}

/* ------------------------------------------------------------------ Steinhardt
 * steinhardt_bond_orientation.cpp:188-224 CG table, 243-286 Legendre / prefactor,
 * 288-576 _compute_ql, 578-675 identifySolidLiquid. */
/* The reference tabulates n! as 15-significant-digit decimal literals (cpp:12-181).  For
 * n <= 78 every literal equals strtod("%.15g" of n!) -- checked against the table by
 * tests/test_oracle_port.py through w_l parity with oracle/_ref -- which covers l <= 25
 * (3l+1 <= 78); beyond that a few literals are truncated instead of rounded and the
 * Clebsch-Gordan coefficients agree only to ~1e-15 relative. */
#include <stdio.h>
static double fact(int n)
{
    long double f = 1.0L;
    for (int i = 2; i <= n; ++i) f *= i;
    char buf[64];
    snprintf(buf, sizeof buf, "%.15Lg", f);
    return strtod(buf, NULL);
}

static double legendre(int l, int m, double xv)
{
    double res = 0.0;
    if (l >= m) {
        double p = 1.0, pm1 = 0.0, pm2 = 0.0;
        if (m != 0) {
            const double sq = sqrt(1.0 - xv * xv);
            for (int i = 1; i < m + 1; ++i) p *= (2 * i - 1) * sq;
        }
        for (int i = m + 1; i < l + 1; ++i) {
            pm2 = pm1;
            pm1 = p;
            p = ((2 * i - 1) * xv * pm1 - (i + m - 1) * pm2) / (i - m);
        }
        res = p;
    }
    return res;
}

static double polar_pref(int l, int m, double ct)
{
    const double PI = 3.14159265358979323846;
    const int ma = abs(m);
    double pf = 1.0;
    for (int i = l - ma + 1; i < l + ma + 1; ++i) pf *= i;
    pf = sqrt((2 * l + 1) / (4 * PI * pf)) * legendre(l, ma, ct);
    if ((m < 0) & (m % 2)) pf = -pf;
    return pf;
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

void port_get_sq(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
                 const int *boundary3, const int *verlet, int M, const double *dist, const int *nn,
                 const double *weight, const int *llist, int ndeg, int nnn, int lmax, int wl, int wlhat, int average,
                 int use_voronoi, double rc, int use_weight, double *qr, double *qi, double *qn, int ncol, int num_t)
{
    const double EPS = 1e-15, PI = 3.14159265358979323846;
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const int nz = 2 * lmax + 1, stride = ndeg * nz;
#pragma omp parallel for num_threads(num_t) schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        double wsum = 0.0;
        int cnt = nn[i];
        if (!use_voronoi && nnn > 0) cnt = nnn;
        double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
        for (int jj = 0; jj < cnt; ++jj) {
            const size_t at = (size_t)i * M + jj;
            const int j = verlet[at];
            if (j < 0) continue;
            double dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
            min_image(&c, &dx, &dy, &dz);
            const double r = dist[at];
            if (!(r > EPS && r <= rc)) continue;
            const double w = use_weight ? weight[at] : 1.0;
            wsum += w;
            const double rinv = 1.0 / r, ct = dz * rinv;
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
            for (int il = 0; il < ndeg; ++il) {
                const int l = llist[il];
                double *R = Qr + il * nz, *I = Qi + il * nz;
                R[l] += w * polar_pref(l, 0, ct);
                double pr = er, pi = ei;
                for (int m = 1; m < l + 1; ++m) {
                    const double pf = polar_pref(l, m, ct);
                    const double cr = pf * pr, ci = pf * pi;
                    const double wr = w * cr, wi = w * ci;
                    R[l + m] += wr;
                    I[l + m] += wi;
                    if (m & 1) {
                        R[l - m] -= wr;
                        I[l - m] += wi;
                    } else {
                        R[l - m] += wr;
                        I[l - m] -= wi;
                    }
                    const double tr = pr * er - pi * ei, ti = pr * ei + pi * er;
                    pr = tr;
                    pi = ti;
                }
            }
        }
        const double fac = 1.0 / wsum;
        for (int il = 0; il < ndeg; ++il)
            for (int m = 0; m < 2 * llist[il] + 1; ++m) {
                Qr[il * nz + m] *= fac;
                Qi[il * nz + m] *= fac;
            }
    }
    if (average) {
        const size_t tot = (size_t)N * stride;
        double *ar = (double *)malloc(sizeof(double) * tot), *ai = (double *)malloc(sizeof(double) * tot);
        memcpy(ar, qr, sizeof(double) * tot);
        memcpy(ai, qi, sizeof(double) * tot);
#pragma omp parallel for num_threads(num_t) schedule(dynamic, 16)
        for (int i = 0; i < N; ++i) {
            int cnt = nn[i];
            if (!use_voronoi && nnn > 0) cnt = nnn;
            int used = 1;
            double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
            for (int jj = 0; jj < cnt; ++jj) {
                const int j = verlet[(size_t)i * M + jj];
                if (j < 0) continue;
                for (int il = 0; il < ndeg; ++il)
                    for (int m = 0; m < 2 * llist[il] + 1; ++m) {
                        Qr[il * nz + m] += ar[(size_t)j * stride + il * nz + m];
                        Qi[il * nz + m] += ai[(size_t)j * stride + il * nz + m];
                    }
                ++used;
            }
            const double inv = 1.0 / used;
            for (int il = 0; il < ndeg; ++il)
                for (int m = 0; m < 2 * llist[il] + 1; ++m) {
                    Qr[il * nz + m] *= inv;
                    Qi[il * nz + m] *= inv;
                }
        }
        free(ar);
        free(ai);
    }
    /* Clebsch-Gordan table (cpp:188-224) */
    double *cg = NULL;
    if (wl || wlhat) {
        int ncg = 0;
        for (int il = 0; il < ndeg; ++il) {
            const int l = llist[il];
            for (int m1 = 0; m1 < 2 * l + 1; ++m1)
                for (int m2 = imax(0, l - m1); m2 < imin(2 * l + 1, 3 * l - m1 + 1); ++m2) ++ncg;
        }
        cg = (double *)malloc(sizeof(double) * (size_t)(ncg + 1));
        int t = 0;
        for (int il = 0; il < ndeg; ++il) {
            const int l = llist[il];
            for (int m1 = 0; m1 < 2 * l + 1; ++m1) {
                const int aa = m1 - l;
                for (int m2 = imax(0, l - m1); m2 < imin(2 * l + 1, 3 * l - m1 + 1); ++m2) {
                    const int bb = m2 - l, m = aa + bb + l;
                    double sums = 0.0;
                    for (int zz = imax(0, imax(-aa, bb)); zz < imin(l, imin(l - aa, l + bb)) + 1; ++zz) {
                        const int sgn = (zz % 2) ? -1 : 1;
                        sums += sgn / (fact(zz) * fact(l - zz) * fact(l - aa - zz) * fact(l + bb - zz) * fact(aa + zz) *
                                       fact(-bb + zz));
                    }
                    const int cc = m - l;
                    const double sf = sqrt(fact(l + aa) * fact(l - aa) * fact(l + bb) * fact(l - bb) * fact(l + cc) *
                                           fact(l - cc) * (2 * l + 1));
                    const double f1 = fact(3 * l + 1), f2 = fact(l);
                    cg[t++] = sums * sqrt(f2 * f2 * f2 / f1) * sf;
                }
            }
        }
    }
#pragma omp parallel for num_threads(num_t) schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        const double *Qr = qr + (size_t)i * stride, *Qi = qi + (size_t)i * stride;
        double *out = qn + (size_t)i * ncol;
        for (int il = 0; il < ndeg; ++il) {
            const int l = llist[il];
            double s = 0.0;
            for (int m = 0; m < 2 * l + 1; ++m) s += Qr[il * nz + m] * Qr[il * nz + m] + Qi[il * nz + m] * Qi[il * nz + m];
            out[il] = sqrt(4 * PI / (2 * l + 1)) * sqrt(s);
        }
        if (wl || wlhat) {
            int t = 0;
            for (int il = 0; il < ndeg; ++il) {
                const int l = llist[il];
                const double *R = Qr + il * nz, *I = Qi + il * nz;
                double ws = 0.0;
                for (int m1 = 0; m1 < 2 * l + 1; ++m1)
                    for (int m2 = imax(0, l - m1); m2 < imin(2 * l + 1, 3 * l - m1 + 1); ++m2, ++t) {
                        const int m = m1 + m2 - l;
                        const double ar = R[m1] * R[m2] - I[m1] * I[m2];
                        const double ai = R[m1] * I[m2] + I[m1] * R[m2];
                        ws += (ar * R[m] + ai * I[m]) * cg[t];
                    }
                const double wf = ws / sqrt(2 * l + 1.0);
                if (wl) out[il + ndeg] = wf;
                if (wlhat && out[il] > EPS) {
                    const double q = sqrt(4 * PI / (2 * l + 1)) / out[il];
                    out[il + (wl ? 1 : 0) * ndeg + ndeg] = wf * (q * q * q);
                }
            }
        }
    }
    free(cg);
}

void port_solid_liquid(int q6index, const double *Q6, const int *verlet, int N, int M, const double *dist,
                       const int *nn, const double *qlm_r, const double *qlm_i, int ndeg, int nz, double threshold,
                       int n_bond, int *solidliquid, int *nbond, int use_voronoi, int nnn, double rc, int num_t)
{
    const double PI = 3.14159265358979323846;
    const int stride = ndeg * nz;
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        int cnt = nn[i], solid = 0;
        if (!use_voronoi && nnn > 0) cnt = nnn;
        for (int jj = 0; jj < cnt; ++jj) {
            const int j = verlet[(size_t)i * M + jj];
            if (j < 0) continue;
            if (dist[(size_t)i * M + jj] > rc) continue;
            const double *ar = qlm_r + (size_t)i * stride + q6index * nz, *ai = qlm_i + (size_t)i * stride + q6index * nz;
            const double *br = qlm_r + (size_t)j * stride + q6index * nz, *bi = qlm_i + (size_t)j * stride + q6index * nz;
            double s = 0.0;
            for (int m = 0; m < 13; ++m) s += ar[m] * br[m] + ai[m] * bi[m];
            s = s / Q6[i] / Q6[j] * 4 * PI / 13;
            if (s > threshold) ++solid;
        }
        if (solid >= n_bond) solidliquid[i] = 1;
        nbond[i] = solid;
    }
    /* second sweep reads the first sweep's flags of other atoms while clearing its own
     * (cpp:645-674): do it from a snapshot so the result does not depend on thread timing */
    int *snap = (int *)malloc(sizeof(int) * (size_t)N);
    memcpy(snap, solidliquid, sizeof(int) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        if (snap[i] != 1) continue;
        int cnt = nn[i], any = 0;
        if (!use_voronoi && nnn > 0) cnt = nnn;
        for (int jj = 0; jj < cnt; ++jj) {
            const int j = verlet[(size_t)i * M + jj];
            if (j < 0) continue;
            if (solidliquid[j] == 1) {
                any = 1;
                break;
            }
        }
        if (!any) solidliquid[i] = 0;
    }
    free(snap);
}

/* ------------------------------------------------------------------ RDF
 * radial_distribution_function.cpp:22-85 list kernels, 143-317 streaming kernel */
void port_rdf(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list, double *g,
              int ntype, double rc, int nbin)
{
    const double dr = rc / nbin;
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < nn[i]; ++q) {
            const double d = dist[(size_t)i * M + q];
            if (d < rc) {
                const int j = verlet[(size_t)i * M + q];
                g[((size_t)type_list[i] * ntype + type_list[j]) * nbin + (int)(d / dr)] += 1.;
            }
        }
}

void port_rdf_single(const int *verlet, int N, int M, const double *dist, const int *nn, double *g, double rc,
                     int nbin)
{
    const double dr = rc / nbin;
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < nn[i]; ++q) {
            const double d = dist[(size_t)i * M + q];
            if (verlet[(size_t)i * M + q] > i && d < rc) g[(int)(d / dr)] += 2.0;
        }
}

/* Pair membership does not depend on the cell decomposition (the 27-cell window of
 * cpp:221-262 covers every pair within rc once when >= 3 cells per periodic axis, and the
 * fallback 266-305 is all pairs), so the restatement is the all-pairs form. */
void port_rdf_streaming(const double *x, const double *y, const double *z, int N, const int *type_list,
                        const double *box9, const double *origin3, const int *boundary3, double *g, int ntype,
                        double rc, int nbin, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const double dr = rc / nbin, rcsq = rc * rc;
    const size_t hs = (size_t)ntype * ntype * nbin;
#pragma omp parallel num_threads(num_t)
    {
        double *loc = (double *)calloc(hs, sizeof(double));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < N; ++i) {
            double xi = x[i], yi = y[i], zi = z[i];
            if (c.anyp) wrap_point(&c, &xi, &yi, &zi);
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
                min_image(&c, &dx, &dy, &dz);
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < rcsq) {
                    const int k = (int)(sqrt(r2) / dr);
                    if (k < nbin) loc[((size_t)type_list[i] * ntype + type_list[j]) * nbin + k] += 1.0;
                }
            }
        }
#pragma omp critical
        for (size_t t = 0; t < hs; ++t) g[t] += loc[t];
        free(loc);
    }
}

/* ------------------------------------------------------------------ replication
 * repeat_cell.cpp:19-61 */
void port_repeat_cell(double *new_pos, const double *old_box, const double *old_pos, int n_old, int nx, int ny,
                      int nz, int num_t)
{
    (void)num_t;
    size_t cellno = 0;
    for (int ix = 0; ix < nx; ++ix)
        for (int iy = 0; iy < ny; ++iy)
            for (int iz = 0; iz < nz; ++iz, ++cellno) {
                double sh[3];
                for (int d = 0; d < 3; ++d) sh[d] = ix * old_box[d] + iy * old_box[3 + d] + iz * old_box[6 + d];
                for (int a = 0; a < n_old; ++a)
                    for (int d = 0; d < 3; ++d)
                        new_pos[(cellno * n_old + a) * 3 + d] = old_pos[3 * a + d] + sh[d];
            }
}

/* ------------------------------------------------------------------ common neighbour parameter
 * common_neighbor_parameter.cpp:10-136 */
void port_cnp(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
              const int *boundary3, const int *verlet, int M, const double *dist, const int *nn, double *cnp, double rc,
              int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
#pragma omp parallel for num_threads(num_t) schedule(dynamic, 64)
    for (int i = 0; i < N; ++i) {
        int cnt = 0;
        double acc = 0.0;
        const int ni = nn[i];
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        for (int m = 0; m < ni; ++m) {
            if (!(di[m] <= rc)) continue;
            const int j = vi[m];
            ++cnt;
            double rx = 0.0, ry = 0.0, rz = 0.0;
            const int nj = nn[j];
            const int *vj = verlet + (size_t)j * M;
            const double *dj = dist + (size_t)j * M;
            for (int s2 = 0; s2 < nj; ++s2) {
                const int k = vj[s2];
                for (int h = 0; h < ni; ++h) {
                    if (vi[h] != k) continue;
                    if (dj[s2] <= rc && di[h] <= rc) {
                        double ax = x[i] - x[k], ay = y[i] - y[k], az = z[i] - z[k];
                        double bx = x[j] - x[k], by = y[j] - y[k], bz = z[j] - z[k];
                        min_image(&c, &ax, &ay, &az);
                        min_image(&c, &bx, &by, &bz);
                        rx += ax + bx;
                        ry += ay + by;
                        rz += az + bz;
                    }
                    break;
                }
            }
            acc += rx * rx + ry * ry + rz * rz;
        }
        cnp[i] = cnt > 0 ? acc / cnt : 1000.0;
    }
}

/* ------------------------------------------------------------------ Warren-Cowley parameter
 * warren_cowley_parameter.cpp:9-80 */
void port_wcp(const int *verlet, int N, int M, const int *nn, const int *type_list, int T, double *WCP, int num_t)
{
    (void)num_t;
    long long *Zmn = (long long *)calloc((size_t)T * T, sizeof(long long));
    long long *Zm = (long long *)calloc((size_t)T, sizeof(long long));
    double *alpha = (double *)calloc((size_t)T, sizeof(double));
    for (int i = 0; i < N; ++i) {
        const int it = type_list[i];
        alpha[it] += 1.0;
        Zm[it] += nn[i];
        for (int q = 0; q < nn[i]; ++q) Zmn[it * T + type_list[verlet[(size_t)i * M + q]]]++;
    }
    for (int i = 0; i < T; ++i) alpha[i] /= N;
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < T; ++j) {
            const int zmn = (int)Zmn[i * T + j], zm = (int)Zm[i];
            WCP[i * T + j] = (alpha[j] > 0 && zm > 0) ? 1.0 - (double)zmn / (alpha[j] * zm) : 0.0;
        }
    free(Zmn);
    free(Zm);
    free(alpha);
}

/* ------------------------------------------------------------------ average_by_neighbor
 * neighbor.cpp:704-743 */
void port_average_by_neighbor(double rc, const int *verlet, int N, int M, const double *dist, const int *nn,
                              const double *value, double *value_ave, int include_self, int num_t)
{
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        double sum = 0.0;
        int n = 0;
        if (include_self) {
            sum += value[i];
            ++n;
        }
        for (int q = 0; q < nn[i]; ++q)
            if (dist[(size_t)i * M + q] <= rc) {
                sum += value[verlet[(size_t)i * M + q]];
                ++n;
            }
        value_ave[i] = n > 0 ? sum / n : 0.0;
    }
}

/* ------------------------------------------------------------------ cluster analysis
 * cluster.cpp:9-60 (rc > 0: bond iff distance <= rc), 62-112 (dist == NULL: bond iff entry > -1), 114-150 */
int port_get_cluster(const int *verlet, int N, int M, const double *dist, const int *nn, double rc, int *clusters)
{
    int *queue = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int id = 0;
    for (int seed = 0; seed < N; ++seed) {
        if (clusters[seed] != -1) continue;
        int head = 0, tail = 0;
        queue[tail++] = seed;
        ++id;
        while (head < tail) {
            const int cur = queue[head++];
            int nl = 0;
            for (int q = 0; q < nn[cur]; ++q) {
                const int j = verlet[(size_t)cur * M + q];
                const int bond = dist ? (dist[(size_t)cur * M + q] <= rc) : (j > -1);
                if (!bond) continue;
                ++nl;
                if (clusters[j] == -1) {
                    clusters[j] = id;
                    queue[tail++] = j;
                }
            }
            if (nl == 0) clusters[cur] = id;
        }
    }
    free(queue);
    return id;
}

void port_filter_by_type(int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                         const int *t1, const int *t2, const double *r, int npair, int num_t)
{
    (void)num_t;
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < nn[i]; ++q) {
            const int j = verlet[(size_t)i * M + q];
            for (int k = 0; k < npair; ++k)
                if ((t1[k] == type_list[i]) & (t2[k] == type_list[j]) & (dist[(size_t)i * M + q] > r[k]))
                    verlet[(size_t)i * M + q] = -1;
        }
}

/* ------------------------------------------------------------------ structure entropy
 * structure_entropy.cpp:11-103 */
void port_structure_entropy(double rc, double sigma, int use_local_density, double volume, const double *dist, int N,
                            int M, const int *nn, double *entropy, int num_t)
{
    const double MY_PI = 3.14159265358979323846;
    const int nbins = (int)floor(rc / sigma) + 1;
    const double global_density = N / volume;
    const double step = rc / (nbins - 1);
    const double factor = (4. * MY_PI * global_density * sqrt(2. * MY_PI * sigma * sigma));
    const double sigma_sq = sigma * sigma;
    const double local_vol = 4. / 3. * MY_PI * rc * rc * rc;
    double *rl = (double *)malloc(sizeof(double) * (size_t)nbins);
    double *rsq = (double *)malloc(sizeof(double) * (size_t)nbins);
    double *pref = (double *)malloc(sizeof(double) * (size_t)nbins);
    for (int j = 0; j < nbins; ++j) {
        rl[j] = j * step;
        rsq[j] = rl[j] * rl[j];
        pref[j] = rsq[j] * factor;
    }
    pref[0] = pref[1];
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        const double *di = dist + (size_t)i * M;
        int n_neigh = 0;
        for (int k = 0; k < nn[i]; ++k) n_neigh += di[k] <= rc;
        double density = global_density, fac = 1.0;
        if (use_local_density) {
            density = n_neigh / local_vol;
            fac = global_density / density;
        }
        double sum = 0.0, prev = 0.0;
        for (int j = 0; j < nbins; ++j) {
            double g = 0.0;
            for (int k = 0; k < nn[i]; ++k)
                if (di[k] <= rc) {
                    const double delta = rl[j] - di[k];
                    g += exp(-(delta * delta) / (2.0 * sigma_sq)) / pref[j];
                }
            if (use_local_density) g *= fac;
            const double integrand = g >= 1e-10 ? (g * log(g) - g + 1.0) * rsq[j] : rsq[j];
            if (j > 0) sum += prev + integrand;
            prev = integrand;
        }
        entropy[i] = -MY_PI * density * sum * sigma;
    }
    free(rl);
    free(rsq);
    free(pref);
}

/* ------------------------------------------------------------------ atomic temperature
 * atomic_temperature.cpp:7-112 */
void port_compute_temp(const int *verlet, int N, int M, const double *dist, const double *vx, const double *vy,
                       const double *vz, const double *mass, double *T, double rc, int num_t)
{
    const double kb = 1.380649e-23, dim = 3.0, afu = 6.022140857e23;
    const double mass_factor = 1.0 / afu / 1000.0, vel_conv = 1e4;
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) {
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        const double mi = mass[i];
        double sx = vx[i] * mi, sy = vy[i] * mi, sz = vz[i] * mi, mn = mi;
        int n = 1;
        for (int q = 0; q < M; ++q) {
            const int j = vi[q];
            if (j < 0) break;
            if (j != i && di[q] <= rc) {
                sx += vx[j] * mass[j];
                sy += vy[j] * mass[j];
                sz += vz[j] * mass[j];
                ++n;
                mn += mass[j];
            }
        }
        const double mx = sx / mn, my = sy / mn, mz = sz / mn;
        double dx = vx[i] - mx, dy = vy[i] - my, dz = vz[i] - mz;
        double vsq = dx * dx + dy * dy + dz * dz;
        double ke = 0.0;
        ke += 0.5 * mi * mass_factor * vsq * vel_conv;
        for (int q = 0; q < M; ++q) {
            const int j = vi[q];
            if (j < 0) break;
            if (j != i && di[q] <= rc) {
                dx = vx[j] - mx;
                dy = vy[j] - my;
                dz = vz[j] - mz;
                vsq = dx * dx + dy * dy + dz * dz;
                ke += 0.5 * mass[j] * mass_factor * vsq * vel_conv;
            }
        }
        T[i] = ke * 2.0 / (dim * n * kb);
    }
}

/* ------------------------------------------------------------------ bond analysis / angular distribution
 * bond_analysis.cpp:7-118 and 120-240 */
static int angle_bin(const cell_t *c, const double *x, const double *y, const double *z, int i, int j, int k, double rij,
                     double rik, double dti, int nbins, int clamp_low)
{
    const double PI = 3.14159265358979323846;
    double ax = x[j] - x[i], ay = y[j] - y[i], az = z[j] - z[i];
    double bx = x[k] - x[i], by = y[k] - y[i], bz = z[k] - z[i];
    min_image(c, &ax, &ay, &az);
    min_image(c, &bx, &by, &bz);
    const double dot = ax * bx + ay * by + az * bz;
    double ct = dot / (rij * rik);
    if (ct > 1.0) ct = 1.0;
    if (ct < -1.0) ct = -1.0;
    const double theta = acos(ct) * 180.0 / PI;
    int index = (int)floor(theta * dti);
    if (clamp_low && index < 0) index = 0;
    if (index > nbins - 1) index = nbins - 1;
    return index;
}

void port_compute_bond(const double *x, const double *y, const double *z, int N, const double *box9,
                       const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                       const int *nn, int *blen, int *bang, double delta_r, double delta_theta, double rc, int nbins,
                       int num_t)
{
    (void)num_t;
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const double dri = 1.0 / delta_r, dti = 1.0 / delta_theta;
    for (int i = 0; i < N; ++i) {
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        for (int jj = 0; jj < nn[i]; ++jj)
            if (vi[jj] > i && di[jj] <= rc) {
                int index = (int)floor(di[jj] * dri);
                if (index > nbins - 1) index = nbins - 1;
                blen[index] += 1;
            }
        for (int jj = 0; jj < nn[i]; ++jj) {
            if (!(di[jj] <= rc)) continue;
            for (int kk = jj + 1; kk < nn[i]; ++kk)
                if (di[kk] <= rc) bang[angle_bin(&c, x, y, z, i, vi[jj], vi[kk], di[jj], di[kk], dti, nbins, 0)] += 1;
        }
    }
}

void port_compute_adf(const double *x, const double *y, const double *z, int N, const double *box9,
                      const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                      const int *nn, double delta_theta, const double *rcs, const int *pairs, int npair,
                      const int *types, int nbins, int *bang, int num_t)
{
    (void)num_t;
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
    const double dti = 1.0 / delta_theta;
    for (int i = 0; i < N; ++i) {
        const int *vi = verlet + (size_t)i * M;
        const double *di = dist + (size_t)i * M;
        for (int m = 0; m < npair; ++m) {
            if (types[i] != pairs[m * 3]) continue;
            const int jt = pairs[m * 3 + 1], kt = pairs[m * 3 + 2], same = jt == kt;
            for (int jj = 0; jj < nn[i]; ++jj) {
                if (types[vi[jj]] != jt) continue;
                if (!(di[jj] <= rcs[m * 4 + 1] && di[jj] >= rcs[m * 4 + 0])) continue;
                for (int kk = same ? jj + 1 : 0; kk < nn[i]; ++kk) {
                    if (kk == jj || types[vi[kk]] != kt) continue;
                    if (!(di[kk] <= rcs[m * 4 + 3] && di[kk] >= rcs[m * 4 + 2])) continue;
                    bang[(size_t)m * nbins + angle_bin(&c, x, y, z, i, vi[jj], vi[kk], di[jj], di[kk], dti, nbins, 1)] += 1;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ wrap_positions, neighbor.cpp:675-702 */
void port_wrap_positions(double *x, double *y, double *z, int N, const double *box9, const double *origin3,
                         const int *boundary3, int num_t)
{
    cell_t c;
    cell_init(&c, box9, origin3, boundary3);
#pragma omp parallel for num_threads(num_t)
    for (int i = 0; i < N; ++i) wrap_point(&c, x + i, y + i, z + i);
}
