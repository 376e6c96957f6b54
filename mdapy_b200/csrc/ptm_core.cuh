// mdapy_b200/csrc/ptm_core.cuh
//
// Polyhedral template matching (Larsen, Schmidt, Schiotz 2016), per-atom core, written from the
// published method and from the BEHAVIOUR of the reference's driver + vendored library
// (src/polyhedral_template_matching.cpp:135-319, extern/ptm/ptm_index.cpp:114-193,
// ptm_structure_matcher.cpp:29-186, ptm_neighbour_ordering.cpp:48-201).  Everything here is
// __host__ __device__ so the same arithmetic can be exercised on the CPU by the tests; the product
// only ever runs it inside the CUDA kernel of ptm.cu.
//
// Pipeline for one atom (SC / FCC / HCP / ICO / BCC; the two-shell structures DCUB / DHEX / graphene go through
// the same steps on the neighbours of the four / three first neighbours, see match_two_shell below):
//   1. neighbour vectors r_k = min_image(x_k - x_i) of the <= 18 listed nearest neighbours;
//   2. pre-ordering: Voronoi cell of the atom against those points, faces ranked by solid angle
//      (descending), ties by distance, then input order -- ptm_neighbour_ordering.cpp:28-201.
//      The cell is obtained face by face: the polygon of plane k is clipped by every other bisector
//      plane and by the +-10 r_max cube, its solid angle summed over a fan of spherical triangles
//      (Van Oosterom & Strackee, as ptm_solid_angles.cpp:36-48);
//   3. per structure: first n ordered points -> incremental convex hull -> facet count / vertex degree
//      checks -> canonical code of the oriented triangulation -> look-up among the structure's template
//      triangulations -> for every automorphism an optimal-rotation RMSD (quaternion characteristic
//      polynomial, Horn / Theobald) -> smallest RMSD wins (ptm_structure_matcher.cpp:57-105);
//   4. scale -> interatomic distance, orientation rotated into the fundamental zone of the structure's
//      rotation group (ptm_quat.cpp:180-207), alloy ordering (ptm_alloy_types.cpp).
//
// Template triangulations, their automorphisms and the rotation groups are GENERATED at start-up
// (ptm_tables.cu), not tabulated: every triangulation of the coplanar faces of the ideal polyhedron
// is enumerated and canonicalised with the same routine used for the atomic environments.
#pragma once
#include <cmath>
#include <cstdint>
#include "box.cuh"

// large cold-ish building blocks are kept out of line on the device: the fully inlined kernel was 460 KB
// of SASS and stalled on instruction fetch
#ifdef __CUDACC__
#define MDB_HDN __host__ __device__ __noinline__
#else
#define MDB_HDN inline
#endif

namespace ptm {

constexpr int MAX_IN = 18;      // neighbours offered per atom (PTM_MAX_INPUT_POINTS - 1)
constexpr int MAX_NB = 16;      // neighbours of the largest supported structure (diamond: 4 + 12)
constexpr int MAX_FACETS = 28;  // 2n - 4 for n = 16
constexpr int MAX_CODE = 84;    // 3 * facets = 2 * edges
constexpr int MAX_MULTISHELL = 13;  // two-shell environments only use the 13 nearest listed neighbours (ptm_multishell.h:20)
constexpr int MAX_POLY = 28;    // vertices of one Voronoi face during clipping (<= 17 bisectors + 4 + cube corners)

enum { S_SC = 0, S_FCC = 1, S_HCP = 2, S_ICO = 3, S_BCC = 4, S_DCUB = 5, S_DHEX = 6, S_GRAPHENE = 7, NSTRUCT = 8 };
// reference structure ids (ptm_constants.h): FCC 1, HCP 2, BCC 3, ICO 4, SC 5, DCUB 6, DHEX 7, graphene 8
constexpr int CHECK_FCC = 1, CHECK_HCP = 2, CHECK_BCC = 4, CHECK_ICO = 8, CHECK_SC = 16, CHECK_DCUB = 32,
              CHECK_DHEX = 64, CHECK_GRAPHENE = 128;

struct Tables {
    int n_nbrs[NSTRUCT], n_facets[NSTRUCT], max_degree[NSTRUCT], type_id[NSTRUCT], group[NSTRUCT];
    int n_inner[NSTRUCT];                 // two-shell structures: neighbours of the first shell (coloured 1), else 0
    double tpl[NSTRUCT][MAX_NB + 1][3];   // template points (0 = centre), barycentre 0, mean distance 1
    double c_dist[NSTRUCT];               // |template[1]|: interatomic distance = c_dist / scale
    int graph_begin[NSTRUCT + 1];         // graphs of structure s: [graph_begin[s], graph_begin[s+1]) sorted by hash
    const unsigned long long *hash;       // [n_graphs]
    const int *aut_begin;                 // [n_graphs + 1]
    const signed char *aut_label;         // [n_aut][MAX_NB]: template neighbour -> canonical label
    int gen_begin[4];                     // rotation groups: 0 cubic (24), 1 hexagonal conventional (12), 2 icosahedral (60)
    const double *gen;                    // [n_gen][4] unit quaternions
    unsigned fcc_plane[3];                // bit masks (template point indices 1..12) of the three {100} planes
};

// ---------------------------------------------------------------------------------------------
// small vector helpers
MDB_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
MDB_HD void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---------------------------------------------------------------------------------------------
// 2. Voronoi solid angles

// clip polygon (n vertices, in order) by the half-space nrm.x <= d; returns new vertex count
MDB_HD int clip_poly(const double (*in)[3], int n, const double *nrm, double d, double (*out)[3])
{
    int m = 0;
    if (n == 0) return 0;
    double sp = dot3(in[n - 1], nrm) - d;
    for (int i = 0; i < n; ++i) {
        const double *cur = in[i];
        const double *prv = in[i == 0 ? n - 1 : i - 1];
        const double sc = dot3(cur, nrm) - d;
        if ((sc <= 0) != (sp <= 0)) {  // edge crosses the plane
            const double t = sp / (sp - sc);
            if (m < MAX_POLY) {
                out[m][0] = prv[0] + t * (cur[0] - prv[0]);
                out[m][1] = prv[1] + t * (cur[1] - prv[1]);
                out[m][2] = prv[2] + t * (cur[2] - prv[2]);
                ++m;
            }
        }
        if (sc <= 0 && m < MAX_POLY) {
            out[m][0] = cur[0];
            out[m][1] = cur[1];
            out[m][2] = cur[2];
            ++m;
        }
        sp = sc;
    }
    return m;
}

// 0: the plane cuts nothing (polygon unchanged), 1: it cuts, 2: nothing is left
MDB_HD int clip_class(const double (*in)[3], int n, const double *nrm, double d)
{
    bool out = false, in_ = false;
    for (int i = 0; i < n; ++i) {
        const double sc = dot3(in[i], nrm) - d;
        if (sc <= 0) in_ = true;
        else out = true;
    }
    return !out ? 0 : (in_ ? 1 : 2);
}

// solid angle subtended at the origin by the Voronoi face of point k (0 when the face is empty).
// The face is the bisector plane of k clipped by the other bisectors (nearest first: the polygon shrinks
// to its final size after a few planes and the remaining ones are rejected by clip_class without touching
// it) and finally by the bounding cube of the reference's cell (+-10 r_max), which only matters for atoms
// whose cell is open.  The solid angle of the fan is the argument of the product of the triangles'
// (den + i num) terms (Van Oosterom-Strackee per triangle): one atan2 per face.
MDB_HDN double voronoi_face_solid_angle(int num, const double (*pts)[3], const double *normsq, double box_half, int k)
{
    double bufa[MAX_POLY][3], bufb[MAX_POLY][3];
    // a large square on the bisector plane of point k
    const double *p = pts[k];
    const double pn = sqrt(normsq[k]);
    double u[3], v[3], e[3] = {0, 0, 0};
    // axis least aligned with p
    int ax = 0;
    if (fabs(p[1]) < fabs(p[ax])) ax = 1;
    if (fabs(p[2]) < fabs(p[ax])) ax = 2;
    e[ax] = 1;
    cross3(p, e, u);
    double un = sqrt(dot3(u, u));
    u[0] /= un;
    u[1] /= un;
    u[2] /= un;
    cross3(p, u, v);
    v[0] /= pn;
    v[1] /= pn;
    v[2] /= pn;
    const double S = 4 * box_half;
    const double c0[3] = {0.5 * p[0], 0.5 * p[1], 0.5 * p[2]};
    const double sg[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
    for (int i = 0; i < 4; ++i)
        for (int d = 0; d < 3; ++d) bufa[i][d] = c0[d] + S * (sg[i][0] * u[d] + sg[i][1] * v[d]);
    int n = 4;
    double(*src)[3] = bufa;
    double(*dst)[3] = bufb;
    for (int j = 0; j < num + 6 && n; ++j) {
        if (j == k) continue;
        double cube[3] = {0, 0, 0};
        const double *nrm;
        double d;
        if (j < num) {
            nrm = pts[j];
            d = 0.5 * normsq[j];
        } else {
            cube[(j - num) >> 1] = ((j - num) & 1) ? -1.0 : 1.0;
            nrm = cube;
            d = box_half;
        }
        const int cls = clip_class(src, n, nrm, d);
        if (cls == 0) continue;
        if (cls == 2) {
            n = 0;
            break;
        }
        n = clip_poly(src, n, nrm, d, dst);
        double(*t)[3] = src;
        src = dst;
        dst = t;
    }
    if (n < 3) return 0.0;
    // unit vectors, fan from the first vertex
    for (int i = 0; i < n; ++i) {
        const double nr = sqrt(dot3(src[i], src[i]));
        src[i][0] /= nr;
        src[i][1] /= nr;
        src[i][2] /= nr;
    }
    double re = 1.0, im = 0.0;
    for (int i = 2; i < n; ++i) {
        double c[3];
        cross3(src[i - 1], src[i], c);
        const double tn = dot3(src[0], c);
        const double td = 1 + dot3(src[0], src[i - 1]) + dot3(src[i], src[0]) + dot3(src[i - 1], src[i]);
        const double r2 = re * td - im * tn;
        im = re * tn + im * td;
        re = r2;
    }
    return fabs(2 * atan2(im, re));
}

// The same face in the 2-D coordinates (a, b) of its own plane, x = p/2 + a u + b v: a clipping plane
// becomes the line s0 + a su + b sv <= 0 (three dot products per plane, two multiply-adds per vertex), and
// the two small vertex buffers fit in shared memory on the device (element i of a buffer lives at
// buf[i * STRIDE]; STRIDE = 1 on the host, = block size on the device with buf offset by the thread).
// Returns -1 when a polygon outgrows the buffers: the caller then uses the 3-D routine above.
constexpr int MAX_POLY2 = 16;

template <int STRIDE>
MDB_HD double voronoi_face_solid_angle_2d(int num, const double (*pts)[3], const double *normsq, double box_half, int k,
                                          double *buf)
{
    const double *p = pts[k];
    const double pn = sqrt(normsq[k]);
    double u[3], v[3], e[3] = {0, 0, 0};
    int ax = 0;
    if (fabs(p[1]) < fabs(p[ax])) ax = 1;
    if (fabs(p[2]) < fabs(p[ax])) ax = 2;
    e[ax] = 1;
    cross3(p, e, u);
    const double un = sqrt(dot3(u, u));
    u[0] /= un;
    u[1] /= un;
    u[2] /= un;
    cross3(p, u, v);
    v[0] /= pn;
    v[1] /= pn;
    v[2] /= pn;
    const double c0[3] = {0.5 * p[0], 0.5 * p[1], 0.5 * p[2]};
    const double S = 4 * box_half;
    // buffer layout: [which (2)][coordinate (2)][vertex (MAX_POLY2)]
    double *A0 = buf, *B0 = buf + MAX_POLY2 * STRIDE, *A1 = buf + 2 * MAX_POLY2 * STRIDE, *B1 = buf + 3 * MAX_POLY2 * STRIDE;
    A0[0] = S;
    B0[0] = S;
    A0[STRIDE] = -S;
    B0[STRIDE] = S;
    A0[2 * STRIDE] = -S;
    B0[2 * STRIDE] = -S;
    A0[3 * STRIDE] = S;
    B0[3 * STRIDE] = -S;
    int n = 4;
    double *sa = A0, *sb = B0, *da = A1, *db = B1;
    for (int j = 0; j < num + 6 && n; ++j) {
        if (j == k) continue;
        double s0, su, sv;
        if (j < num) {
            const double *q = pts[j];
            s0 = dot3(c0, q) - 0.5 * normsq[j];
            su = dot3(u, q);
            sv = dot3(v, q);
        } else {
            const int d = (j - num) >> 1;
            const double sg = ((j - num) & 1) ? -1.0 : 1.0;
            s0 = sg * c0[d] - box_half;
            su = sg * u[d];
            sv = sg * v[d];
        }
        bool any_out = false, any_in = false;
        for (int i = 0; i < n; ++i) {
            const double sc = s0 + sa[i * STRIDE] * su + sb[i * STRIDE] * sv;
            if (sc <= 0) any_in = true;
            else any_out = true;
        }
        if (!any_out) continue;
        if (!any_in) {
            n = 0;
            break;
        }
        double pa = sa[(n - 1) * STRIDE], pb = sb[(n - 1) * STRIDE];
        double sp = s0 + pa * su + pb * sv;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const double ca = sa[i * STRIDE], cb = sb[i * STRIDE];
            const double sc = s0 + ca * su + cb * sv;
            if ((sc <= 0) != (sp <= 0)) {
                if (m >= MAX_POLY2) return -1.0;
                const double t = sp / (sp - sc);
                da[m * STRIDE] = pa + t * (ca - pa);
                db[m * STRIDE] = pb + t * (cb - pb);
                ++m;
            }
            if (sc <= 0) {
                if (m >= MAX_POLY2) return -1.0;
                da[m * STRIDE] = ca;
                db[m * STRIDE] = cb;
                ++m;
            }
            pa = ca;
            pb = cb;
            sp = sc;
        }
        n = m;
        double *t1 = sa;
        sa = da;
        da = t1;
        t1 = sb;
        sb = db;
        db = t1;
    }
    if (n < 3) return 0.0;
    // fan of spherical triangles from vertex 0; product of the (den + i num) terms, one atan2
    double f0[3], fp[3], fc[3];
    auto unit = [&](int i, double *o) {
        const double a = sa[i * STRIDE], b = sb[i * STRIDE];
        o[0] = c0[0] + a * u[0] + b * v[0];
        o[1] = c0[1] + a * u[1] + b * v[1];
        o[2] = c0[2] + a * u[2] + b * v[2];
        const double nr = sqrt(dot3(o, o));
        o[0] /= nr;
        o[1] /= nr;
        o[2] /= nr;
    };
    unit(0, f0);
    unit(1, fp);
    double re = 1.0, im = 0.0;
    for (int i = 2; i < n; ++i) {
        unit(i, fc);
        double c[3];
        cross3(fp, fc, c);
        const double tn = dot3(f0, c);
        const double td = 1 + dot3(f0, fp) + dot3(fc, f0) + dot3(fp, fc);
        const double r2 = re * td - im * tn;
        im = re * tn + im * td;
        re = r2;
        fp[0] = fc[0];
        fp[1] = fc[1];
        fp[2] = fc[2];
    }
    return fabs(2 * atan2(im, re));
}

// order[0..num) = input indices ranked by (solid angle desc, distance asc, input order).
// buf: 4 * MAX_POLY2 doubles of scratch per caller, strided by STRIDE (see above).
template <int STRIDE>
MDB_HD void preorder_neighbours(int num, const double (*pts)[3], int *order, double *buf)
{
    double normsq[MAX_IN], area[MAX_IN];
    double mx = 0;
    for (int i = 0; i < num; ++i) {
        normsq[i] = dot3(pts[i], pts[i]);
        mx = mx > normsq[i] ? mx : normsq[i];
    }
    const double box_half = 10 * sqrt(mx);
    for (int i = 0; i < num; ++i) {
        double a = voronoi_face_solid_angle_2d<STRIDE>(num, pts, normsq, box_half, i, buf);
        if (a < 0) a = voronoi_face_solid_angle(num, pts, normsq, box_half, i);
        area[i] = a;
    }
    // stable insertion sort
    for (int i = 0; i < num; ++i) order[i] = i;
    for (int i = 1; i < num; ++i) {
        const int o = order[i];
        int j = i - 1;
        while (j >= 0) {
            const int q = order[j];
            const bool before = area[o] > area[q] || (area[o] == area[q] && normsq[o] < normsq[q]);
            if (!before) break;
            order[j + 1] = q;
            --j;
        }
        order[j + 1] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// 3a. incremental convex hull of points[0..np) (0 = central atom).  Facets are returned over the
// neighbour indices 0..np-2, counter-clockwise seen from outside.  Returns the facet count, or a
// negative value when the point set is degenerate, the hull overflows or the centre lies on it.
MDB_HDN int convex_hull(int np, const double (*P)[3], signed char (*facets)[3])
{
    const double TOL = 1e-12;
    signed char F[2 * MAX_NB + 8][3];
    double Nrm[2 * MAX_NB + 8][3];
    int nf = 0;
    // initial simplex: extreme pair along the longest axis, farthest from the line, farthest from the plane
    int a = 0, b = 0;
    {
        double best = -1;
        for (int d = 0; d < 3; ++d) {
            int lo = 0, hi = 0;
            for (int i = 1; i < np; ++i) {
                if (P[i][d] < P[lo][d]) lo = i;
                if (P[i][d] > P[hi][d]) hi = i;
            }
            const double dd[3] = {P[hi][0] - P[lo][0], P[hi][1] - P[lo][1], P[hi][2] - P[lo][2]};
            const double len = dot3(dd, dd);
            if (lo != hi && len > best) {
                best = len;
                a = lo;
                b = hi;
            }
        }
        if (best <= 0) return -1;
    }
    int c = -1, d4 = -1;
    {
        const double ab[3] = {P[b][0] - P[a][0], P[b][1] - P[a][1], P[b][2] - P[a][2]};
        const double ab2 = dot3(ab, ab);
        double best = 0;
        for (int i = 0; i < np; ++i) {
            if (i == a || i == b) continue;
            const double w[3] = {P[a][0] - P[i][0], P[a][1] - P[i][1], P[a][2] - P[i][2]};
            const double dt = dot3(w, ab);
            const double dist = (dot3(w, w) * ab2 - dt * dt) / ab2;
            if (dist > best) {
                best = dist;
                c = i;
            }
        }
        if (!(best > TOL)) return -2;
        double n[3];
        const double ac[3] = {P[c][0] - P[a][0], P[c][1] - P[a][1], P[c][2] - P[a][2]};
        cross3(ab, ac, n);
        const double nn = sqrt(dot3(n, n));
        best = 0;
        for (int i = 0; i < np; ++i) {
            if (i == a || i == b || i == c) continue;
            const double w[3] = {P[i][0] - P[a][0], P[i][1] - P[a][1], P[i][2] - P[a][2]};
            const double dist = fabs(dot3(w, n)) / nn;
            if (dist > best) {
                best = dist;
                d4 = i;
            }
        }
        if (!(best > TOL)) return -3;
    }
    double ctr[3];
    for (int d = 0; d < 3; ++d) ctr[d] = 0.25 * (P[a][d] + P[b][d] + P[c][d] + P[d4][d]);
    bool done[MAX_NB + 1];
    for (int i = 0; i < np; ++i) done[i] = false;
    done[a] = done[b] = done[c] = done[d4] = true;

    // adds facet (u, v, w), oriented so that the interior point ctr is behind it
    auto add = [&](int u, int v, int w) -> bool {
        if (nf >= 2 * MAX_NB + 8) return false;
        const double e1[3] = {P[v][0] - P[u][0], P[v][1] - P[u][1], P[v][2] - P[u][2]};
        const double e2[3] = {P[w][0] - P[u][0], P[w][1] - P[u][1], P[w][2] - P[u][2]};
        double n[3];
        cross3(e1, e2, n);
        const double nn = sqrt(dot3(n, n));
        n[0] /= nn;
        n[1] /= nn;
        n[2] /= nn;
        const double dc[3] = {ctr[0] - P[u][0], ctr[1] - P[u][1], ctr[2] - P[u][2]};
        if (dot3(n, dc) > 0) {
            n[0] = -n[0];
            n[1] = -n[1];
            n[2] = -n[2];
            const int t = v;
            v = w;
            w = t;
        }
        F[nf][0] = (signed char)u;
        F[nf][1] = (signed char)v;
        F[nf][2] = (signed char)w;
        Nrm[nf][0] = n[0];
        Nrm[nf][1] = n[1];
        Nrm[nf][2] = n[2];
        ++nf;
        return true;
    };
    add(a, b, c);
    add(a, b, d4);
    add(a, c, d4);
    add(b, c, d4);

    for (int i = 0; i < np; ++i) {
        if (done[i]) continue;
        done[i] = true;
        // classify facets, remember the directed edges of the visible ones
        unsigned vis_edge[MAX_NB + 1];  // bit v of vis_edge[u]: directed edge u->v belongs to a visible facet
        for (int q = 0; q < np; ++q) vis_edge[q] = 0;
        bool any = false;
        bool visible[2 * MAX_NB + 8];
        for (int f = 0; f < nf; ++f) {
            const double w[3] = {P[i][0] - P[F[f][0]][0], P[i][1] - P[F[f][0]][1], P[i][2] - P[F[f][0]][2]};
            visible[f] = dot3(w, Nrm[f]) > TOL;
            if (visible[f]) {
                any = true;
                vis_edge[F[f][0]] |= 1u << F[f][1];
                vis_edge[F[f][1]] |= 1u << F[f][2];
                vis_edge[F[f][2]] |= 1u << F[f][0];
            }
        }
        if (!any) continue;  // inside (or on) the current hull: dropped, the facet count will not match
        // horizon: directed edges u->v of visible facets whose reverse v->u is not visible
        signed char hz[2 * MAX_NB + 8][2];
        int nh = 0;
        int keep = 0;
        for (int f = 0; f < nf; ++f) {
            if (visible[f]) {
                for (int e = 0; e < 3; ++e) {
                    const int u = F[f][e], v = F[f][(e + 1) % 3];
                    if (!(vis_edge[v] >> u & 1u)) {
                        if (nh >= 2 * MAX_NB + 8) return -4;
                        hz[nh][0] = (signed char)u;
                        hz[nh][1] = (signed char)v;
                        ++nh;
                    }
                }
            } else {
                if (keep != f) {
                    for (int e = 0; e < 3; ++e) {
                        F[keep][e] = F[f][e];
                        Nrm[keep][e] = Nrm[f][e];
                    }
                }
                ++keep;
            }
        }
        nf = keep;
        for (int h = 0; h < nh; ++h)
            if (!add(hz[h][0], hz[h][1], i)) return -4;
    }
    if (nf > MAX_FACETS) return -4;
    for (int f = 0; f < nf; ++f) {
        if (F[f][0] == 0 || F[f][1] == 0 || F[f][2] == 0) return -6;  // central atom on the hull
        facets[f][0] = (signed char)(F[f][0] - 1);
        facets[f][1] = (signed char)(F[f][1] - 1);
        facets[f][2] = (signed char)(F[f][2] - 1);
    }
    return nf;
}

// ---------------------------------------------------------------------------------------------
// 3b. canonical code of an oriented triangulation with n vertices.
// nxt[u][v] = neighbour following v in the counter-clockwise rotation around u.
// For a start dart (s -> t): label s = 0, t = 1, then vertices are processed in label order, each one
// listing its neighbours in rotation order starting from the neighbour it was discovered from;
// unlabelled neighbours receive the next labels.  The code is the concatenated list of labels; the
// canonical code is the lexicographic minimum over all start darts, and every dart attaining it is an
// orientation-preserving automorphism.
struct Rotation {
    signed char nxt[MAX_NB][MAX_NB];
    signed char deg[MAX_NB];
};

MDB_HD bool build_rotation(int n, int nf, const signed char (*facets)[3], Rotation &R)
{
    for (int u = 0; u < n; ++u) {
        R.deg[u] = 0;
        for (int v = 0; v < n; ++v) R.nxt[u][v] = -1;
    }
    for (int f = 0; f < nf; ++f)
        for (int e = 0; e < 3; ++e) {
            const int u = facets[f][e], v = facets[f][(e + 1) % 3], w = facets[f][(e + 2) % 3];
            if (R.nxt[u][v] >= 0) return false;  // not a closed oriented surface
            R.nxt[u][v] = (signed char)w;
            ++R.deg[u];
        }
    return true;
}

// Code for one start dart, compared on the fly with the best code so far (best_len < 0: none yet).
// Returns 1 when the new code is lexicographically smaller (code[0..len) is then complete), 0 when it
// equals the best, -1 as soon as it is known to be larger (the walk stops there), -2 when the walk breaks
// (malformed surface).  All codes of one graph have the same length (one entry per dart).
// n_inner > 0: vertices 0..n_inner-1 carry colour 1 and enter the code as n + label
// (ptm_canonical_coloured.cpp:33-44), so only colour-preserving relabellings compare equal.
MDB_HD int dart_code(int n, const Rotation &R, int s, int t, signed char *label, signed char *code, int &len,
                     const signed char *best, int best_len, int n_inner = 0)
{
    signed char ref[MAX_NB], byl[MAX_NB];
    for (int u = 0; u < n; ++u) label[u] = -1;
    label[s] = 0;
    label[t] = 1;
    byl[0] = (signed char)s;
    byl[1] = (signed char)t;
    ref[s] = (signed char)t;
    ref[t] = (signed char)s;
    int count = 2;
    int state = best_len < 0 ? 1 : 0;
    int total = 0;
    for (int u = 0; u < n; ++u) total += R.deg[u];
    if (total > MAX_CODE) return -2;
    // one flat loop over the darts (same trip count for every atom of a shell: SIMT friendly);
    // (k, v, j, w) = position in the label-ordered vertex list, that vertex, neighbours listed so far, next neighbour
    int k = 0, v = s, j = 0, w = t, dv = R.deg[s];
    for (len = 0; len < total; ++len) {
        if (w < 0) return -2;
        if (label[w] < 0) {
            label[w] = (signed char)count;
            byl[count] = (signed char)w;
            ref[w] = (signed char)v;
            ++count;
        }
        const signed char c = (signed char)(label[w] + (w < n_inner ? n : 0));
        if (state == 0) {
            if (c > best[len]) return -1;
            if (c < best[len]) state = 1;
        }
        code[len] = c;
        w = R.nxt[v][w];
        if (++j == dv) {
            if (++k >= count) {
                if (len + 1 < total) return -2;  // disconnected
                continue;
            }
            v = byl[k];
            w = ref[v];
            j = 0;
            dv = R.deg[v];
        }
    }
    return count == n ? state : -2;
}

MDB_HD unsigned long long code_hash(const signed char *code, int len)
{
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) {
        h ^= (unsigned long long)(unsigned char)code[i];
        h *= 1099511628211ull;
    }
    return h ^ ((unsigned long long)len << 56);
}

// Start darts are restricted to those whose (deg s, deg t, deg third-vertex-of-the-facet) triple is the
// largest in the graph -- an isomorphism-invariant choice (the reference prunes the same way,
// ptm_canonical_coloured.cpp:120-167), which leaves a handful of darts instead of 3 * facets.
MDB_HD int dart_key(const Rotation &R, int s, int t, int w) { return (R.deg[s] << 16) | (R.deg[t] << 8) | R.deg[w]; }

// canonical code + ONE canonical labelling of an environment graph (the first start dart, in facet order,
// that attains the minimum)
MDB_HDN bool canonical_form(int n, int nf, const signed char (*facets)[3], const Rotation &R, signed char *best_label,
                           unsigned long long &hash, int n_inner = 0)
{
    signed char bufa[MAX_CODE], bufb[MAX_CODE], laba[MAX_NB], labb[MAX_NB];
    signed char *best = bufa, *cur = bufb, *blab = laba, *clab = labb;
    int best_len = -1;
    int top = 0;
    for (int f = 0; f < nf; ++f)
        for (int e = 0; e < 3; ++e) {
            const int k = dart_key(R, facets[f][e], facets[f][(e + 1) % 3], facets[f][(e + 2) % 3]);
            top = top > k ? top : k;
        }
    // the qualifying start darts are listed first, so that the threads of a warp walk their lists in step
    // (testing the key inside the walk loop left ~4 of 32 lanes busy)
    unsigned short dart[MAX_CODE];
    int nd = 0;
    for (int f = 0; f < nf; ++f)
        for (int e = 0; e < 3; ++e) {
            const int s = facets[f][e], t = facets[f][(e + 1) % 3];
            if (dart_key(R, s, t, facets[f][(e + 2) % 3]) == top && nd < MAX_CODE) dart[nd++] = (unsigned short)(s | (t << 8));
        }
    for (int i = 0; i < nd; ++i) {
        const int s = dart[i] & 255, t = dart[i] >> 8;
        int len;
        const int r = dart_code(n, R, s, t, clab, cur, len, best, best_len, n_inner);
        if (r == -2) return false;
        if (r == 1) {
            best_len = len;
            signed char *x = best;
            best = cur;
            cur = x;
            x = blab;
            blab = clab;
            clab = x;
        }
    }
    if (best_len < 0) return false;
    for (int u = 0; u < n; ++u) best_label[u] = blab[u];
    hash = code_hash(best, best_len);
    return true;
}

// ---------------------------------------------------------------------------------------------
// 3c. optimal rotation (unit quaternion, w first) taking template points onto observed points.
// A[3*a+b] = sum_i tpl_i[a] * obs_i[b].  Largest eigenvalue of Horn's 4x4 matrix by Newton iteration on
// its characteristic quartic, eigenvector from the adjugate (Theobald's QCP formulation).
MDB_HDN void optimal_rotation(const double *A, double E0, double *q)
{
    const double Sxx = A[0], Sxy = A[1], Sxz = A[2], Syx = A[3], Syy = A[4], Syz = A[5], Szx = A[6], Szy = A[7],
                 Szz = A[8];
    const double Sxx2 = Sxx * Sxx, Syy2 = Syy * Syy, Szz2 = Szz * Szz, Sxy2 = Sxy * Sxy, Syz2 = Syz * Syz,
                 Sxz2 = Sxz * Sxz, Syx2 = Syx * Syx, Szy2 = Szy * Szy, Szx2 = Szx * Szx;
    const double fn2 = Sxx2 + Syy2 + Szz2 + Sxy2 + Syz2 + Sxz2 + Syx2 + Szy2 + Szx2;
    const double t1 = 2.0 * (Syz * Szy - Syy * Szz);
    const double t2 = Syy2 + Szz2 - Sxx2 + Syz2 + Szy2;
    const double xzp = Sxz + Szx, yzp = Syz + Szy, xyp = Sxy + Syx, yzm = Syz - Szy, xzm = Sxz - Szx,
                 xym = Sxy - Syx, xxpyy = Sxx + Syy, xxmyy = Sxx - Syy;
    const double t3 = Sxy2 + Sxz2 - Syx2 - Szx2;
    const double C0 = t3 * t3 + (t2 + t1) * (t2 - t1) +
                      (-(xzp) * (yzm) + (xym) * (xxmyy - Szz)) * (-(xzm) * (yzp) + (xym) * (xxmyy + Szz)) +
                      (-(xzp) * (yzp) - (xyp) * (xxpyy - Szz)) * (-(xzm) * (yzm) - (xyp) * (xxpyy + Szz)) +
                      (+(xyp) * (yzp) + (xzp) * (xxmyy + Szz)) * (-(xym) * (yzm) + (xzp) * (xxpyy + Szz)) +
                      (+(xyp) * (yzm) + (xzm) * (xxmyy - Szz)) * (-(xym) * (yzp) + (xzm) * (xxpyy - Szz));
    const double C1 = 8.0 * (Sxx * Syz * Szy + Syy * Szx * Sxz + Szz * Sxy * Syx - Sxx * Syy * Szz - Syz * Szx * Sxy -
                             Szy * Syx * Sxz);
    const double C2 = -2.0 * fn2;
    double ev = E0;
    if (ev > 1e-11) {
        for (int it = 0; it < 50; ++it) {
            const double old = ev;
            const double x2 = ev * ev;
            const double bb = (x2 + C2) * ev;
            const double aa = bb + C1;
            ev -= (aa * ev + C0) / (2 * x2 * ev + bb + aa);
            if (fabs(ev - old) < fabs(1e-11 * ev)) break;
        }
    } else {
        ev = 0.0;
    }
    const double a11 = xxpyy + Szz - ev, a12 = yzm, a13 = -xzm, a14 = xym;
    const double a21 = yzm, a22 = xxmyy - Szz - ev, a23 = xyp, a24 = xzp;
    const double a31 = a13, a32 = a23, a33 = Syy - Sxx - Szz - ev, a34 = yzp;
    const double a41 = a14, a42 = a24, a43 = a34, a44 = Szz - xxpyy - ev;
    const double m3344 = a33 * a44 - a43 * a34, m3244 = a32 * a44 - a42 * a34, m3243 = a32 * a43 - a42 * a33,
                 m3143 = a31 * a43 - a41 * a33, m3144 = a31 * a44 - a41 * a34, m3142 = a31 * a42 - a41 * a32,
                 m1324 = a13 * a24 - a14 * a23, m1224 = a12 * a24 - a14 * a22, m1223 = a12 * a23 - a13 * a22,
                 m1124 = a11 * a24 - a14 * a21, m1123 = a11 * a23 - a13 * a21, m1122 = a11 * a22 - a12 * a21;
    double r[4][4];
    r[0][0] = a12 * m3344 - a13 * m3244 + a14 * m3243;
    r[0][1] = -a11 * m3344 + a13 * m3144 - a14 * m3143;
    r[0][2] = a11 * m3244 - a12 * m3144 + a14 * m3142;
    r[0][3] = -a11 * m3243 + a12 * m3143 - a13 * m3142;
    r[1][0] = a22 * m3344 - a23 * m3244 + a24 * m3243;
    r[1][1] = -a21 * m3344 + a23 * m3144 - a24 * m3143;
    r[1][2] = a21 * m3244 - a22 * m3144 + a24 * m3142;
    r[1][3] = -a21 * m3243 + a22 * m3143 - a23 * m3142;
    r[2][0] = a32 * m1324 - a33 * m1224 + a34 * m1223;
    r[2][1] = -a31 * m1324 + a33 * m1124 - a34 * m1123;
    r[2][2] = a31 * m1224 - a32 * m1124 + a34 * m1122;
    r[2][3] = -a31 * m1223 + a32 * m1123 - a33 * m1122;
    r[3][0] = a42 * m1324 - a43 * m1224 + a44 * m1223;
    r[3][1] = -a41 * m1324 + a43 * m1124 - a44 * m1123;
    r[3][2] = a41 * m1224 - a42 * m1124 + a44 * m1122;
    r[3][3] = -a41 * m1223 + a42 * m1123 - a43 * m1122;
    int bi = 0;
    double mx = 0;
    for (int i = 0; i < 4; ++i) {
        const double s = r[i][0] * r[i][0] + r[i][1] * r[i][1] + r[i][2] * r[i][2] + r[i][3] * r[i][3];
        if (s > mx) {
            mx = s;
            bi = i;
        }
    }
    if (mx < 1e-6) {
        q[0] = 1;
        q[1] = q[2] = q[3] = 0;
    } else {
        const double nq = sqrt(mx);
        for (int d = 0; d < 4; ++d) q[d] = r[bi][d] / nq;
    }
}

MDB_HD void quat_to_matrix(const double *q, double *u)
{
    const double a = q[0], b = q[1], c = q[2], d = q[3];
    u[0] = a * a + b * b - c * c - d * d;
    u[1] = 2 * b * c - 2 * a * d;
    u[2] = 2 * b * d + 2 * a * c;
    u[3] = 2 * b * c + 2 * a * d;
    u[4] = a * a - b * b + c * c - d * d;
    u[5] = 2 * c * d - 2 * a * b;
    u[6] = 2 * b * d - 2 * a * c;
    u[7] = 2 * c * d + 2 * a * b;
    u[8] = a * a - b * b - c * c + d * d;
}

MDB_HD void quat_mul(const double *r, const double *a, double *b)
{
    b[0] = r[0] * a[0] - r[1] * a[1] - r[2] * a[2] - r[3] * a[3];
    b[1] = r[0] * a[1] + r[1] * a[0] + r[2] * a[3] - r[3] * a[2];
    b[2] = r[0] * a[2] - r[1] * a[3] + r[2] * a[0] + r[3] * a[1];
    b[3] = r[0] * a[3] + r[1] * a[2] - r[2] * a[1] + r[3] * a[0];
}

// rotate q by the group element closest to its inverse, w >= 0 (ptm_quat.cpp:180-207)
MDB_HD void into_fundamental_zone(int ng, const double *gen, double *q)
{
    double mx = 0;
    int bi = 0;
    for (int i = 0; i < ng; ++i) {
        const double *g = gen + 4 * i;
        const double t = fabs(q[0] * g[0] - q[1] * g[1] - q[2] * g[2] - q[3] * g[3]);
        if (t > mx) {
            mx = t;
            bi = i;
        }
    }
    double f[4];
    quat_mul(q, gen + 4 * bi, f);
    const double sgn = f[0] < 0 ? -1.0 : 1.0;
    for (int d = 0; d < 4; ++d) q[d] = sgn * f[d];
}

// ---------------------------------------------------------------------------------------------
// result of one atom
struct Result {
    int type;       // reference structure id, 0 = none
    int ordering;   // alloy ordering (ptm_constants.h): 0 none, 1 pure, 2 L1_0, 3 L1_2(Cu), 4 L1_2(Au), 5 B2
    double rmsd, scale, q[4], interatomic_distance;
    int struct_index;            // S_* of the winner, -1 none
    signed char mapping[MAX_NB + 1];  // template point -> environment point (0 = centre)
    // two-shell winners (diamond, graphene): atom index and type of every environment point
    int env_idx[MAX_NB + 1], env_type[MAX_NB + 1];
};

// try every template triangulation of structure s whose hash matches
MDB_HDN void check_structure(const Tables &T, int s, unsigned long long hash, const signed char *env_label,
                            const double (*centred)[3], Result &res)
{
    const int n = T.n_nbrs[s], np = n + 1;
    signed char inv[MAX_NB];
    for (int u = 0; u < n; ++u) inv[env_label[u]] = (signed char)u;
    double G1 = 0, G2 = 0;
    for (int i = 0; i < np; ++i) {
        G1 += T.tpl[s][i][0] * T.tpl[s][i][0] + T.tpl[s][i][1] * T.tpl[s][i][1] + T.tpl[s][i][2] * T.tpl[s][i][2];
        G2 += centred[i][0] * centred[i][0] + centred[i][1] * centred[i][1] + centred[i][2] * centred[i][2];
    }
    const double E0 = (G1 + G2) / 2;
    // binary search for the first graph with this hash
    int lo = T.graph_begin[s], hi = T.graph_begin[s + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (T.hash[mid] < hash) lo = mid + 1;
        else hi = mid;
    }
    for (int g = lo; g < T.graph_begin[s + 1] && T.hash[g] == hash; ++g) {
        for (int au = T.aut_begin[g]; au < T.aut_begin[g + 1]; ++au) {
            const signed char *tl = T.aut_label + (size_t)au * MAX_NB;
            signed char mapping[MAX_NB + 1];
            mapping[0] = 0;
            for (int p = 0; p < n; ++p) mapping[p + 1] = (signed char)(inv[tl[p]] + 1);
            double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < np; ++i) {
                const double *t = T.tpl[s][i];
                const double *o = centred[mapping[i]];
                A[0] += t[0] * o[0];
                A[1] += t[0] * o[1];
                A[2] += t[0] * o[2];
                A[3] += t[1] * o[0];
                A[4] += t[1] * o[1];
                A[5] += t[1] * o[2];
                A[6] += t[2] * o[0];
                A[7] += t[2] * o[1];
                A[8] += t[2] * o[2];
            }
            double q[4], rot[9];
            optimal_rotation(A, E0, q);
            quat_to_matrix(q, rot);
            double k0 = 0;
            for (int i = 0; i < np; ++i) {
                const double *t = T.tpl[s][i];
                const double *o = centred[mapping[i]];
                for (int j = 0; j < 3; ++j) {
                    double v = 0.0;
                    for (int k = 0; k < 3; ++k) v += rot[j * 3 + k] * t[k];
                    k0 += v * o[j];
                }
            }
            const double scale = k0 / G2;
            const double rmsd = sqrt(fabs(G1 - scale * k0) / np);
            if (rmsd < res.rmsd) {
                res.rmsd = rmsd;
                res.scale = scale;
                res.struct_index = s;
                for (int d = 0; d < 4; ++d) res.q[d] = q[d];
                for (int i = 0; i < np; ++i) res.mapping[i] = mapping[i];
            }
        }
    }
}

// hull + canonical form of the first n ordered points, then the listed structures (all with n neighbours)
MDB_HDN void match_shell(const Tables &T, const int *structs, int ns, const double (*hull_pts)[3],
                        const double (*raw_pts)[3], Result &res)
{
    const int s0 = structs[0];
    const int n = T.n_nbrs[s0], np = n + 1;
    signed char facets[MAX_FACETS][3];
    const int nf = convex_hull(np, hull_pts, facets);
    if (nf != T.n_facets[s0]) return;
    Rotation R;
    if (!build_rotation(n, nf, facets, R)) return;
    int maxdeg = 0;
    for (int u = 0; u < n; ++u) maxdeg = maxdeg > R.deg[u] ? maxdeg : R.deg[u];
    if (maxdeg > T.max_degree[s0]) return;
    if (s0 == S_SC)
        for (int u = 0; u < n; ++u)
            if (R.deg[u] != 4) return;
    // barycentre of the n+1 raw points removed (ptm_normalize_vertices.cpp:23-44)
    double centred[MAX_NB + 1][3];
    double sum[3] = {0, 0, 0};
    for (int i = 0; i < np; ++i)
        for (int d = 0; d < 3; ++d) sum[d] += raw_pts[i][d];
    for (int d = 0; d < 3; ++d) sum[d] /= np;
    for (int i = 0; i < np; ++i)
        for (int d = 0; d < 3; ++d) centred[i][d] = raw_pts[i][d] - sum[d];
    signed char label[MAX_NB];
    unsigned long long hash;
    if (!canonical_form(n, nf, facets, R, label, hash)) return;
    for (int k = 0; k < ns; ++k) check_structure(T, structs[k], hash, label, centred, res);
}

// barycentre of the np points removed (ptm_normalize_vertices.cpp:21-44)
MDB_HD void subtract_barycentre(int np, const double (*raw)[3], double (*out)[3])
{
    double sum[3] = {0, 0, 0};
    for (int i = 0; i < np; ++i)
        for (int d = 0; d < 3; ++d) sum[d] += raw[i][d];
    for (int d = 0; d < 3; ++d) sum[d] /= np;
    for (int i = 0; i < np; ++i)
        for (int d = 0; d < 3; ++d) out[i][d] = raw[i][d] - sum[d];
}

// hull coordinates: barycentre removed, then divided by (sum of the neighbour lengths) / np
// (ptm_normalize_vertices.cpp:46-66 -- the divisor counts the centre as well)
MDB_HD void normalize_vertices(int np, const double (*raw)[3], double (*out)[3])
{
    subtract_barycentre(np, raw, out);
    double scale = 0;
    for (int i = 1; i < np; ++i) scale += sqrt(dot3(out[i], out[i]));
    scale /= np;
    for (int i = 0; i < np; ++i)
        for (int d = 0; d < 3; ++d) out[i][d] /= scale;
}

// Two-shell environment of the diamond structures (ptm_structure_matcher.cpp:193-310).  Points: centre,
// 4 first neighbours ("inner", vertices 0..3 of the graph), 12 second neighbours (vertices 4..15, three
// per inner atom).  The hull of the 16 neighbours normally consists of the outer atoms only; every inner
// atom is then put back as the apex over the facet spanned by its own three outer atoms, which yields
// the 28-facet graph the templates are tabulated with.  An inner atom that does reach the hull
// ("inverted") already brings its three facets along.
MDB_HDN void match_diamond(const Tables &T, int flags, const double (*hull_pts)[3], const double (*raw_pts)[3],
                          Result &res)
{
    const int n = 16, np = 17;
    signed char facets[MAX_FACETS][3];
    int nf = convex_hull(np, hull_pts, facets);
    if (nf < 0) return;
    bool inverted[4] = {false, false, false, false};
    for (int f = 0; f < nf; ++f) {
        int cnt = 0;
        for (int e = 0; e < 3; ++e)
            if (facets[f][e] <= 3) {
                inverted[facets[f][e]] = true;
                ++cnt;
            }
        if (cnt > 1) return;  // a facet with two inner atoms
    }
    int n_inv = 0;
    for (int i = 0; i < 4; ++i) n_inv += inverted[i] ? 1 : 0;
    if (nf != 20 + 2 * n_inv) return;
    int n_found = 0;
    signed char toadd[4][3];
    for (int f = 0; f < nf; ++f) {
        const int a = facets[f][0], b = facets[f][1], c = facets[f][2];
        if (a <= 3 || b <= 3 || c <= 3) continue;
        const int i0 = (a - 4) / 3, i1 = (b - 4) / 3, i2 = (c - 4) / 3;
        if (i0 == i1 && i0 == i2) {
            if (n_found + n_inv >= 4) return;
            toadd[n_found][0] = (signed char)a;
            toadd[n_found][1] = (signed char)b;
            toadd[n_found][2] = (signed char)c;
            ++n_found;
            for (int e = 0; e < 3; ++e) facets[f][e] = facets[nf - 1][e];
            --nf;
            --f;
        }
    }
    if (n_found + n_inv != 4) return;
    for (int k = 0; k < n_found; ++k) {
        const signed char a = toadd[k][0], b = toadd[k][1], c = toadd[k][2];
        const signed char i0 = (signed char)((a - 4) / 3);
        if (nf + 3 > MAX_FACETS) return;
        facets[nf][0] = i0, facets[nf][1] = b, facets[nf][2] = c;
        ++nf;
        facets[nf][0] = a, facets[nf][1] = i0, facets[nf][2] = c;
        ++nf;
        facets[nf][0] = a, facets[nf][1] = b, facets[nf][2] = i0;
        ++nf;
    }
    Rotation R;
    if (!build_rotation(n, nf, facets, R)) return;
    int maxdeg = 0;
    for (int u = 0; u < n; ++u) maxdeg = maxdeg > R.deg[u] ? maxdeg : R.deg[u];
    if (maxdeg > T.max_degree[S_DCUB]) return;
    double centred[MAX_NB + 1][3];
    subtract_barycentre(np, raw_pts, centred);
    signed char label[MAX_NB];
    unsigned long long hash;
    if (!canonical_form(n, nf, facets, R, label, hash, 4)) return;
    if (flags & CHECK_DCUB) check_structure(T, S_DCUB, hash, label, centred, res);
    if (flags & CHECK_DHEX) check_structure(T, S_DHEX, hash, label, centred, res);
}

// RMSD of one fixed correspondence (ptm_structure_matcher.cpp:29-56); mapping: template point -> env point
MDB_HDN void try_mapping(const Tables &T, int s, const double (*centred)[3], const signed char *mapping, double G1,
                        double G2, Result &res)
{
    const int np = T.n_nbrs[s] + 1;
    const double E0 = (G1 + G2) / 2;
    double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < np; ++i) {
        const double *t = T.tpl[s][i];
        const double *o = centred[mapping[i]];
        A[0] += t[0] * o[0];
        A[1] += t[0] * o[1];
        A[2] += t[0] * o[2];
        A[3] += t[1] * o[0];
        A[4] += t[1] * o[1];
        A[5] += t[1] * o[2];
        A[6] += t[2] * o[0];
        A[7] += t[2] * o[1];
        A[8] += t[2] * o[2];
    }
    double q[4], rot[9];
    optimal_rotation(A, E0, q);
    quat_to_matrix(q, rot);
    double k0 = 0;
    for (int i = 0; i < np; ++i) {
        const double *t = T.tpl[s][i];
        const double *o = centred[mapping[i]];
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += rot[j * 3 + k] * t[k];
            k0 += v * o[j];
        }
    }
    const double scale = k0 / G2;
    const double rmsd = sqrt(fabs(G1 - scale * k0) / np);
    if (rmsd < res.rmsd) {
        res.rmsd = rmsd;
        res.scale = scale;
        res.struct_index = s;
        for (int d = 0; d < 4; ++d) res.q[d] = q[d];
        for (int i = 0; i < np; ++i) res.mapping[i] = mapping[i];
    }
}

// graphene: 3 + 3 x 2 neighbours, planar, no hull: the two outer atoms of every inner atom are tried in
// both orders (ptm_structure_matcher.cpp:348-380, same sequence of the eight correspondences)
MDB_HD void match_graphene(const Tables &T, const double (*raw_pts)[3], Result &res)
{
    const int np = 10;
    double centred[MAX_NB + 1][3];
    subtract_barycentre(np, raw_pts, centred);
    double G1 = 0, G2 = 0;
    for (int i = 0; i < np; ++i) {
        G1 += T.tpl[S_GRAPHENE][i][0] * T.tpl[S_GRAPHENE][i][0] + T.tpl[S_GRAPHENE][i][1] * T.tpl[S_GRAPHENE][i][1] +
              T.tpl[S_GRAPHENE][i][2] * T.tpl[S_GRAPHENE][i][2];
        G2 += centred[i][0] * centred[i][0] + centred[i][1] * centred[i][1] + centred[i][2] * centred[i][2];
    }
    signed char mapping[MAX_NB + 1];
    for (int i = 0; i < np; ++i) mapping[i] = (signed char)i;
    auto swp = [&](int a, int b) {
        const signed char t = mapping[a];
        mapping[a] = mapping[b];
        mapping[b] = t;
    };
    for (int i = 0; i < 2; ++i) {
        swp(4, 5);
        for (int j = 0; j < 2; ++j) {
            swp(6, 7);
            for (int k = 0; k < 2; ++k) {
                swp(8, 9);
                try_mapping(T, S_GRAPHENE, centred, mapping, G1, G2, res);
            }
        }
    }
}

// Two-shell neighbour ordering (ptm_multishell.cpp:94-184).  Src gives access to other atoms:
//   int gather(int atom, double (*pts)[3], int *nbr)   listed neighbours (vectors, indices), returns their number
//   const unsigned char *order_of(int atom)            rank -> list position from the pre-ordering pass (255 = none)
//   int type_of(int atom)
// Only the MAX_MULTISHELL nearest listed neighbours of an atom take part.  Returns false when the
// environment cannot be completed.  out_*: centre, n_inner first neighbours, then n_outer per inner.
struct ShellEnv {
    double pts[MAX_NB + 1][3];
    int idx[MAX_NB + 1], type[MAX_NB + 1];
};

template <class Src>
MDB_HDN bool two_shell_env(const Src &src, int atom, int num, const double (*pts)[3], const int *nbr, const int *order,
                          const int *types, int n_inner, int n_outer, ShellEnv &out)
{
    // filtered, ordered first shell
    int kept = 0;
    out.pts[0][0] = out.pts[0][1] = out.pts[0][2] = 0;
    out.idx[0] = atom;
    out.type[0] = types[0];
    for (int r = 0; r < num; ++r) {
        const int p = order[r];
        if (p + 1 > MAX_MULTISHELL) continue;
        if (kept < n_inner) {
            for (int d = 0; d < 3; ++d) out.pts[1 + kept][d] = pts[p][d];
            out.idx[1 + kept] = nbr[p];
            out.type[1 + kept] = types[1 + p];
        }
        ++kept;
    }
    if (kept + 1 < n_inner + 1) return false;
    double tol = 1e-5 * sqrt(dot3(out.pts[1], out.pts[1]));
    tol = tol > 1e-5 ? tol : 1e-5;
    // filtered, ordered neighbours of every inner atom, relative to the centre
    struct Cand {
        double d[3];
        int idx, type;
    };
    Cand cand[4][MAX_MULTISHELL];
    int ncand[4] = {0, 0, 0, 0};
    for (int i = 0; i < n_inner; ++i) {
        double ip[MAX_IN][3];
        int inb[MAX_IN];
        const int a = out.idx[1 + i];
        const int m = src.gather(a, ip, inb);
        const unsigned char *ord = src.order_of(a);
        int c = 0;
        for (int r = 0; r < m; ++r) {
            const int p = ord[r];
            if (p >= m || p + 1 > MAX_MULTISHELL) continue;
            for (int d = 0; d < 3; ++d) cand[i][c].d[d] = ip[p][d] + out.pts[1 + i][d];
            cand[i][c].idx = inb[p];
            cand[i][c].type = src.type_of(inb[p]);
            ++c;
        }
        ncand[i] = c;
        if (c + 1 < n_inner + 1) return false;
    }
    // rank-major sweep (the reference's stable sort by rank keeps the inner atoms in order within a rank)
    int counts[4] = {0, 0, 0, 0};
    int found = 0;
    const int want = n_inner * n_outer;
    auto claimed = [&](int idx, const double *d) {
        auto near = [&](int slot) {
            if (out.idx[slot] != idx) return false;
            const double dx = d[0] - out.pts[slot][0], dy = d[1] - out.pts[slot][1], dz = d[2] - out.pts[slot][2];
            return sqrt(dx * dx + dy * dy + dz * dz) < tol;
        };
        for (int s = 0; s < n_inner + 1; ++s)
            if (near(s)) return true;
        for (int i = 0; i < n_inner; ++i)
            for (int j = 0; j < counts[i]; ++j)
                if (near(1 + n_inner + n_outer * i + j)) return true;
        return false;
    };
    for (int r = 0; r < MAX_MULTISHELL && found < want; ++r)
        for (int i = 0; i < n_inner && found < want; ++i) {
            if (r >= ncand[i] || counts[i] >= n_outer) continue;
            const Cand &c = cand[i][r];
            if (claimed(c.idx, c.d)) continue;
            const int slot = 1 + n_inner + n_outer * i + counts[i];
            for (int d = 0; d < 3; ++d) out.pts[slot][d] = c.d[d];
            out.idx[slot] = c.idx;
            out.type[slot] = c.type;
            ++counts[i];
            ++found;
        }
    return found == want;
}

// diamond (4 + 4 x 3) and graphene (3 + 3 x 2) environments; kept out of line so that the common
// one-shell path does not pay for its registers and stack
template <class Src>
MDB_HDN void match_two_shell(const Tables &T, int flags, int num, const double (*pts)[3], const int *order,
                            const int *types, const int *nbr, const Src &src, int atom, Result &res)
{
    ShellEnv env;
    if (flags & (CHECK_DCUB | CHECK_DHEX)) {   // ptm_index.cpp:159-170
        if (two_shell_env(src, atom, num, pts, nbr, order, types, 4, 3, env)) {
            double hull[MAX_NB + 1][3];
            normalize_vertices(17, env.pts, hull);
            const double before = res.rmsd;
            match_diamond(T, flags, hull, env.pts, res);
            if (res.rmsd < before)
                for (int p = 0; p < 17; ++p) {
                    res.env_idx[p] = env.idx[p];
                    res.env_type[p] = env.type[p];
                }
        }
    }
    if (flags & CHECK_GRAPHENE) {             // ptm_index.cpp:172-179
        if (two_shell_env(src, atom, num, pts, nbr, order, types, 3, 2, env)) {
            const double before = res.rmsd;
            match_graphene(T, env.pts, res);
            if (res.rmsd < before)
                for (int p = 0; p < 10; ++p) {
                    res.env_idx[p] = env.idx[p];
                    res.env_type[p] = env.type[p];
                }
        }
    }
}

// full per-atom analysis.  pts[0..num): neighbour vectors in list (distance) order; order[0..num): their
// ranking from preorder_neighbours; types: atom type of the centre (types[0]) and of each listed
// neighbour (types[1 + k]).
// nbr: atom index of every listed neighbour; src / atom: access to the neighbours' own lists and rankings
// for the two-shell structures (see two_shell_env).
template <class Src>
MDB_HD void match_atom(const Tables &T, int flags, int num, const double (*pts)[3], const int *order, const int *types,
                       const int *nbr, const Src &src, int atom, Result &res)
{
    res.type = 0;
    res.ordering = 0;
    res.rmsd = 1e300;
    res.scale = 0;
    res.interatomic_distance = 0;
    res.struct_index = -1;
    res.q[0] = res.q[1] = res.q[2] = res.q[3] = 0;
    // ordered points, 0 = centre
    double raw[MAX_IN + 1][3];
    raw[0][0] = raw[0][1] = raw[0][2] = 0;
    for (int i = 0; i < num; ++i)
        for (int d = 0; d < 3; ++d) raw[i + 1][d] = pts[order[i]][d];
    const int np_all = num + 1;
    // hull coordinates: barycentre of ALL offered points removed, mean length 1 (ptm_index.cpp:147-149)
    double hull[MAX_IN + 1][3];
    {
        double sum[3] = {0, 0, 0};
        for (int i = 0; i < np_all; ++i)
            for (int d = 0; d < 3; ++d) sum[d] += raw[i][d];
        for (int d = 0; d < 3; ++d) sum[d] /= np_all;
        double scale = 0;
        for (int i = 0; i < np_all; ++i)
            for (int d = 0; d < 3; ++d) hull[i][d] = raw[i][d] - sum[d];
        for (int i = 1; i < np_all; ++i) scale += sqrt(dot3(hull[i], hull[i]));
        scale /= np_all;
        for (int i = 0; i < np_all; ++i)
            for (int d = 0; d < 3; ++d) hull[i][d] /= scale;
    }
    if ((flags & CHECK_SC) && np_all >= 7) {
        const int st[1] = {S_SC};
        match_shell(T, st, 1, hull, raw, res);
    }
    if ((flags & (CHECK_FCC | CHECK_HCP | CHECK_ICO)) && np_all >= 13) {
        int st[3], ns = 0;
        if (flags & CHECK_FCC) st[ns++] = S_FCC;
        if (flags & CHECK_HCP) st[ns++] = S_HCP;
        if (flags & CHECK_ICO) st[ns++] = S_ICO;
        match_shell(T, st, ns, hull, raw, res);
    }
    if ((flags & CHECK_BCC) && np_all >= 15) {
        const int st[1] = {S_BCC};
        match_shell(T, st, 1, hull, raw, res);
    }
    // one-shell winners so far: environment = the ranked list itself
    if (res.struct_index >= 0) {
        const int n1 = T.n_nbrs[res.struct_index];
        res.env_idx[0] = atom;
        res.env_type[0] = types[0];
        for (int p = 1; p <= n1; ++p) {
            res.env_idx[p] = nbr[order[p - 1]];
            res.env_type[p] = types[1 + order[p - 1]];
        }
    }
    if (flags & (CHECK_DCUB | CHECK_DHEX | CHECK_GRAPHENE))
        match_two_shell(T, flags, num, pts, order, types, nbr, src, atom, res);
    if (res.struct_index < 0) {
        res.rmsd = 0;
        return;
    }
    const int s = res.struct_index;
    res.type = T.type_id[s];
    const int gi = T.group[s];
    into_fundamental_zone(T.gen_begin[gi + 1] - T.gen_begin[gi], T.gen + 4 * T.gen_begin[gi], res.q);
    res.interatomic_distance = T.c_dist[s] / res.scale;
    // alloy ordering (ptm_alloy_types.cpp:92-121) from the types of the matched points
    const int n = T.n_nbrs[s];
    const int t0 = res.env_type[0];
    bool pure = true, binary = true;
    int other = -1;
    unsigned diff = 0;  // bit p: template point p carries a different type than the centre
    for (int p = 1; p <= n; ++p) {
        const int tp = res.env_type[res.mapping[p]];
        if (tp != t0) {
            pure = false;
            diff |= 1u << p;
            if (other == -1) other = tp;
            else if (tp != other) binary = false;
        }
    }
    if (pure) res.ordering = 1;
    else if (!binary) res.ordering = 0;
    else if (s == S_FCC) {
        const unsigned all = 0x1ffeu;
        if (diff == all) res.ordering = 4;
        else
            for (int a = 0; a < 3; ++a) {
                if (diff == T.fcc_plane[a]) res.ordering = 3;             // four unlike neighbours in one {100} plane
                if (diff == (all & ~T.fcc_plane[a])) res.ordering = 2;    // four like neighbours in one {100} plane
            }
    } else if (s == S_BCC) {
        if (diff == 0x1feu) res.ordering = 5;  // first shell unlike, second shell like
    } else if (s == S_DCUB || s == S_DHEX) {
        if (diff == 0x1eu) res.ordering = 6;   // SiC: the four first neighbours unlike, second shell like
    } else if (s == S_GRAPHENE) {
        if (diff == 0xeu) res.ordering = 7;    // BN
    }
}


}  // namespace ptm
