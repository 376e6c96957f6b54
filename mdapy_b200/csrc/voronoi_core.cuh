// mdapy_b200/csrc/voronoi_core.cuh -- per-atom Voronoi cell construction (see voronoi.cu for the method and the
// reference lines).  __host__ __device__ so tests/host/voronoi_host_harness.cpp can run the same arithmetic on the
// CPU; the product only ever runs it inside k_voronoi.
#pragma once
#include "box.cuh"

namespace voro {
constexpr int VT = 128;   // vertices (dual triangles) per cell
constexpr int VP = 64;    // planes per cell, walls included
constexpr int VE = 128;   // boundary edges / removed vertices of one cut (a cut can remove most of a large cell)
constexpr int VB = 64;    // threads per block

template <class Rec> struct VoroArgs {
    const Rec *sorted;
    const int *cell_start;
    int N;
    DBox box;
    CellGrid g;
    double w;       // cell width of the grid (the last cell of an axis takes the remainder)
    double L[3];    // extent of the grid axes: box lengths (orthogonal) / perpendicular thicknesses (triclinic)
    double R0;      // triclinic: half-width of the initial cube (the atom's own images bound the cell inside it)
    int wrapped;    // the records hold coordinates already wrapped into the box
    int has_open;   // some axis is not periodic
    double tolh;    // half of voro++'s tolerance on (2 n.v - |r|^2)
    double *volume;
    int *nfaces;
    double *radius;
    int *row_id;
    double *row_area;
    int W;          // row width of row_id / row_area (0: no rows wanted)
    int *status;    // [0] max faces, [1] cells that outgrew the buffers
};

struct Cell {
    double px[VP], py[VP], pz[VP], pd[VP];
    int pid[VP];
    unsigned char ta[VT], tb[VT], tc[VT];
    double vx[VT], vy[VT], vz[VT];
    unsigned long long em[VP];   // scratch of clip(): edge matrix of the removed triangles
    int np, nt, fail;
    int careful;   // 0: new vertices solved in place; 1: appended and refined (after an ill-conditioned solve was seen)
    int ill;       // fast mode met a vertex whose three planes almost share a line
    double rmax2;

    // Position of vertex t from its three planes (Cramer).  Returns the squared normalised determinant of the three
    // normals: near zero the planes almost share a line and the solve loses accuracy (see refine()).
    MDB_HD double vertex(int t)
    {
        const int a = ta[t], b = tb[t], c = tc[t];
        const double ax = px[a], ay = py[a], az = pz[a];
        const double bx = px[b], by = py[b], bz = pz[b];
        const double cx = px[c], cy = py[c], cz = pz[c];
        const double bcx = by * cz - bz * cy, bcy = bz * cx - bx * cz, bcz = bx * cy - by * cx;
        const double cax = cy * az - cz * ay, cay = cz * ax - cx * az, caz = cx * ay - cy * ax;
        const double abx = ay * bz - az * by, aby = az * bx - ax * bz, abz = ax * by - ay * bx;
        const double det = ax * bcx + ay * bcy + az * bcz;
        const double inv = 1.0 / det;
        const double da = pd[a], db = pd[b], dc = pd[c];
        vx[t] = (da * bcx + db * cax + dc * abx) * inv;
        vy[t] = (da * bcy + db * cay + dc * aby) * inv;
        vz[t] = (da * bcz + db * caz + dc * abz) * inv;
        return det * det / ((ax * ax + ay * ay + az * az) * (bx * bx + by * by + bz * bz) * (cx * cx + cy * cy + cz * cz));
    }
    // An ill-conditioned new vertex (a, b, p): it lies on the cell edge shared by planes a and b, between the removed
    // vertex tu and the kept vertex across the edge (the one holding the opposite dual edge (b, a)); interpolate along
    // that edge to the cutting plane instead, as voro++ places its new vertices.
    MDB_HD void refine(int t, int tu, int limit, double nx, double ny, double nz, double d)
    {
        const int a = ta[t], b = tb[t];
        int tv = -1;
        for (int u = 0; u < limit && tv < 0; ++u)
            if ((ta[u] == b && tb[u] == a) || (tb[u] == b && tc[u] == a) || (tc[u] == b && ta[u] == a)) tv = u;
        if (tv < 0) return;
        const double su = nx * vx[tu] + ny * vy[tu] + nz * vz[tu] - d, sv = nx * vx[tv] + ny * vy[tv] + nz * vz[tv] - d;
        if (!(su > 0.0) || !(sv < 0.0)) return;   // the edge does not cross the plane cleanly: keep the solve
        const double w = sv / (sv - su);          // V + w (U - V) lies on the plane
        vx[t] = vx[tv] + w * (vx[tu] - vx[tv]);
        vy[t] = vy[tv] + w * (vy[tu] - vy[tv]);
        vz[t] = vz[tv] + w * (vz[tu] - vz[tv]);
    }
    MDB_HD void plane(int p, double nx, double ny, double nz, double d, int id)
    {
        px[p] = nx, py[p] = ny, pz[p] = nz, pd[p] = d, pid[p] = id;
    }
    MDB_HD void update_rmax()
    {
        double m = 0.0;
        for (int t = 0; t < nt; ++t) m = fmax(m, vx[t] * vx[t] + vy[t] * vy[t] + vz[t] * vz[t]);
        rmax2 = m;
    }
    // the box lo <= x <= hi around the atom at the origin; wall ids as voro++ (-1 .. -6)
    MDB_HD void init(const double *lo, const double *hi)
    {
        plane(0, -1, 0, 0, -lo[0], -1);
        plane(1, 1, 0, 0, hi[0], -2);
        plane(2, 0, -1, 0, -lo[1], -3);
        plane(3, 0, 1, 0, hi[1], -4);
        plane(4, 0, 0, -1, -lo[2], -5);
        plane(5, 0, 0, 1, hi[2], -6);
        np = 6;
        // the eight corners, each triple ordered so that the normals form a right-handed frame
        const unsigned char T[8][3] = {{0, 2, 4}, {1, 4, 2}, {0, 4, 3}, {1, 3, 4}, {0, 5, 2}, {1, 2, 5}, {0, 3, 5}, {1, 5, 3}};
        for (int t = 0; t < 8; ++t) {
            ta[t] = T[t][0], tb[t] = T[t][1], tc[t] = T[t][2];
            vertex(t);
        }
        nt = 8;
        fail = 0;
        ill = 0;
        for (int q = 0; q < VP; ++q) em[q] = 0;
        update_rmax();
    }
    // drop planes no vertex refers to (they were cut away), keeping the insertion order
    MDB_HD void collect()
    {
        unsigned long long used = 0;
        for (int t = 0; t < nt; ++t) used |= (1ull << ta[t]) | (1ull << tb[t]) | (1ull << tc[t]);
        unsigned char map[VP];
        int w = 0;
        for (int p = 0; p < np; ++p)
            if (used >> p & 1ull) {
                map[p] = (unsigned char)w;
                if (w != p) plane(w, px[p], py[p], pz[p], pd[p], pid[p]);
                ++w;
            }
        np = w;
        for (int t = 0; t < nt; ++t) ta[t] = map[ta[t]], tb[t] = map[tb[t]], tc[t] = map[tc[t]];
    }
    // cut with n.x <= d; returns true when the plane removed something
    MDB_HD bool clip(double nx, double ny, double nz, double d, int id, double tolh)
    {
        unsigned char rem[VE];   // vertices beyond the plane, ascending
        int nout = 0;
#ifdef VORO_COUNT
        ++g_clips, g_scan += nt;
#endif
        for (int t = 0; t < nt; ++t)
            if (nx * vx[t] + ny * vy[t] + nz * vz[t] - d > tolh) {
                if (nout < VE) rem[nout] = (unsigned char)t;
                ++nout;
            }
        if (nout == 0) return false;
        if (nout == nt || nout > VE) {   // nothing left (cannot happen for a bisector of a distinct point) / too many
            fail = 1;
            return false;
        }
        if (np == VP) collect();
        if (np == VP) {
            fail = 1;
            return false;
        }
        const int p = np++;
        plane(p, nx, ny, nz, d, id);
        // em[u] bit v: the removed set holds the oriented edge (u, v).  An edge of the set lies on its boundary when
        // the opposite edge is not in the set (the triangle across it stays).  em is all zero between calls.
        unsigned long long dup = 0;
        for (int k = 0; k < nout; ++k) {
            const int t = rem[k];
            const int a = ta[t], b = tb[t], c = tc[t];
            dup |= (em[a] >> b) & 1ull;
            em[a] |= 1ull << b;
            dup |= (em[b] >> c) & 1ull;
            em[b] |= 1ull << c;
            dup |= (em[c] >> a) & 1ull;
            em[c] |= 1ull << a;
        }
        unsigned char ea[VE], eb[VE], et[VE];   // boundary edges and the removed vertex each one belongs to
        int ne = 0;
        if (!dup) {
            for (int k = 0; k < nout; ++k) {
                const int t = rem[k];
                const unsigned char tri[4] = {ta[t], tb[t], tc[t], ta[t]};
                for (int e = 0; e < 3; ++e) {
                    const unsigned char u = tri[e], v = tri[e + 1];
                    if (em[v] >> u & 1ull) continue;
                    if (ne < VE) {
                        ea[ne] = u, eb[ne] = v, et[ne] = (unsigned char)t;
                        ++ne;
                    } else fail = 1;
                }
            }
        } else {
            // a tolerance decision left the same oriented edge in two triangles (near-degenerate input): pair the
            // edges one by one instead, an edge cancels exactly one opposite edge
            for (int k = 0; k < nout; ++k) {
                const int t = rem[k];
                const unsigned char tri[4] = {ta[t], tb[t], tc[t], ta[t]};
                for (int q = 0; q < 3; ++q) {
                    const unsigned char u = tri[q], v = tri[q + 1];
                    int hit = -1;
                    for (int e = 0; e < ne && hit < 0; ++e)
                        if (ea[e] == v && eb[e] == u) hit = e;
                    if (hit >= 0) {
                        --ne;
                        ea[hit] = ea[ne], eb[hit] = eb[ne], et[hit] = et[ne];
                    } else if (ne < VE) {
                        ea[ne] = u, eb[ne] = v, et[ne] = (unsigned char)t;
                        ++ne;
                    } else fail = 1;
                }
            }
        }
        for (int k = 0; k < nout; ++k) {
            const int t = rem[k];
            em[ta[t]] = 0, em[tb[t]] = 0, em[tc[t]] = 0;
        }
        if (fail || (careful ? nt + ne : nt - nout + ne) > VT) {
            fail = 1;
            return false;
        }
        if (!careful) {
            // the new vertices (boundary edge + new plane) take the freed slots; a cut that removes whole planes
            // frees more slots than it fills, and the tail of the array moves into those
            for (int e = 0; e < ne; ++e) {
                const int slot = e < nout ? rem[e] : nt++;
                ta[slot] = ea[e], tb[slot] = eb[e], tc[slot] = (unsigned char)p;
                if (vertex(slot) < 1.0e-8) ill = 1;   // the caller rebuilds this cell in careful mode
            }
            for (int k = nout - 1; k >= ne; --k) {
                const int h = rem[k], last = nt - 1;
                if (h != last) ta[h] = ta[last], tb[h] = tb[last], tc[h] = tc[last], vx[h] = vx[last], vy[h] = vy[last], vz[h] = vz[last];
                --nt;
            }
        } else {
            // the new vertices go behind the array while the removed ones are still in place (refine() reads them),
            // then the tail moves into the freed slots
            const int base = nt;
            for (int e = 0; e < ne; ++e) {
                const int slot = base + e;
                ta[slot] = ea[e], tb[slot] = eb[e], tc[slot] = (unsigned char)p;
                if (vertex(slot) < 1.0e-8) refine(slot, et[e], base, nx, ny, nz, d);
            }
            nt = base + ne;
            for (int k = nout - 1; k >= 0; --k) {
                const int h = rem[k], last = nt - 1;
                if (h != last) ta[h] = ta[last], tb[h] = tb[last], tc[h] = tc[last], vx[h] = vx[last], vy[h] = vy[last], vz[h] = vz[last];
                --nt;
            }
        }
        update_rmax();
#ifdef VORO_COUNT
        ++g_accept;
#endif
        return true;
    }
    // area of the face of plane p (vertices visited around the plane node), 0 for a plane without vertices;
    // perim receives the length of its outline
    MDB_HD double face_area(int p, double &perim) const
    {
        perim = 0.0;
        int t0 = -1;
        for (int t = 0; t < nt && t0 < 0; ++t)
            if (ta[t] == p || tb[t] == p || tc[t] == p) t0 = t;
        if (t0 < 0) return 0.0;
        const double x0 = vx[t0], y0 = vy[t0], z0 = vz[t0];
        double sx = 0, sy = 0, sz = 0, ux = 0, uy = 0, uz = 0;
        int cur = t0, steps = 0;
        bool first = true;
        while (steps++ < VT) {
            // the plane that follows p's successor in cur: cur = (p, b, c) up to rotation -> next holds (p, c, .)
            const int c = ta[cur] == p ? tc[cur] : (tb[cur] == p ? ta[cur] : tb[cur]);
            int nxt = -1;
            for (int t = 0; t < nt; ++t)
                if ((ta[t] == p && tb[t] == c) || (tb[t] == p && tc[t] == c) || (tc[t] == p && ta[t] == c)) {
                    nxt = t;
                    break;
                }
            if (nxt < 0 || nxt == t0) break;
            const double wx = vx[nxt] - x0, wy = vy[nxt] - y0, wz = vz[nxt] - z0;
            perim += sqrt((wx - ux) * (wx - ux) + (wy - uy) * (wy - uy) + (wz - uz) * (wz - uz));
            if (!first) {
                sx += uy * wz - uz * wy;
                sy += uz * wx - ux * wz;
                sz += ux * wy - uy * wx;
            }
            first = false;
            ux = wx, uy = wy, uz = wz;
            cur = nxt;
        }
        perim += sqrt(ux * ux + uy * uy + uz * uz);   // back to the first vertex
        return 0.5 * sqrt(sx * sx + sy * sy + sz * sz);
    }
};

MDB_HD int floor_div(int a, int n) { return a >= 0 ? a / n : -((-a + n - 1) / n); }

// [lo, hi] of cell k of an axis (k outside [0, n): a periodic image), relative to the box origin
MDB_HD void cell_span(int k, int n, double w, double L, double &lo, double &hi)
{
    const int m = floor_div(k, n), kk = k - m * n;
    // a grid padded to three cells can be wider than the box: the cells past L are empty and have zero width there
    lo = m * L + fmin(kk * w, L);
    hi = kk == n - 1 ? m * L + L : m * L + fmin((kk + 1) * w, L);
}

// One axis of a grid cell seen from the atom: k may lie outside [0, n) (a periodic image).
struct AxisCell {
    bool valid;
    int kk;         // cell index inside the grid
    double gap2;    // squared distance from the atom's slab coordinate to the cell's slab
    double s[3];    // image shift (Cartesian)
};
template <class Rec> MDB_HD AxisCell axis_cell(const VoroArgs<Rec> &A, int d, int k, double pd)
{
    AxisCell a;
    const int n = A.g.n[d];
    a.valid = A.box.pbc[d] || (k >= 0 && k < n);
    const int m = floor_div(k, n);
    a.kk = k - m * n;
    double lo, hi;
    cell_span(k, n, A.w, A.L[d], lo, hi);
    const double gap = fmax(0.0, fmax(lo - pd, pd - hi));
    a.gap2 = gap * gap;
    a.s[0] = m * A.box.h[3 * d], a.s[1] = m * A.box.h[3 * d + 1], a.s[2] = m * A.box.h[3 * d + 2];
    return a;
}

// Cell of the atom at position s of the cell-sorted order.  Returns its face count, -1 when the cell outgrew the
// buffers, 0 for an atom outside an open container.
template <class Rec> MDB_HD int voronoi_atom(const VoroArgs<Rec> &A, int s)
{
    const Rec me = A.sorted[s];
    const DBox &box = A.box;
    double pc[3] = {me.x, me.y, me.z};   // Cartesian, relative to the box origin
    if (!A.wrapped && box.any_pbc) wrap_into_box(box, pc[0], pc[1], pc[2]);
    for (int d = 0; d < 3; ++d) pc[d] -= box.origin[d];
    // p: the coordinate the grid slabs are measured in -- Cartesian for an orthogonal box, fractional coordinate x
    // perpendicular thickness for a triclinic one (cell_of(); a slab of cells is A.w thick along its own normal)
    double p[3] = {pc[0], pc[1], pc[2]};
    if (box.triclinic)
        for (int d = 0; d < 3; ++d) p[d] = (pc[0] * box.hinv[d] + pc[1] * box.hinv[3 + d] + pc[2] * box.hinv[6 + d]) * A.L[d];
    const int i = me.idx;
    bool inside = true;
    for (int d = 0; d < 3; ++d) inside = inside && (box.pbc[d] || (p[d] >= 0.0 && p[d] <= A.L[d]));
    if (!inside) {   // voro++ does not store a particle outside an open container: its outputs keep their zeros
        A.volume[i] = 0.0;
        A.nfaces[i] = 0;
        A.radius[i] = 0.0;
        if (A.W)
            for (int k = 0; k < A.W; ++k) A.row_id[(size_t)i * A.W + k] = -1, A.row_area[(size_t)i * A.W + k] = 0.0;
        return 0;
    }
    Cell C;
    int c[3];
    cell_decode(A.g, me.cell, c[0], c[1], c[2]);
    const int *n = A.g.n;
    // Fast mode first; a cell that met an ill-conditioned vertex solve (three planes almost sharing a line: lattices
    // with ~1e-8 noise) is rebuilt once in careful mode, where such vertices are interpolated along the cut edge.
    for (int careful = 0; careful < 2; ++careful) {
    {
        double lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            lo[d] = box.triclinic ? -A.R0 : (box.pbc[d] ? -0.5 * A.L[d] : -p[d]);
            hi[d] = box.triclinic ? A.R0 : (box.pbc[d] ? 0.5 * A.L[d] : A.L[d] - p[d]);
        }
        C.careful = careful;
        C.init(lo, hi);
    }
    // Schedule: the atom's own cell and the first shell are walked twice, first for the candidates closer than
    // 0.8 cell widths (about the first neighbour shell of a crystal at this density), then for the rest -- planes of
    // near atoms shrink the cell before the far ones are tested, so fewer planes are inserted only to be cut away.
    const double near2 = 0.64 * A.w * A.w;
    for (int it = 0;; ++it) {
        const int sh = it < 4 ? (it & 1) : it - 2;
        const int part = it < 2 ? 1 : (it < 4 ? 2 : 0);   // 1: near candidates only, 2: far only, 0: all
        if (sh > 0) {
            // distance from the atom to the boundary of the block of shells < sh: nothing beyond it can cut
            double dmin = 1.0e300;
            bool any_side = false;
            for (int d = 0; d < 3; ++d) {
                double lo, hi, t;
                if (box.pbc[d] || c[d] - sh >= 0) {
                    cell_span(c[d] - sh + 1, n[d], A.w, A.L[d], lo, t);
                    dmin = fmin(dmin, p[d] - lo);
                    any_side = true;
                }
                if (box.pbc[d] || c[d] + sh <= n[d] - 1) {
                    cell_span(c[d] + sh - 1, n[d], A.w, A.L[d], t, hi);
                    dmin = fmin(dmin, hi - p[d]);
                    any_side = true;
                }
            }
            if (!any_side) break;
            if (dmin > 0.0 && dmin * dmin >= 4.0 * C.rmax2) break;
        }
        // shell cells in three passes: face-, edge-, corner-adjacent offsets (nearest cells cut first)
        for (int pass = (sh == 0 ? 0 : 1); pass <= (sh == 0 ? 0 : 3); ++pass)
            for (int di = -sh; di <= sh; ++di) {
                const AxisCell a0 = axis_cell(A, 0, c[0] + di, p[0]);
                if (!a0.valid) continue;
                for (int dj = -sh; dj <= sh; ++dj) {
                    const AxisCell a1 = axis_cell(A, 1, c[1] + dj, p[1]);
                    if (!a1.valid) continue;
                    for (int dk = -sh; dk <= sh; ++dk) {
                        const int ai = di < 0 ? -di : di, aj = dj < 0 ? -dj : dj, ak = dk < 0 ? -dk : dk;
                        if (sh > 0) {
                            if (ai != sh && aj != sh && ak != sh) continue;
                            if ((ai == sh) + (aj == sh) + (ak == sh) != pass) continue;
                        }
                        const AxisCell a2 = axis_cell(A, 2, c[2] + dk, p[2]);
                        if (!a2.valid) continue;
                        // orthogonal: the gaps are the components of the distance; triclinic: each is a distance
                        // along a slab normal, the largest one bounds the distance from below
                        const double gap2 = box.triclinic ? fmax(a0.gap2, fmax(a1.gap2, a2.gap2)) : a0.gap2 + a1.gap2 + a2.gap2;
                        if (gap2 >= 4.0 * C.rmax2) continue;
                        // candidate - atom = record + image shift - pc
                        const double ox = a0.s[0] + a1.s[0] + a2.s[0] - pc[0], oy = a0.s[1] + a1.s[1] + a2.s[1] - pc[1],
                                     oz = a0.s[2] + a1.s[2] + a2.s[2] - pc[2];
                        const int cell = cell_linear(A.g, a0.kk, a1.kk, a2.kk);
                        const int b0 = A.cell_start[cell], b1 = A.cell_start[cell + 1];
                        for (int q = b0; q < b1; ++q) {
                            if (q == s && di == 0 && dj == 0 && dk == 0) continue;
                            const Rec o = A.sorted[q];
                            double r[3] = {o.x, o.y, o.z};
                            if (!A.wrapped && box.any_pbc) wrap_into_box(box, r[0], r[1], r[2]);
                            for (int d = 0; d < 3; ++d) r[d] -= box.origin[d];
                            if (A.has_open) {   // an atom outside an open container is not stored by voro++
                                bool ok = true;
                                for (int d = 0; d < 3; ++d) ok = ok && (box.pbc[d] || (r[d] >= 0.0 && r[d] <= A.L[d]));
                                if (!ok) continue;
                            }
                            r[0] += ox, r[1] += oy, r[2] += oz;
#ifdef VORO_COUNT
                            ++g_cand;
#endif
                            const double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                            if (!(d2 > 0.0) || d2 >= 4.0 * C.rmax2) continue;
                            if ((part == 1 && d2 >= near2) || (part == 2 && d2 < near2)) continue;
                            C.clip(r[0], r[1], r[2], 0.5 * d2, o.idx, A.tolh);
                        }
                    }
                }
            }
        if (C.fail) break;
    }
    if (C.fail || !C.ill) break;
    }
    if (C.fail) {
        A.volume[i] = 0.0;
        A.nfaces[i] = -1;
        A.radius[i] = 0.0;
        return -1;
    }
    C.collect();
    double vol = 0.0;
    int nf = 0;
    for (int f = 0; f < C.np; ++f) {
        double perim;
        const double area = C.face_area(f, perim);
        const double nlen = sqrt(C.px[f] * C.px[f] + C.py[f] * C.py[f] + C.pz[f] * C.pz[f]);
        vol += area * C.pd[f] / nlen;
        // A plane inserted while the cell was still large can end up touching the final cell in a vertex or along
        // an edge only (second neighbours of a perfect or strained lattice): its polygon survives with a width of
        // rounding-error size.  voro++ meets such a plane last, finds the vertices within its tolerance of the plane
        // and cuts nothing, so the reference has no such face: drop polygons narrower than that tolerance.
        if (!(2.0 * area > perim * (4.0 * A.tolh / nlen))) continue;
        if (A.W && nf < A.W) {
            A.row_id[(size_t)i * A.W + nf] = C.pid[f];
            A.row_area[(size_t)i * A.W + nf] = area;
        }
        ++nf;
    }
    A.volume[i] = vol / 3.0;
    A.nfaces[i] = nf;
    A.radius[i] = 2.0 * sqrt(C.rmax2);
    return nf;
}
}  // namespace voro

