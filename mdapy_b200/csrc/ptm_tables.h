// mdapy_b200/csrc/ptm_tables.h -- host-side generation of everything PTM looks up.
//
// Nothing here is tabulated by hand: the ideal polyhedra are constructed from their definition
// (in the reference's frame conventions, extern/ptm/ptm_templates.h: cubic axes for SC / FCC / BCC,
// c || z with an in-plane neighbour on +x for HCP, a five-fold axis on z with an upper-ring vertex at
// azimuth 90 deg for ICO; barycentre 0, mean neighbour distance 1), every triangulation of their
// coplanar faces is enumerated and canonicalised with ptm::canonical routines, and the rotation
// groups are closed under multiplication from two generators.  The reference ships the same
// information as precomputed tables (ptm_graph_data.cpp, ptm_quat.cpp) -- 8 FCC, 16 HCP, 1 ICO, 1 SC
// and 218 BCC triangulation classes, which is also what this generator finds.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>
#include "ptm_core.cuh"

namespace ptm {

struct HostTables {
    Tables t{};                               // scalar part + (later) device pointers
    std::vector<unsigned long long> hash;     // per graph, sorted inside each structure
    std::vector<int> aut_begin;               // per graph + 1
    std::vector<signed char> aut_label;       // [n_aut][MAX_NB]
    std::vector<double> gen;                  // quaternions of the three rotation groups
};

namespace detail {

inline void add_point(std::vector<std::array<double, 3>> &v, double x, double y, double z) { v.push_back({x, y, z}); }

inline std::vector<std::array<double, 3>> template_points(int s)
{
    std::vector<std::array<double, 3>> p;
    const double PI = 3.14159265358979323846;
    add_point(p, 0, 0, 0);
    if (s == S_SC) {
        for (int d = 0; d < 3; ++d)
            for (int sg = -1; sg <= 1; sg += 2) {
                double v[3] = {0, 0, 0};
                v[d] = sg;
                add_point(p, v[0], v[1], v[2]);
            }
    } else if (s == S_FCC) {
        const double h = std::sqrt(2.0) / 2;
        for (int zero = 2; zero >= 0; --zero)  // z = 0 plane first
            for (int sa = 1; sa >= -1; sa -= 2)
                for (int sb = 1; sb >= -1; sb -= 2) {
                    double v[3];
                    v[zero] = 0;
                    v[(zero + 1) % 3] = sa * h;
                    v[(zero + 2) % 3] = sb * h;
                    add_point(p, v[0], v[1], v[2]);
                }
    } else if (s == S_HCP) {
        for (int k = 0; k < 6; ++k) add_point(p, std::cos(k * PI / 3), std::sin(k * PI / 3), 0);
        const double r = 1 / std::sqrt(3.0), h = std::sqrt(2.0 / 3.0);
        for (int sg = -1; sg <= 1; sg += 2)
            for (int k = 0; k < 3; ++k) {
                const double th = (30 + 120 * k) * PI / 180;
                add_point(p, r * std::cos(th), r * std::sin(th), sg * h);
            }
    } else if (s == S_ICO) {
        add_point(p, 0, 0, 1);
        add_point(p, 0, 0, -1);
        const double zr = 1 / std::sqrt(5.0), rr = 2 / std::sqrt(5.0);
        for (int k = 0; k < 5; ++k) {
            const double th = (90 + 72 * k) * PI / 180;
            add_point(p, rr * std::cos(th), rr * std::sin(th), zr);
        }
        for (int k = 0; k < 5; ++k) {
            const double th = (126 + 72 * k) * PI / 180;
            add_point(p, rr * std::cos(th), rr * std::sin(th), -zr);
        }
    } else if (s == S_BCC) {
        const double s1 = 14.0 / (8 * std::sqrt(3.0) + 12);
        for (int a = 1; a >= -1; a -= 2)
            for (int b = 1; b >= -1; b -= 2)
                for (int c = 1; c >= -1; c -= 2) add_point(p, a * s1, b * s1, c * s1);
        for (int d = 0; d < 3; ++d)
            for (int sg = 1; sg >= -1; sg -= 2) {
                double v[3] = {0, 0, 0};
                v[d] = 2 * s1 * sg;
                add_point(p, v[0], v[1], v[2]);
            }
    } else if (s == S_DCUB || s == S_DHEX || s == S_GRAPHENE) {
        // bond vectors of the centre (unit length), then the first shell, then for every first neighbour its own
        // other neighbours: diamond cubic and graphene are staggered (inner - b_m), hexagonal diamond is
        // eclipsed across the bond along c (inner + mirror_z(b_m)) and staggered across the other three
        std::vector<std::array<double, 3>> b;
        if (s == S_DCUB) {
            const double h = 1 / std::sqrt(3.0);
            b = {{h, h, h}, {h, -h, -h}, {-h, -h, h}, {-h, h, -h}};
        } else if (s == S_DHEX) {
            const double r = std::sqrt(2.0 / 3.0), q = std::sqrt(2.0) / 3;
            b = {{-r, q, -1.0 / 3}, {0, -2 * q, -1.0 / 3}, {r, q, -1.0 / 3}, {0, 0, 1}};
        } else {
            const double h = std::sqrt(3.0) / 2;
            b = {{0, 1, 0}, {h, -0.5, 0}, {-h, -0.5, 0}};
        }
        const int ni = (int)b.size();
        for (auto &v : b) add_point(p, v[0], v[1], v[2]);
        for (int i = 0; i < ni; ++i)
            for (int m = 0; m < ni; ++m) {
                if (m == i) continue;
                if (s == S_DHEX && i == 3) add_point(p, b[i][0] + b[m][0], b[i][1] + b[m][1], b[i][2] - b[m][2]);
                else add_point(p, b[i][0] - b[m][0], b[i][1] - b[m][1], b[i][2] - b[m][2]);
            }
        double mean = 0;
        for (size_t i = 1; i < p.size(); ++i) mean += std::sqrt(p[i][0] * p[i][0] + p[i][1] * p[i][1] + p[i][2] * p[i][2]);
        mean /= (double)(p.size() - 1);
        for (auto &q : p)
            for (double &c : q) c /= mean;
    }
    // snap tiny trigonometric residue so coplanarity tests are clean
    for (auto &q : p)
        for (double &c : q)
            if (std::fabs(c) < 1e-15) c = 0;
    return p;
}

// faces of the convex hull of the neighbour points (indices 0..n-1), each as a counter-clockwise polygon
inline std::vector<std::vector<int>> polyhedron_faces(const std::vector<std::array<double, 3>> &pts)
{
    const int n = (int)pts.size() - 1;
    auto P = [&](int i) { return pts[i + 1].data(); };
    std::vector<std::vector<int>> faces;
    std::vector<std::array<double, 3>> normals;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            for (int k = j + 1; k < n; ++k) {
                double e1[3], e2[3], nr[3];
                for (int d = 0; d < 3; ++d) {
                    e1[d] = P(j)[d] - P(i)[d];
                    e2[d] = P(k)[d] - P(i)[d];
                }
                cross3(e1, e2, nr);
                const double nn = std::sqrt(dot3(nr, nr));
                if (nn < 1e-9) continue;
                for (double &c : nr) c /= nn;
                const double off = dot3(nr, P(i));
                int pos = 0, neg = 0;
                for (int m = 0; m < n; ++m) {
                    const double dd = dot3(nr, P(m)) - off;
                    if (dd > 1e-9) ++pos;
                    if (dd < -1e-9) ++neg;
                }
                if (pos && neg) continue;  // not a supporting plane
                if (pos) {
                    for (double &c : nr) c = -c;
                }
                bool known = false;
                for (auto &q : normals)
                    if (std::fabs(q[0] - nr[0]) + std::fabs(q[1] - nr[1]) + std::fabs(q[2] - nr[2]) < 1e-7) known = true;
                if (known) continue;
                normals.push_back({nr[0], nr[1], nr[2]});
                const double off2 = dot3(nr, P(i));
                std::vector<int> on;
                double cen[3] = {0, 0, 0};
                for (int m = 0; m < n; ++m)
                    if (std::fabs(dot3(nr, P(m)) - off2) < 1e-9) {
                        on.push_back(m);
                        for (int d = 0; d < 3; ++d) cen[d] += P(m)[d];
                    }
                for (double &c : cen) c /= on.size();
                // angular order around the outward normal
                double ux[3], uy[3];
                for (int d = 0; d < 3; ++d) ux[d] = P(on[0])[d] - cen[d];
                cross3(nr, ux, uy);
                std::sort(on.begin(), on.end(), [&](int a, int b) {
                    double va[3], vb[3];
                    for (int d = 0; d < 3; ++d) {
                        va[d] = P(a)[d] - cen[d];
                        vb[d] = P(b)[d] - cen[d];
                    }
                    return std::atan2(dot3(va, uy), dot3(va, ux)) < std::atan2(dot3(vb, uy), dot3(vb, ux));
                });
                faces.push_back(on);
            }
    return faces;
}

inline void close_group(std::vector<std::array<double, 4>> &g)
{
    auto same = [](const std::array<double, 4> &a, const std::array<double, 4> &b) {
        double dp = 0, dm = 0;
        for (int d = 0; d < 4; ++d) {
            dp += std::fabs(a[d] - b[d]);
            dm += std::fabs(a[d] + b[d]);
        }
        return dp < 1e-9 || dm < 1e-9;
    };
    bool grew = true;
    while (grew) {
        grew = false;
        const size_t n = g.size();
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) {
                std::array<double, 4> r;
                quat_mul(g[i].data(), g[j].data(), r.data());
                bool known = false;
                for (auto &q : g)
                    if (same(q, r)) known = true;
                if (!known) {
                    if (r[0] < 0 || (r[0] == 0 && (r[1] < 0 || (r[1] == 0 && (r[2] < 0 || (r[2] == 0 && r[3] < 0))))))
                        for (double &c : r) c = -c;
                    for (double &c : r)
                        if (std::fabs(c) < 1e-15) c = 0;
                    g.push_back(r);
                    grew = true;
                }
            }
    }
}

}  // namespace detail

inline void build_tables(HostTables &H)
{
    Tables &T = H.t;
    const int nn[NSTRUCT] = {6, 12, 12, 12, 14, 16, 16, 9};
    const int nfac[NSTRUCT] = {8, 20, 20, 20, 24, 28, 28, 0};
    const int mdeg[NSTRUCT] = {4, 6, 6, 6, 8, 8, 8, 0};
    const int tid[NSTRUCT] = {5, 1, 2, 4, 3, 6, 7, 8};   // reference ids: SC 5, FCC 1, HCP 2, ICO 4, BCC 3, DCUB 6, DHEX 7, graphene 8
    const int grp[NSTRUCT] = {0, 0, 1, 2, 0, 0, 1, 1};   // cubic / hexagonal-conventional / icosahedral fundamental zones
    const int ninner[NSTRUCT] = {0, 0, 0, 0, 0, 4, 4, 3};
    H.aut_begin.clear();
    H.hash.clear();
    H.aut_label.clear();
    // rotation groups (unit quaternions, w first), closed from two generators each
    const double PI = 3.14159265358979323846;
    const double hq = std::sqrt(0.5);
    std::vector<std::array<double, 4>> groups[3];
    groups[0] = {{1, 0, 0, 0}, {hq, hq, 0, 0}, {hq, 0, hq, 0}};                     // cubic: 90 deg about x and y
    groups[1] = {{1, 0, 0, 0}, {std::sqrt(3.0) / 2, 0, 0, 0.5}, {0, 1, 0, 0}};       // 60 deg about z, 180 deg about x
    {   // icosahedral: five-fold about z, two-fold about the midpoint of the edge top vertex -- ring vertex at 90 deg
        double ax[3] = {0, 2 / std::sqrt(5.0), 1 + 1 / std::sqrt(5.0)};
        const double an = std::sqrt(dot3(ax, ax));
        groups[2] = {{1, 0, 0, 0}, {std::cos(PI / 5), 0, 0, std::sin(PI / 5)}, {0, ax[0] / an, ax[1] / an, ax[2] / an}};
    }
    for (auto &g : groups) detail::close_group(g);
    for (int s = 0; s < NSTRUCT; ++s) {
        T.n_nbrs[s] = nn[s];
        T.n_facets[s] = nfac[s];
        T.max_degree[s] = mdeg[s];
        T.type_id[s] = tid[s];
        T.group[s] = grp[s];
        T.n_inner[s] = ninner[s];
        auto pts = detail::template_points(s);
        for (int i = 0; i <= MAX_NB; ++i)
            for (int d = 0; d < 3; ++d) T.tpl[s][i][d] = i < (int)pts.size() ? pts[i][d] : 0.0;
        T.c_dist[s] = std::sqrt(dot3(T.tpl[s][1], T.tpl[s][1]));
        const int n = nn[s];
        if (s == S_GRAPHENE) {   // matched without a graph (ptm_structure_matcher.cpp:348)
            T.graph_begin[s] = (int)H.hash.size();
            continue;
        }
        const int n_col = (s == S_DCUB || s == S_DHEX) ? 4 : 0;
        // permutations of the neighbour points induced by the proper rotations that map the template onto itself
        std::vector<std::vector<int>> perms;
        for (auto &q : groups[grp[s]]) {
            double rot[9];
            quat_to_matrix(q.data(), rot);
            std::vector<int> pi(n, -1);
            bool ok = true;
            for (int a = 0; a < n && ok; ++a) {
                double v[3];
                for (int j = 0; j < 3; ++j) v[j] = rot[3 * j] * pts[a + 1][0] + rot[3 * j + 1] * pts[a + 1][1] + rot[3 * j + 2] * pts[a + 1][2];
                for (int b = 0; b < n; ++b)
                    if (std::fabs(v[0] - pts[b + 1][0]) + std::fabs(v[1] - pts[b + 1][1]) + std::fabs(v[2] - pts[b + 1][2]) < 1e-9) pi[a] = b;
                if (pi[a] < 0) ok = false;
            }
            if (ok) perms.push_back(pi);
        }
        // faces, then every triangulation of the quadrilateral ones
        auto faces = detail::polyhedron_faces(pts);
        std::vector<int> quads;
        for (size_t f = 0; f < faces.size(); ++f)
            if (faces[f].size() == 4) quads.push_back((int)f);
        struct Entry {
            unsigned long long hash;
            std::vector<std::array<signed char, MAX_NB>> labels;
        };
        // one entry per GEOMETRIC class (orbit under the template's rotations); classes that happen to share a
        // canonical code stay separate entries with equal hash -- each contributes its own correspondences
        std::map<std::vector<std::array<int, 3>>, Entry> classes;
        auto facet_key = [](std::vector<std::array<int, 3>> fs) {
            for (auto &f : fs) {
                const int m = std::min(f[0], std::min(f[1], f[2]));
                while (f[0] != m) f = {f[1], f[2], f[0]};
            }
            std::sort(fs.begin(), fs.end());
            return fs;
        };
        for (unsigned mask = 0; mask < (1u << quads.size()); ++mask) {
            signed char facets[MAX_FACETS][3];
            int nf = 0;
            size_t qi = 0;
            for (size_t f = 0; f < faces.size(); ++f) {
                const auto &v = faces[f];
                if (v.size() == 3) {
                    facets[nf][0] = v[0], facets[nf][1] = v[1], facets[nf][2] = v[2];
                    ++nf;
                } else {
                    const bool alt = mask >> qi & 1u;
                    ++qi;
                    const int a = alt ? 1 : 0;
                    facets[nf][0] = v[a], facets[nf][1] = v[(a + 1) % 4], facets[nf][2] = v[(a + 2) % 4];
                    ++nf;
                    facets[nf][0] = v[a], facets[nf][1] = v[(a + 2) % 4], facets[nf][2] = v[(a + 3) % 4];
                    ++nf;
                }
            }
            if (n_col) {   // inner atoms back in as apexes (same surgery as match_diamond)
                for (int f = 0; f < nf; ++f) {
                    const int a = facets[f][0], b = facets[f][1], c = facets[f][2];
                    if (a < 4 || b < 4 || c < 4) continue;
                    const int i0 = (a - 4) / 3;
                    if ((b - 4) / 3 != i0 || (c - 4) / 3 != i0) continue;
                    facets[f][0] = (signed char)i0, facets[f][1] = (signed char)b, facets[f][2] = (signed char)c;
                    facets[nf][0] = (signed char)a, facets[nf][1] = (signed char)i0, facets[nf][2] = (signed char)c;
                    ++nf;
                    facets[nf][0] = (signed char)a, facets[nf][1] = (signed char)b, facets[nf][2] = (signed char)i0;
                    ++nf;
                }
            }
            // orbit representative: smallest facet list over the template's rotations
            std::vector<std::array<int, 3>> rep;
            for (auto &pi : perms) {
                std::vector<std::array<int, 3>> fs(nf);
                for (int f = 0; f < nf; ++f) fs[f] = {pi[facets[f][0]], pi[facets[f][1]], pi[facets[f][2]]};
                fs = facet_key(fs);
                if (rep.empty() || fs < rep) rep = fs;
            }
            if (classes.find(rep) != classes.end()) continue;
            Rotation R;
            build_rotation(n, nf, facets, R);
            // all start darts: minimal code and every labelling attaining it
            std::vector<signed char> best;
            std::vector<std::array<signed char, MAX_NB>> labs;
            int top = 0;
            for (int f = 0; f < nf; ++f)
                for (int e = 0; e < 3; ++e)
                    top = std::max(top, dart_key(R, facets[f][e], facets[f][(e + 1) % 3], facets[f][(e + 2) % 3]));
            for (int f = 0; f < nf; ++f)
                for (int e = 0; e < 3; ++e) {
                    if (dart_key(R, facets[f][e], facets[f][(e + 1) % 3], facets[f][(e + 2) % 3]) != top) continue;
                    signed char lab[MAX_NB], code[MAX_CODE];
                    int len;
                    if (dart_code(n, R, facets[f][e], facets[f][(e + 1) % 3], lab, code, len, nullptr, -1, n_col) != 1) continue;
                    std::vector<signed char> c(code, code + len);
                    std::array<signed char, MAX_NB> la{};
                    for (int u = 0; u < n; ++u) la[u] = lab[u];
                    if (best.empty() || c < best) {
                        best = c;
                        labs.clear();
                        labs.push_back(la);
                    } else if (c == best) {
                        if (std::find(labs.begin(), labs.end(), la) == labs.end()) labs.push_back(la);
                    }
                }
            // Two labellings that differ by a rigid rotation of the ideal template (t_sigma(u) = R t_u) give
            // correlation matrices R A and A: the same optimal-rotation eigenvalue, hence the same RMSD up to
            // rounding, whatever the environment.  The reference evaluates all of them (ptm_structure_matcher.cpp:
            // 57-105) and rounding noise picks the winner; here one labelling per class of rotation-equivalent
            // labellings is kept.  (All 24 labellings of the BCC hull are rotations of one another.)
            if (!getenv("MDB_PTM_ALL_AUTOMORPHISMS")) {
                std::vector<std::array<signed char, MAX_NB>> kept;
                for (const auto &la : labs) {
                    bool equivalent = false;
                    for (const auto &lb : kept) {
                        int sigma[MAX_NB];   // template neighbour u of `la` plays the role of sigma[u] in `lb`
                        for (int u = 0; u < n; ++u)
                            for (int v = 0; v < n; ++v)
                                if (lb[v] == la[u]) sigma[u] = v;
                        bool iso = true;
                        for (int u = 0; u < n && iso; ++u) {
                            const double *tu = T.tpl[s][1 + u], *su = T.tpl[s][1 + sigma[u]];
                            if (std::fabs(dot3(tu, tu) - dot3(su, su)) > 1e-9) iso = false;
                            for (int v = u + 1; v < n && iso; ++v) {
                                const double *tv = T.tpl[s][1 + v], *sv = T.tpl[s][1 + sigma[v]];
                                const double a[3] = {tu[0] - tv[0], tu[1] - tv[1], tu[2] - tv[2]};
                                const double b[3] = {su[0] - sv[0], su[1] - sv[1], su[2] - sv[2]};
                                if (std::fabs(dot3(a, a) - dot3(b, b)) > 1e-9) iso = false;
                            }
                        }
                        if (iso) {
                            equivalent = true;
                            break;
                        }
                    }
                    if (!equivalent) kept.push_back(la);
                }
                labs.swap(kept);
            }
            classes[rep] = Entry{code_hash(best.data(), (int)best.size()), labs};
        }
        std::vector<Entry> sorted;
        for (auto &kv : classes) sorted.push_back(kv.second);
        std::stable_sort(sorted.begin(), sorted.end(), [](const Entry &a, const Entry &b) { return a.hash < b.hash; });
        T.graph_begin[s] = (int)H.hash.size();
        for (auto &e : sorted) {
            H.hash.push_back(e.hash);
            H.aut_begin.push_back((int)(H.aut_label.size() / MAX_NB));
            for (auto &la : e.labels)
                for (int u = 0; u < MAX_NB; ++u) H.aut_label.push_back(la[u]);
        }
    }
    T.graph_begin[NSTRUCT] = (int)H.hash.size();
    H.aut_begin.push_back((int)(H.aut_label.size() / MAX_NB));
    H.gen.clear();
    T.gen_begin[0] = 0;
    for (int g = 0; g < 3; ++g) {
        for (auto &q : groups[g]) H.gen.insert(H.gen.end(), q.begin(), q.end());
        T.gen_begin[g + 1] = T.gen_begin[g] + (int)groups[g].size();
    }
    // {100} planes of the FCC template
    for (int a = 0; a < 3; ++a) {
        T.fcc_plane[a] = 0;
        for (int p = 1; p <= 12; ++p)
            if (T.tpl[S_FCC][p][a] == 0) T.fcc_plane[a] |= 1u << p;
    }
    // host pointers for CPU-side use (tests); ptm.cu replaces them by device copies
    T.hash = H.hash.data();
    T.aut_begin = H.aut_begin.data();
    T.aut_label = H.aut_label.data();
    T.gen = H.gen.data();
}

}  // namespace ptm
