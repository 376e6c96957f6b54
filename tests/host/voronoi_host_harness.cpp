// tests/host/voronoi_host_harness.cpp -- TEST-ONLY host build of the Voronoi cell core (voronoi_core.cuh is
// __host__ __device__).  Lets the CPU test-suite run the per-atom construction against the reference's fixtures
// without a GPU.  The product never uses this: mdapy_b200 runs the same functions only inside k_voronoi.
#include "../../mdapy_b200/csrc/voronoi_core.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

struct HostRec {
    double x, y, z;
    int idx, cell;
};

extern "C" int voronoi_host(const double *x, const double *y, const double *z, int N, const double *box9,
                            const double *origin3, const int *boundary3, double cell_scale, double *volume, int *nfaces,
                            double *radius, int *row_id, double *row_area, int W)
{
    DBox b;
    if (dbox_make(b, box9, origin3, boundary3)) return -1;
    if (b.triclinic && !(b.pbc[0] && b.pbc[1] && b.pbc[2])) return -2;
    const double vol = std::fabs(dbox_volume(b));
    const double w = 1.75 * std::cbrt(vol / N) * cell_scale;
    CellGrid g = cellgrid_make(b, w);
    std::vector<int> cell(N), start((size_t)g.total + 1, 0);
    for (int i = 0; i < N; ++i) {
        double xi = x[i], yi = y[i], zi = z[i];
        if (b.any_pbc) wrap_into_box(b, xi, yi, zi);
        int ic, jc, kc;
        cell_of(b, g, xi, yi, zi, ic, jc, kc);
        cell[i] = cell_linear(g, ic, jc, kc);
        ++start[cell[i] + 1];
    }
    for (int c = 0; c < g.total; ++c) start[c + 1] += start[c];
    std::vector<HostRec> sorted(N);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < N; ++i) sorted[fill[cell[i]]++] = HostRec{x[i], y[i], z[i], i, cell[i]};
    voro::VoroArgs<HostRec> A{};
    A.sorted = sorted.data();
    A.cell_start = start.data();
    A.N = N;
    A.box = b;
    A.g = g;
    A.w = w;
    double len2 = 0.0;
    A.R0 = 0.0;
    for (int d = 0; d < 3; ++d) {
        A.L[d] = b.triclinic ? b.thick[d] : b.h[d * 4];
        const double edge = std::sqrt(b.h[3 * d] * b.h[3 * d] + b.h[3 * d + 1] * b.h[3 * d + 1] + b.h[3 * d + 2] * b.h[3 * d + 2]);
        A.R0 += edge;
        len2 += edge * edge * (b.pbc[d] ? 0.25 : 1.0);
    }
    A.wrapped = 0;
    A.has_open = !(b.pbc[0] && b.pbc[1] && b.pbc[2]);
    A.tolh = 0.5 * 10.0 * 2.220446049250313e-16 * len2;
    A.volume = volume;
    A.nfaces = nfaces;
    A.radius = radius;
    A.row_id = row_id;
    A.row_area = row_area;
    A.W = W;
    int worst = 0;
    for (int s = 0; s < N; ++s) {
        const int nf = voro::voronoi_atom(A, s);
        if (nf < 0) return -3;
        worst = std::max(worst, nf);
    }
    return worst;
}
