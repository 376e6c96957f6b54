// tests/host/ptm_host_harness.cpp -- TEST-ONLY host build of the PTM core (ptm_core.cuh is
// __host__ __device__).  Lets the CPU test-suite check the per-atom arithmetic, the generated
// template tables and the symmetry groups against the reference without a GPU.  The product never
// uses this: mdapy_b200 runs the same functions only inside the CUDA kernel of ptm.cu.
#include "../../mdapy_b200/csrc/ptm_tables.h"
#include <cstring>

static ptm::HostTables g_tables;
static bool g_ready = false;

static void ensure()
{
    if (!g_ready) {
        ptm::build_tables(g_tables);
        g_ready = true;
    }
}

extern "C" {

// counts[0..4]: template triangulation classes of SC, FCC, HCP, ICO, BCC; counts[5..7]: group orders;
// counts[8]: total automorphism labellings
void ptmh_tables_info(int *counts)
{
    ensure();
    for (int s = 0; s < ptm::NSTRUCT; ++s) counts[s] = g_tables.t.graph_begin[s + 1] - g_tables.t.graph_begin[s];
    for (int g = 0; g < 3; ++g) counts[5 + g] = g_tables.t.gen_begin[g + 1] - g_tables.t.gen_begin[g];
    counts[8] = (int)(g_tables.aut_label.size() / ptm::MAX_NB);
}

void ptmh_template(int s, double *out)  // (MAX_NB+1) x 3
{
    ensure();
    memcpy(out, g_tables.t.tpl[s], sizeof(double) * 3 * (ptm::MAX_NB + 1));
}

void ptmh_index(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
                const int *boundary3, const int *verlet, int M, const int *types, int flags, double rmsd_threshold,
                double *output, int *indices)
{
    ensure();
    DBox box;
    dbox_make(box, box9, origin3, boundary3);
    for (int i = 0; i < N; ++i) {
        double pts[ptm::MAX_IN][3];
        int nbr[ptm::MAX_IN], ty[ptm::MAX_IN + 1];
        int num = 0;
        ty[0] = types ? types[i] : 0;
        for (int k = 0; k < M && num < ptm::MAX_IN; ++k) {
            const int j = verlet[(size_t)i * M + k];
            if (j < 0 || j >= N) break;
            if (j == i) continue;
            double dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
            min_image(box, dx, dy, dz);
            pts[num][0] = dx;
            pts[num][1] = dy;
            pts[num][2] = dz;
            nbr[num] = j;
            ty[1 + num] = types ? types[j] : 0;
            ++num;
        }
        ptm::Result r;
        int order[ptm::MAX_IN];
        ptm::index_atom(g_tables.t, flags, num, pts, ty, r, order);
        double *o = output + (size_t)i * 8;
        int *ind = indices + (size_t)i * 18;
        for (int k = 0; k < 18; ++k) ind[k] = -1;
        if (r.struct_index >= 0) {
            const int n = g_tables.t.n_nbrs[r.struct_index];
            ind[0] = i;
            for (int p = 1; p <= n; ++p) ind[p] = nbr[order[r.mapping[p] - 1]];
        }
        int type = r.type, ordering = r.ordering;
        if (r.rmsd > rmsd_threshold || type == 0) {
            type = 0;
            ordering = 0;
        }
        o[0] = type;
        o[1] = ordering;
        o[2] = r.rmsd;
        o[3] = r.interatomic_distance;
        o[4] = r.q[0];
        o[5] = r.q[1];
        o[6] = r.q[2];
        o[7] = r.q[3];
    }
}
}
