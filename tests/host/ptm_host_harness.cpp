// tests/host/ptm_host_harness.cpp -- TEST-ONLY host build of the PTM core (ptm_core.cuh is
// __host__ __device__).  Lets the CPU test-suite check the per-atom arithmetic, the generated
// template tables and the symmetry groups against the reference without a GPU.  The product never
// uses this: mdapy_b200 runs the same functions only inside the CUDA kernel of ptm.cu.
#include "../../mdapy_b200/csrc/ptm_tables.h"
#include <cstring>
#include <vector>

static ptm::HostTables g_tables;
static bool g_ready = false;

static void ensure()
{
    if (!g_ready) {
        ptm::build_tables(g_tables);
        g_ready = true;
    }
}

// access to other atoms' lists and rankings, as the kernel has it (ptm.cu DeviceSrc)
struct HostSrc {
    const double *x, *y, *z;
    int N;
    DBox box;
    const int *verlet;
    int M;
    const int *types;
    const unsigned char *order;  // [N][MAX_IN]

    int gather(int i, double (*pts)[3], int *nbr) const
    {
        int num = 0;
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
            ++num;
        }
        return num;
    }
    const unsigned char *order_of(int i) const { return order + (size_t)i * ptm::MAX_IN; }
    int type_of(int i) const { return types ? types[i] : 0; }
};

extern "C" {

// counts[0..7]: template triangulation classes of SC, FCC, HCP, ICO, BCC, DCUB, DHEX, graphene (none);
// counts[8..10]: group orders; counts[11]: total automorphism labellings
void ptmh_tables_info(int *counts)
{
    ensure();
    for (int s = 0; s < ptm::NSTRUCT; ++s) counts[s] = g_tables.t.graph_begin[s + 1] - g_tables.t.graph_begin[s];
    for (int g = 0; g < 3; ++g) counts[8 + g] = g_tables.t.gen_begin[g + 1] - g_tables.t.gen_begin[g];
    counts[11] = (int)(g_tables.aut_label.size() / ptm::MAX_NB);
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
    HostSrc src{x, y, z, N, {}, verlet, M, types, nullptr};
    dbox_make(src.box, box9, origin3, boundary3);
    // pass 1: ranking of every atom's listed neighbours
    std::vector<unsigned char> order((size_t)N * ptm::MAX_IN, 255);
    for (int i = 0; i < N; ++i) {
        double pts[ptm::MAX_IN][3], buf[4 * ptm::MAX_POLY2];
        int nbr[ptm::MAX_IN], ord[ptm::MAX_IN];
        const int num = src.gather(i, pts, nbr);
        ptm::preorder_neighbours<1>(num, pts, ord, buf);
        for (int k = 0; k < num; ++k) order[(size_t)i * ptm::MAX_IN + k] = (unsigned char)ord[k];
    }
    src.order = order.data();
    // pass 2: matching
    for (int i = 0; i < N; ++i) {
        double pts[ptm::MAX_IN][3];
        int nbr[ptm::MAX_IN], ty[ptm::MAX_IN + 1], ord[ptm::MAX_IN];
        const int num = src.gather(i, pts, nbr);
        ty[0] = src.type_of(i);
        for (int k = 0; k < num; ++k) {
            ty[1 + k] = src.type_of(nbr[k]);
            ord[k] = order[(size_t)i * ptm::MAX_IN + k];
        }
        ptm::Result r;
        ptm::match_atom(g_tables.t, flags, num, pts, ord, ty, nbr, src, i, r);
        double *o = output + (size_t)i * 8;
        int *ind = indices + (size_t)i * 18;
        for (int k = 0; k < 18; ++k) ind[k] = -1;
        if (r.struct_index >= 0) {
            const int n = g_tables.t.n_nbrs[r.struct_index];
            for (int p = 0; p <= n && p < 18; ++p) ind[p] = r.env_idx[r.mapping[p]];
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
