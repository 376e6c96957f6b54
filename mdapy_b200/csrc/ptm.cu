// mdapy_b200/csrc/ptm.cu -- polyhedral template matching on the device.
// Replaces _ptm.get_ptm (src/polyhedral_template_matching.cpp:135-319): the serial Voronoi pre-ordering
// stage (211-255) and the OpenMP ptm_index loop (258-318) become ONE kernel, one thread per atom, running
// the per-atom core of ptm_core.cuh.  Look-up tables are generated on the host at first use
// (ptm_tables.h) and kept in device memory.
#include "internal.cuh"
#include "ptm_tables.h"
#include <mutex>

constexpr int PTM_RESIDENT_DEFAULT = 512;   // threads per SM for k_ptm_match (see launch_ptm)

namespace {

struct DeviceTables {
    ptm::Tables *d_tables{nullptr};
    void *d_hash{nullptr}, *d_aut_begin{nullptr}, *d_aut_label{nullptr}, *d_gen{nullptr};
    bool ready{false};
};
DeviceTables g_dev[16];
std::mutex g_mu;

const ptm::Tables *device_tables(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    MDB_REQUIRE(device >= 0 && device < 16, MDB_ERR_VALUE, "device index %d out of range", device);
    DeviceTables &D = g_dev[device];
    if (D.ready) return D.d_tables;
    ptm::HostTables H;
    ptm::build_tables(H);
    ptm::Tables T = H.t;
    auto up = [](const void *src, size_t bytes, void **dst) {
        CUDA_TRY(cudaMalloc(dst, bytes ? bytes : 8));
        CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    };
    up(H.hash.data(), H.hash.size() * sizeof(unsigned long long), &D.d_hash);
    up(H.aut_begin.data(), H.aut_begin.size() * sizeof(int), &D.d_aut_begin);
    up(H.aut_label.data(), H.aut_label.size(), &D.d_aut_label);
    up(H.gen.data(), H.gen.size() * sizeof(double), &D.d_gen);
    T.hash = static_cast<const unsigned long long *>(D.d_hash);
    T.aut_begin = static_cast<const int *>(D.d_aut_begin);
    T.aut_label = static_cast<const signed char *>(D.d_aut_label);
    T.gen = static_cast<const double *>(D.d_gen);
    void *dt = nullptr;
    up(&T, sizeof(T), &dt);
    D.d_tables = static_cast<ptm::Tables *>(dt);
    D.ready = true;
    return D.d_tables;
}

// neighbour vectors of atom i in list order: polyhedral_template_matching.cpp:222-248
__device__ __forceinline__ int gather_points(const double *__restrict__ x, const double *__restrict__ y,
                                             const double *__restrict__ z, int N, const DBox &box,
                                             const int *__restrict__ row, int M, int i, double (*pts)[3], int *nbr)
{
    const double xi = x[i], yi = y[i], zi = z[i];
    int num = 0;
    for (int k = 0; k < M && num < ptm::MAX_IN; ++k) {
        const int j = row[k];
        if (j < 0 || j >= N) break;
        if (j == i) continue;
        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
        min_image(box, dx, dy, dz);
        pts[num][0] = dx;
        pts[num][1] = dy;
        pts[num][2] = dz;
        nbr[num] = j;
        ++num;
    }
    return num;
}

// Pass 1 (the reference's pre-ordering loop, polyhedral_template_matching.cpp:213-252): Voronoi solid-angle
// ranking of every atom's listed neighbours.  order[i][r] = list position of the rank-r neighbour.
__global__ void __launch_bounds__(64) k_ptm_order(const double *__restrict__ x, const double *__restrict__ y,
                                                  const double *__restrict__ z, int N, int n_rows,
                                                  const __grid_constant__ DBox box, const int *__restrict__ verlet, int M,
                                                  unsigned char *__restrict__ order_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double pts[ptm::MAX_IN][3];
    int nbr[ptm::MAX_IN];
    const int num = gather_points(x, y, z, N, box, verlet + (size_t)i * M, M, i, pts, nbr);
    int order[ptm::MAX_IN];
    // Voronoi face polygons of this thread: shared memory, one column per thread (conflict-free)
    __shared__ double poly[4 * ptm::MAX_POLY2 * 64];
    ptm::preorder_neighbours<64>(num, pts, order, poly + threadIdx.x);
    unsigned char *o = order_out + (size_t)i * ptm::MAX_IN;
    for (int k = 0; k < ptm::MAX_IN; ++k) o[k] = (unsigned char)(k < num ? order[k] : 255);
}

// access to other atoms' lists and rankings for the two-shell structures (ptm::two_shell_env)
struct DeviceSrc {
    const double *x, *y, *z;
    int N;
    const DBox *box;
    const int *verlet;
    int M;
    const int *types;
    const unsigned char *order;
    __device__ int gather(int i, double (*pts)[3], int *nbr) const
    {
        return gather_points(x, y, z, N, *box, verlet + (size_t)i * M, M, i, pts, nbr);
    }
    __device__ const unsigned char *order_of(int i) const { return order + (size_t)i * ptm::MAX_IN; }
    __device__ int type_of(int i) const { return types ? types[i] : 0; }
};

// Pass 2 (polyhedral_template_matching.cpp:255-316): template matching on the ranked neighbours.
template <int MINB>
__global__ void __launch_bounds__(64, MINB) k_ptm_match(const double *__restrict__ x, const double *__restrict__ y,
                                                  const double *__restrict__ z, int N, int n_rows,
                                                  const __grid_constant__ DBox box, const int *__restrict__ verlet, int M,
                                                  const unsigned char *__restrict__ order_in,
                                                  const int *__restrict__ types, int flags, double rmsd_threshold,
                                                  const ptm::Tables *__restrict__ tables, double *__restrict__ output,
                                                  int ocols, int *__restrict__ indices, int icols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    double pts[ptm::MAX_IN][3];
    int nbr[ptm::MAX_IN], ty[ptm::MAX_IN + 1], order[ptm::MAX_IN];
    const int num = gather_points(x, y, z, N, box, verlet + (size_t)i * M, M, i, pts, nbr);
    ty[0] = types ? types[i] : 0;
    for (int k = 0; k < num; ++k) {
        ty[1 + k] = types ? types[nbr[k]] : 0;
        order[k] = order_in[(size_t)i * ptm::MAX_IN + k];
    }
    const DeviceSrc src{x, y, z, N, &box, verlet, M, types, order_in};
    ptm::Result r;
    ptm::match_atom(*tables, flags, num, pts, order, ty, nbr, src, i, r);
    // outputs: polyhedral_template_matching.cpp:265-314
    int type = r.type, ordering = r.ordering;
    if (r.rmsd > rmsd_threshold || type == 0) {
        type = 0;
        ordering = 0;
    }
    double vals[8] = {(double)type, (double)ordering, r.rmsd, r.interatomic_distance, r.q[0], r.q[1], r.q[2], r.q[3]};
    double *o = output + (size_t)i * ocols;
    for (int c = 0; c < ocols; ++c) o[c] = c < 8 ? vals[c] : 0.0;
    if (indices) {
        int *ind = indices + (size_t)i * icols;
        const int n = r.struct_index >= 0 ? tables->n_nbrs[r.struct_index] : -1;
        for (int c = 0; c < icols; ++c) ind[c] = c <= n ? r.env_idx[r.mapping[c]] : -1;
    }
}

}  // namespace

// structure string -> flag set, same grammar as polyhedral_template_matching.cpp:168-206
int ptm_parse_flags(const char *structure)
{
    static const char *names[] = {"fcc", "hcp", "bcc", "ico", "sc", "dcub", "dhex", "graphene", "all", "default"};
    static const int flags[] = {ptm::CHECK_FCC, ptm::CHECK_HCP, ptm::CHECK_BCC, ptm::CHECK_ICO, ptm::CHECK_SC,
                                ptm::CHECK_DCUB, ptm::CHECK_DHEX, ptm::CHECK_GRAPHENE, 255,
                                ptm::CHECK_FCC | ptm::CHECK_HCP | ptm::CHECK_BCC | ptm::CHECK_ICO};
    auto sep = [](char c) { return c == '\0' || c == ' ' || c == ',' || c == '-' || c == '_' || c == '|'; };
    int out = 0;
    const char *p = structure ? structure : "";
    while (*p) {
        if (sep(*p)) {
            ++p;
            continue;
        }
        bool found = false;
        for (int k = 0; k < 10; ++k) {
            const size_t len = strlen(names[k]);
            if (strncmp(p, names[k], len) == 0 && sep(p[len])) {
                out |= flags[k];
                p += len;
                found = true;
                break;
            }
        }
        if (!found) ++p;
    }
    if (out == 0) out = ptm::CHECK_FCC | ptm::CHECK_HCP | ptm::CHECK_BCC | ptm::CHECK_ICO;
    return out;
}

void launch_ptm(MdbSystem &s, int flags, const int *verlet, int M, const int *types, double rmsd_threshold,
                double *output, int ocols, int *indices, int icols)
{
    MDB_REQUIRE(!(flags & (ptm::CHECK_DCUB | ptm::CHECK_DHEX | ptm::CHECK_GRAPHENE)) || s.slab_nx == 0, MDB_ERR_STATE,
                "PTM structures dcub / dhex / graphene read neighbours of neighbours: not available on a decomposed frame");
    const ptm::Tables *T = device_tables(s.device);
    const int R = s.n_rows;
    unsigned char *order = s.scratch.ensure<unsigned char>((size_t)R * ptm::MAX_IN);
    MDB_LAUNCH(k_ptm_order, (R + 63) / 64, 64, 0, s.stream, s.x, s.y, s.z, s.N, R, s.box, verlet, M, order);
    // The matching kernel keeps ~5 KB of per-thread scratch in local memory.  With all 512 resident threads per SM
    // that is 390 MB chip-wide -- three times the L2 -- and every scratch line is written back to DRAM (7 KB of
    // traffic per atom).  Reserving shared memory caps the residency so that the scratch of all resident threads
    // stays inside the L2 (MDB_PTM_RESIDENT threads per SM; measured optimum in profiles/r2_ptm_residency.txt).
    static const int resident = getenv("MDB_PTM_RESIDENT") ? atoi(getenv("MDB_PTM_RESIDENT")) : PTM_RESIDENT_DEFAULT;
    size_t pad = 0;
    if (resident > 0 && resident < 512) {
        const int blocks = resident / 64 > 0 ? resident / 64 : 1;
        pad = (size_t)(220 * 1024) / blocks - 1024;
        // per launch: the attribute belongs to the current device (several devices per process: group.cu)
        CUDA_TRY(cudaFuncSetAttribute(k_ptm_match<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    }
    // Register budget: 16 CTAs/SM (64 registers, 1024 resident threads) beats 8 CTAs/SM (128 registers, 512 threads)
    // by 8 % (4.19 M rattled BCC: 166 vs 179 ms) -- the kernel is latency bound, residency wins over spills.  The
    // 128-register build is kept for the MDB_PTM_RESIDENT / MDB_PTM_MINB=8 experiments.
    int minb = 16;
    if (const char *e = getenv("MDB_PTM_MINB")) minb = atoi(e);
    if (minb == 8 || pad)
        MDB_LAUNCH(k_ptm_match<8>, (R + 63) / 64, 64, pad, s.stream, s.x, s.y, s.z, s.N, R, s.box, verlet, M, order, types,
                   flags & 255, rmsd_threshold, T, output, ocols, indices, icols);
    else
        MDB_LAUNCH(k_ptm_match<16>, (R + 63) / 64, 64, 0, s.stream, s.x, s.y, s.z, s.N, R, s.box, verlet, M, order, types,
                   flags & 255, rmsd_threshold, T, output, ocols, indices, icols);
    CUDA_TRY(cudaGetLastError());
}
