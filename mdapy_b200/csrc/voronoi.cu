// mdapy_b200/csrc/voronoi.cu -- Voronoi cells of every atom.
//
// Reference: src/voronoi.cpp:16-71 (get_voronoi_volume_number_radius) and 307-447 (get_voronoi_neighbor) drive the
// vendored voro++ (extern/voro++: container_3d + voronoicell_neighbor_3d, a vertex/edge-table polyhedron cut plane by
// plane while a block worklist spirals outwards).  This is a different construction of the same cells, one GPU
// thread per atom:
//
//   * the cell is a list of half-spaces n.x <= d (the six walls of voro++'s initial box first: +-L/2 around the atom
//     on periodic axes, the container faces on open ones) and its vertices in DUAL form -- every vertex is the triple
//     of planes that meet there, the triples form an oriented triangulation of the "sphere of planes"
//     (Ray, Sokolov, Lefebvre, Levy, "Meshless Voronoi on the GPU", 2018);
//   * cutting with a bisector plane = find the vertices beyond it (the same 10 eps L^2 tolerance voro++ applies, so
//     a plane that merely touches a vertex of a perfect lattice cuts nothing and FCC keeps 12 faces), cancel the
//     interior edges of that triangle set, drop the set and fan the new plane over the boundary loop;
//   * candidates come from the cut-off cell grid of binning.cu, walked shell by shell around the atom's cell (every
//     (cell, periodic image) pair once, so a box smaller than the cell still meets its own images); a cell is skipped
//     when its nearest point is farther than 2 R_max (no bisector from there can reach the cell), a candidate when
//     |r| >= 2 R_max, and the walk ends at the first shell whose inner boundary is that far.
//
// Outputs: volume, face count and 2 R_max per atom (voro++ keeps vertices at double scale and the reference reports
// sqrt(max_radius_squared()) unscaled: voronoi.cpp:65), and per face the neighbour id (-1..-6 for walls) and area.
// Faces are listed in the order their planes were inserted (nearest cells first), not in voro++'s vertex-table order:
// the rows hold the same SET of (neighbour, area, distance) as the reference's.
#include "internal.cuh"

#include "voronoi_core.cuh"

namespace {
using voro::VB;
using voro::VP;
using voro::VT;
typedef voro::VoroArgs<SortedAtom> VoroArgs;

__global__ void __launch_bounds__(voro::VB, 16) k_voronoi(const VoroArgs A)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.N) return;
    const int nf = voro::voronoi_atom(A, s);
    if (nf < 0) atomicAdd(A.status + 1, 1);
    else atomicMax(A.status, nf);
}

// raw rows (W wide) -> the reference's arrays (M wide): voronoi.cpp:394-437
__global__ void __launch_bounds__(128) k_voronoi_rows(const double *__restrict__ x, const double *__restrict__ y,
                                                      const double *__restrict__ z, int N, DBox box,
                                                      const int *__restrict__ row_id, const double *__restrict__ row_area,
                                                      const int *__restrict__ nfaces, int W, int M, double a_thr,
                                                      double r_thr, int *__restrict__ verlet, double *__restrict__ dist,
                                                      double *__restrict__ area)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int nf = nfaces[i] < 0 ? 0 : nfaces[i];
    const int *ids = row_id + (size_t)i * W;
    const double *ar = row_area + (size_t)i * W;
    double area_min = 0.0;
    if (a_thr > 0) area_min = a_thr;
    if (r_thr > 0.0) {
        double sum = 0.0;
        for (int k = 0; k < nf; ++k) sum += ar[k];
        area_min = sum * r_thr;
    }
    if (a_thr > area_min) area_min = a_thr;
    const double xi = x[i], yi = y[i], zi = z[i];
    for (int k = 0; k < M; ++k) {
        const size_t e = (size_t)i * M + k;
        int j = -1;
        double a = 0.0, d = 10000.0;
        if (k < nf && ids[k] >= 0 && ar[k] > area_min) {
            j = ids[k];
            a = ar[k];
            double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
            min_image(box, dx, dy, dz);
            d = sqrt(dx * dx + dy * dy + dz * dz);
        }
        verlet[e] = j;
        area[e] = a;
        dist[e] = d;
    }
}
}  // namespace

// Cells of all atoms of the handle.  want_rows: also keep (neighbour id, face area) rows in s.vor_id / s.vor_area
// (width s.vor_W >= the largest face count, which is returned).
int launch_voronoi(MdbSystem &s, bool want_rows, double *volume, int *nfaces, double *radius)
{
    MDB_REQUIRE(!s.box.triclinic || (s.box.pbc[0] && s.box.pbc[1] && s.box.pbc[2]), MDB_ERR_VALUE,
                "Voronoi cells of a triclinic box need periodic boundaries on all three axes (the reference's "
                "container_triclinic is periodic; its Python side triples the open axes first)");
    MDB_REQUIRE(s.n_rows == s.N, MDB_ERR_STATE, "Voronoi cells need the whole frame on one device");
    const int N = s.N;
    const DBox &b = s.box;
    const double L[3] = {b.triclinic ? b.thick[0] : b.h[0], b.triclinic ? b.thick[1] : b.h[4],
                         b.triclinic ? b.thick[2] : b.h[8]};
    // cell width: ~2 R_max of a close-packed crystal at this density, so the first shell usually ends the walk
    const double vol = fabs(dbox_volume(b));
    double w = 1.75 * cbrt(vol / N);
    if (const char *e = getenv("MDB_VORONOI_CELL")) w *= atof(e);
    // the records carry WRAPPED coordinates (one wrap per atom here instead of one per candidate in the kernel)
    {
        double *wx = s.wx.ensure<double>(N), *wy = s.wy.ensure<double>(N), *wz = s.wz.ensure<double>(N);
        const size_t bytes = sizeof(double) * N;
        CUDA_TRY(cudaMemcpyAsync(wx, s.x, bytes, cudaMemcpyDeviceToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(wy, s.y, bytes, cudaMemcpyDeviceToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(wz, s.z, bytes, cudaMemcpyDeviceToDevice, s.stream));
        if (b.any_pbc) launch_wrap_positions(s, wx, wy, wz, N);
        const double *rx = s.x, *ry = s.y, *rz = s.z;
        s.x = wx, s.y = wy, s.z = wz;
        try {
            launch_binning(s, w);
        } catch (...) {
            s.x = rx, s.y = ry, s.z = rz;
            throw;
        }
        s.x = rx, s.y = ry, s.z = rz;
    }
    VoroArgs A{};
    A.sorted = s.sorted.as<SortedAtom>();
    A.cell_start = s.cell_start.as<int>();
    A.N = N;
    A.box = b;
    A.g = s.grid;
    A.w = w;
    A.wrapped = 1;
    A.has_open = !(b.pbc[0] && b.pbc[1] && b.pbc[2]);
    double len2 = 0.0;
    A.R0 = 0.0;
    for (int d = 0; d < 3; ++d) {
        A.L[d] = L[d];
        const double edge = sqrt(b.h[3 * d] * b.h[3 * d] + b.h[3 * d + 1] * b.h[3 * d + 1] + b.h[3 * d + 2] * b.h[3 * d + 2]);
        A.R0 += edge;
        len2 += edge * edge * (b.pbc[d] ? 0.25 : 1.0);   // container_3d: max_len_sq of the initial cell
    }
    A.tolh = 0.5 * 10.0 * 2.220446049250313e-16 * len2;
    A.volume = volume;
    A.nfaces = nfaces;
    A.radius = radius;
    int *status = s.counters.ensure<int>(8);
    A.status = status;
    int W = want_rows ? (s.vor_W > 0 ? s.vor_W : 32) : 0;
    int h[2] = {0, 0};
    for (int attempt = 0; attempt < 2; ++attempt) {
        A.W = W;
        if (W) {
            A.row_id = s.vor_id.ensure<int>((size_t)N * W);
            A.row_area = s.vor_area.ensure<double>((size_t)N * W);
        }
        CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * 2, s.stream));
        // MDB_VORONOI_RESIDENT caps the resident threads per SM with dynamic shared memory padding (experiment knob:
        // throughput grows with residency up to the register limit -- the kernel is latency bound, not capacity bound)
        int resident = 0;
        if (const char *e = getenv("MDB_VORONOI_RESIDENT")) resident = atoi(e);
        size_t pad = 0;
        if (resident > 0 && resident < 896) {
            const int blocks = resident / VB > 0 ? resident / VB : 1;
            pad = (size_t)(220 * 1024) / blocks - 1024;
            CUDA_TRY(cudaFuncSetAttribute(k_voronoi, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        }
        MDB_LAUNCH(k_voronoi, (N + VB - 1) / VB, VB, pad, s.stream, A);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h, status, sizeof(h), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        MDB_REQUIRE(h[1] == 0, MDB_ERR_VALUE, "%d Voronoi cells outgrew the per-thread buffers (%d planes, %d vertices)",
                    h[1], VP, VT);
        if (!W || h[0] <= W) break;
        W = (h[0] + 7) / 8 * 8;   // wider rows than expected: run again at the measured width
    }
    s.vor_W = W;
    s.bin_rc = -1.0;   // the grid belongs to this call: the next cut-off build bins again
    return h[0];
}

void launch_voronoi_rows(MdbSystem &s, const int *nfaces, int M, double a_thr, double r_thr, int *verlet, double *dist,
                         double *area)
{
    MDB_LAUNCH(k_voronoi_rows, (s.N + 127) / 128, 128, 0, s.stream, s.x, s.y, s.z, s.N, s.box, s.vor_id.as<int>(),
               s.vor_area.as<double>(), nfaces, s.vor_W, M, a_thr, r_thr, verlet, dist, area);
    CUDA_TRY(cudaGetLastError());
}
