// mdapy_b200/csrc/builders.cu
//
// Input side of the hot path on the device: benchmark-size frames are generated in HBM
// instead of being shipped over PCIe.
//   k_repeat_cell           src/repeat_cell.cpp:19-61   supercell replication (cell-major, iz fastest)
//   k_transform_flags/...   src/polycrystal.cpp:21-127  rotate a lattice block into a grain and keep the atoms
//                                                      inside the grain's convex cell (plane tests), order kept
//   k_filter_overlap        src/neighbor.cpp:390-487    drop the higher-index atom of every pair closer than rc
// Arithmetic follows the reference operation by operation (left-to-right sums, no FMA contraction), so the
// generated coordinates and the kept sets are bit-identical to the reference's.
#include "internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_repeat_cell(const double *__restrict__ old_pos, int n_old, DBox b, int nx, int ny,
                                                     int nz, double *__restrict__ ox, double *__restrict__ oy,
                                                     double *__restrict__ oz, double *__restrict__ o3, size_t total)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t cell = t / n_old;
    const int i = (int)(t - cell * n_old);
    const int ix = (int)(cell / ((size_t)ny * nz));
    const size_t rem = cell % ((size_t)ny * nz);
    const int iy = (int)(rem / nz), iz = (int)(rem % nz);
    const double sx = ix * b.h[0] + iy * b.h[3] + iz * b.h[6];
    const double sy = ix * b.h[1] + iy * b.h[4] + iz * b.h[7];
    const double sz = ix * b.h[2] + iy * b.h[5] + iz * b.h[8];
    const double X = old_pos[3 * i + 0] + sx, Y = old_pos[3 * i + 1] + sy, Z = old_pos[3 * i + 2] + sz;
    if (o3) {
        o3[3 * t + 0] = X;
        o3[3 * t + 1] = Y;
        o3[3 * t + 2] = Z;
    }
    if (ox) {
        ox[t] = X;
        oy[t] = Y;
        oz[t] = Z;
    }
}

struct GrainXf {
    double c[3], t[3], RT[9];   // RT[a*3+b] = R(b, a): pos_new = (pos - c) @ R^T + t
};

__device__ __forceinline__ void grain_transform(const GrainXf &g, double x, double y, double z, double &px, double &py,
                                                double &pz)
{
    const double dx = x - g.c[0], dy = y - g.c[1], dz = z - g.c[2];
    px = dx * g.RT[0] + dy * g.RT[3] + dz * g.RT[6] + g.t[0];
    py = dx * g.RT[1] + dy * g.RT[4] + dz * g.RT[7] + g.t[1];
    pz = dx * g.RT[2] + dy * g.RT[5] + dz * g.RT[8] + g.t[2];
}

__global__ void __launch_bounds__(256) k_transform_flags(const double *__restrict__ x, const double *__restrict__ y,
                                                         const double *__restrict__ z, int N, GrainXf g,
                                                         const double *__restrict__ planes, int nfaces,
                                                         int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double px, py, pz;
    grain_transform(g, x[i], y[i], z[i], px, py, pz);
    int inside = 1;
    for (int j = 0; j < nfaces; ++j) {
        const double val = px * planes[4 * j] + py * planes[4 * j + 1] + pz * planes[4 * j + 2] + planes[4 * j + 3];
        if (val >= 0.0) {
            inside = 0;
            break;
        }
    }
    flag[i] = inside;
}

__global__ void __launch_bounds__(256) k_transform_scatter(const double *__restrict__ x, const double *__restrict__ y,
                                                           const double *__restrict__ z, int N, GrainXf g,
                                                           const int *__restrict__ flag, const int *__restrict__ offs,
                                                           double *__restrict__ out3)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !flag[i]) return;
    double px, py, pz;
    grain_transform(g, x[i], y[i], z[i], px, py, pz);
    double *o = out3 + 3 * (size_t)offs[i];
    o[0] = px;
    o[1] = py;
    o[2] = pz;
}

// keep[j] = 0 iff some atom i < j lies within rc of j (distance evaluated like the reference's loop over i:
// xi wrapped, x[j] raw, minimum image).  One thread per atom j on the cell-sorted copy.
__global__ void __launch_bounds__(128) k_filter_overlap(const SortedAtom *__restrict__ sorted,
                                                        const int *__restrict__ cell_start, int N, DBox box, CellGrid g,
                                                        double rcsq, unsigned char *__restrict__ keep)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const SortedAtom me = sorted[s];
    int ic, jc, kc;
    cell_decode(g, me.cell, ic, jc, kc);
    unsigned char k = 1;
    for (int di = -1; di <= 1 && k; ++di)
        for (int dj = -1; dj <= 1 && k; ++dj)
            for (int dk = -1; dk <= 1 && k; ++dk) {
                const int c = cell_linear(g, wrap_cell(ic + di, g.n[0]), wrap_cell(jc + dj, g.n[1]), wrap_cell(kc + dk, g.n[2]));
                if (c < 0) continue;
                const int b = __ldg(cell_start + c), e = __ldg(cell_start + c + 1);
                for (int q = b; q < e; ++q) {
                    const SortedAtom o = sorted[q];
                    if (o.idx >= me.idx) continue;   // only a LOWER index removes me (neighbor.cpp:465)
                    double xi = o.x, yi = o.y, zi = o.z;
                    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
                    double dx = me.x - xi, dy = me.y - yi, dz = me.z - zi;
                    min_image(box, dx, dy, dz);
                    if (dx * dx + dy * dy + dz * dz <= rcsq) {
                        k = 0;
                        break;
                    }
                }
            }
    keep[me.idx] = k;
}

}  // namespace

void launch_repeat_cell(MdbSystem &s, const double *old_pos_dev, int n_old, const DBox &b, int nx, int ny, int nz,
                        double *ox, double *oy, double *oz, double *o3)
{
    const size_t total = (size_t)n_old * nx * ny * nz;
    if (!total) return;
    MDB_LAUNCH(k_repeat_cell, (unsigned)((total + 255) / 256), 256, 0, s.stream, old_pos_dev, n_old, b, nx, ny, nz, ox, oy,
               oz, o3, total);
    CUDA_TRY(cudaGetLastError());
}

// returns the number of atoms kept; out3 (device, >= 3 N doubles) receives them in input order
int launch_transform_and_filter(MdbSystem &s, const double *x, const double *y, const double *z, int N, const double *R9,
                                const double *center3, const double *target3, const double *planes_dev, int nfaces,
                                double *out3)
{
    if (N <= 0) return 0;
    GrainXf g;
    for (int a = 0; a < 3; ++a) {
        g.c[a] = center3[a];
        g.t[a] = target3[a];
        for (int b = 0; b < 3; ++b) g.RT[a * 3 + b] = R9[b * 3 + a];
    }
    int *flag = s.scratch.ensure<int>((size_t)N + 1);
    int *offs = s.scratch2.ensure<int>((size_t)N + 1);
    MDB_LAUNCH(k_transform_flags, (N + 255) / 256, 256, 0, s.stream, x, y, z, N, g, planes_dev, nfaces, flag);
    CUDA_TRY(cudaMemsetAsync(flag + N, 0, sizeof(int), s.stream));
    device_exclusive_scan(s, flag, offs, N + 1);
    int count = 0;
    CUDA_TRY(cudaMemcpyAsync(&count, offs + N, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    MDB_LAUNCH(k_transform_scatter, (N + 255) / 256, 256, 0, s.stream, x, y, z, N, g, flag, offs, out3);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return count;
}

void launch_filter_overlap(MdbSystem &s, double rc, unsigned char *keep)
{
    if (s.bin_rc != rc) launch_binning(s, rc);
    MDB_LAUNCH(k_filter_overlap, (s.N + 127) / 128, 128, 0, s.stream, s.sorted.as<SortedAtom>(), s.cell_start.as<int>(), s.N,
               s.box, s.grid, rc * rc, keep);
    CUDA_TRY(cudaGetLastError());
}
