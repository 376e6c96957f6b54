// mdapy_b200/csrc/group.cu -- one process, several GPUs: the neighbour search + CNA path of ONE unpartitioned
// host frame sharded over a list of devices (include/mdapy_b200.h, section "device group").
//
// The reference has no counterpart (src/neighbor.cpp and src/cna.cpp are OpenMP loops over one address
// space); this is the layer System(..., devices=[...]) sits on.  Data path, all of it library kernels and
// CUDA copies (no PyTorch, no NCCL, no host-side partitioning):
//
//   1. member d uploads the contiguous chunk [d*chunk, (d+1)*chunk) of x, y, z over its own PCIe link
//      (three copy streams per member, issued from host threads so pageable input stages in parallel);
//   2. k_route<false> counts, per destination slab, the atoms each chunk owns there and the atoms of the
//      slab's boundary planes its two neighbours need as ghosts (x cell planes of the GLOBAL cut-off grid,
//      src/neighbor.cpp:30-62 arithmetic); the D x D x 2 count matrix crosses the host once;
//   3. k_route<true> PUSHES every record (x, y, z raw, global id) straight into the destination member's slab
//      buffers with peer stores over NVLink -- each source owns a private range there, so there are no remote
//      atomics;
//   4. every member bins its slab and runs the fused neighbour + CNA kernel (neighbor_tiled.cu) on it;
//   5. k_route_labels pushes the labels back to the member that holds the atom's input chunk (peer stores),
//      and every member copies its chunk of labels to the caller's array over its own link.
//
// Frames the slab decomposition cannot take (fewer than 3 x-planes per member, or a frame the fused kernel
// declines) are gathered on member 0 with peer copies and take the ordinary single-GPU path there.
#include "internal.cuh"

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
constexpr int MAX_DEV = 16;

struct RouteDest {
    double *x, *y, *z;
    int *gid;
    int base[2];  // first row of this source's owned / ghost range in the destination slab
};
struct RouteTable {
    RouteDest d[MAX_DEV];
};
struct RoutePlan {
    int D;
    int n0;
    int bounds[MAX_DEV + 1];
};

// owner of plane p and the members that need an atom of that plane as a ghost (-1: none)
__device__ __forceinline__ void route_of(const RoutePlan &P, int p, int &own, int &gl, int &gr)
{
    own = 0;
#pragma unroll 1
    for (int r = 1; r < P.D; ++r) own += (p >= P.bounds[r]);
    gl = (p == P.bounds[own]) ? (own + P.D - 1) % P.D : -1;          // first plane of its slab: left neighbour's ghost
    gr = (p == P.bounds[own + 1] - 1) ? (own + 1) % P.D : -1;        // last plane: right neighbour's ghost
}

// WRITE = false: counts[2*d + kind] += atoms of this chunk bound for member d (kind 0 owned, 1 ghost)
// WRITE = true : the same walk, records stored at base + cursor (block-aggregated cursors, peer stores)
template <bool WRITE>
__global__ void __launch_bounds__(256) k_route(const double *__restrict__ x, const double *__restrict__ y,
                                               const double *__restrict__ z, int n, int gid0, DBox box, CellGrid g,
                                               const __grid_constant__ RoutePlan P, const __grid_constant__ RouteTable T,
                                               int *__restrict__ counts)
{
    __shared__ int s_cnt[2 * MAX_DEV];
    __shared__ int s_base[2 * MAX_DEV];
    if (threadIdx.x < 2 * MAX_DEV) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int own = -1, gl = -1, gr = -1, so = 0, sl = 0, sr = 0;
    double xr = 0, yr = 0, zr = 0;
    if (i < n) {
        xr = x[i], yr = y[i], zr = z[i];
        double xi = xr, yi = yr, zi = zr;
        if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
        int ic, jc, kc;
        cell_of(box, g, xi, yi, zi, ic, jc, kc);
        route_of(P, ic, own, gl, gr);
        so = atomicAdd(&s_cnt[2 * own], 1);
        if (gl >= 0) sl = atomicAdd(&s_cnt[2 * gl + 1], 1);
        if (gr >= 0) sr = atomicAdd(&s_cnt[2 * gr + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x < 2 * P.D) {
        const int c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd(&counts[threadIdx.x], c) : 0;
    }
    if (!WRITE) return;
    __syncthreads();
    if (i >= n) return;
    auto put = [&](int dest, int kind, int slot) {
        const RouteDest &R = T.d[dest];
        const int row = R.base[kind] + s_base[2 * dest + kind] + slot;
        R.x[row] = xr;
        R.y[row] = yr;
        R.z[row] = zr;
        R.gid[row] = gid0 + i;
    };
    put(own, 0, so);
    if (gl >= 0) put(gl, 1, sl);
    if (gr >= 0) put(gr, 1, sr);
}

struct LabelTable {
    int *lab[MAX_DEV];
};

// owned row i of this slab carries global id gid[i]; its label goes to the member holding input chunk gid / chunk
__global__ void __launch_bounds__(256) k_route_labels(const int *__restrict__ pat, const int *__restrict__ gid,
                                                      int n_owned, int chunk, const __grid_constant__ LabelTable T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    const int g = gid[i], src = g / chunk;
    T.lab[src][g - src * chunk] = pat[i];
}

struct Member {
    int device{0};
    mdb_system *sys{nullptr};
    cudaStream_t up[3]{};
    cudaEvent_t up_ev[3]{}, ready{};
    DevBuf cx, cy, cz;            // this member's input chunk
    int start{0}, count{0};
    DevBuf sx, sy, sz, sgid;      // its slab: [owned by source 0..D-1 | ghosts by source 0..D-1]
    DevBuf counts, lab;
    int h_counts[2 * MAX_DEV]{};
    int n_owned{0}, n_local{0};
    int used{0};
    int code{MDB_OK};
    std::string err;
};

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

struct mdb_group {
    int D{0};
    std::vector<Member> m;
    int N{0}, chunk{0};
    double box9[9]{}, origin[3]{};
    int boundary[3]{};
    bool has_atoms{false};
    int sharded{0};        // members the last frame ran on
    float times[6]{};      // upload issue, route, compute, labels, download, total (host clock, ms)
    double t_upload0{0};
};

namespace {
std::mutex g_err_mu;

// run fn(d) for every member on its own host thread (device selected), rethrow the first failure here
template <class F> void for_members(mdb_group &g, int count, F fn)
{
    std::vector<std::thread> th;
    th.reserve(count);
    for (int d = 0; d < count; ++d) {
        th.emplace_back([&g, d, &fn]() {
            Member &W = g.m[d % g.D];
            try {
                cudaError_t e = cudaSetDevice(W.device);
                if (e != cudaSuccess) {
                    mdb_set_error("cudaSetDevice(%d): %s", W.device, cudaGetErrorString(e));
                    throw MdbError{MDB_ERR_CUDA};
                }
                fn(d);
            } catch (const MdbError &e) {
                std::lock_guard<std::mutex> lk(g_err_mu);
                W.code = e.code;
                W.err = mdb_last_error();
            } catch (const std::exception &e) {
                std::lock_guard<std::mutex> lk(g_err_mu);
                W.code = MDB_ERR_CUDA;
                W.err = e.what();
            }
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < g.D; ++d) {
        Member &W = g.m[d];
        if (W.code != MDB_OK) {
            const int code = W.code;
            W.code = MDB_OK;
            mdb_set_error("device %d: %s", W.device, W.err.c_str());
            throw MdbError{code};
        }
    }
}

void check_call(int rc)
{
    if (rc != MDB_OK) throw MdbError{rc};   // the message is already in this thread's error slot
}

void wait_uploads(Member &W)
{
    for (int k = 0; k < 3; ++k) CUDA_TRY(cudaStreamWaitEvent(W.sys->stream, W.up_ev[k], 0));
}

// every chunk to member 0 (peer copies), then the ordinary single-GPU path there
void gathered_cna(mdb_group &g, double rc, int *pattern_host)
{
    Member &A = g.m[0];
    CUDA_TRY(cudaSetDevice(A.device));
    double *X = A.sx.ensure<double>(g.N), *Y = A.sy.ensure<double>(g.N), *Z = A.sz.ensure<double>(g.N);
    CUDA_TRY(cudaStreamSynchronize(A.sys->stream));   // recycled blocks: drained before a peer's stream writes them
    for (int d = 0; d < g.D; ++d) {
        Member &W = g.m[d];
        if (!W.count) continue;
        CUDA_TRY(cudaSetDevice(W.device));
        wait_uploads(W);
        const size_t bytes = sizeof(double) * W.count;
        // the copy runs on the SOURCE member's stream (its uploads are ordered there); member 0 waits for it below
        CUDA_TRY(cudaMemcpyPeerAsync(X + W.start, A.device, W.cx.as<double>(), W.device, bytes, W.sys->stream));
        CUDA_TRY(cudaMemcpyPeerAsync(Y + W.start, A.device, W.cy.as<double>(), W.device, bytes, W.sys->stream));
        CUDA_TRY(cudaMemcpyPeerAsync(Z + W.start, A.device, W.cz.as<double>(), W.device, bytes, W.sys->stream));
        CUDA_TRY(cudaEventRecord(W.ready, W.sys->stream));
    }
    CUDA_TRY(cudaSetDevice(A.device));
    for (int d = 0; d < g.D; ++d)
        if (g.m[d].count) CUDA_TRY(cudaStreamWaitEvent(A.sys->stream, g.m[d].ready, 0));
    check_call(mdb_system_set_atoms_device(A.sys, X, Y, Z, g.N, g.box9, g.origin, g.boundary));
    int used = 0;
    check_call(mdb_system_fused_cna(A.sys, rc, pattern_host, &used));
    if (!used) {
        int M = 0, mx = 0;
        check_call(mdb_system_build_neighbor(A.sys, rc, 0, &M, &mx));
        check_call(mdb_system_fcna(A.sys, rc, pattern_host));
    }
    CUDA_TRY(cudaStreamSynchronize(A.sys->stream));
    g.sharded = 1;
}
}  // namespace

extern "C" {

int mdb_group_create(const int *devices, int ndev, mdb_group **out)
{
    try {
        MDB_REQUIRE(out && devices, MDB_ERR_VALUE, "devices and out are required");
        MDB_REQUIRE(ndev >= 1 && ndev <= MAX_DEV, MDB_ERR_VALUE, "a device group holds 1..%d members, got %d", MAX_DEV,
                    ndev);
        mdb_group *g = new mdb_group();
        g->D = ndev;
        g->m.resize(ndev);
        *out = g;
        for (int d = 0; d < ndev; ++d) {
            Member &W = g->m[d];
            W.device = devices[d];
            check_call(mdb_system_create(devices[d], &W.sys));
            CUDA_TRY(cudaSetDevice(W.device));
            for (int k = 0; k < 3; ++k) {
                CUDA_TRY(cudaStreamCreateWithFlags(&W.up[k], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&W.up_ev[k], cudaEventDisableTiming));
            }
            CUDA_TRY(cudaEventCreateWithFlags(&W.ready, cudaEventDisableTiming));
            for (DevBuf *b : {&W.cx, &W.cy, &W.cz, &W.sx, &W.sy, &W.sz, &W.sgid, &W.counts, &W.lab})
                b->owner = &W.sys->stream;
        }
        // peer stores need every ordered pair of distinct devices mapped
        for (int a = 0; a < ndev; ++a)
            for (int b = 0; b < ndev; ++b) {
                const int da = devices[a], db = devices[b];
                if (da == db) continue;
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, da, db));
                MDB_REQUIRE(can, MDB_ERR_CUDA, "device %d cannot access device %d's memory (no peer path)", da, db);
                CUDA_TRY(cudaSetDevice(da));
                const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else CUDA_TRY(e);
            }
    } catch (const MdbError &e) {
        if (out && *out) {
            mdb_group_destroy(*out);
            *out = nullptr;
        }
        return e.code;
    }
    return MDB_OK;
}

void mdb_group_destroy(mdb_group *g)
{
    if (!g) return;
    for (Member &W : g->m) {
        if (!W.sys) continue;
        cudaSetDevice(W.device);
        for (int k = 0; k < 3; ++k)
            if (W.up[k]) cudaStreamSynchronize(W.up[k]);
        cudaStreamSynchronize(W.sys->stream);
    }
    for (Member &W : g->m) {
        if (!W.sys) continue;
        cudaSetDevice(W.device);
        for (DevBuf *b : {&W.cx, &W.cy, &W.cz, &W.sx, &W.sy, &W.sz, &W.sgid, &W.counts, &W.lab}) b->release();
        for (int k = 0; k < 3; ++k) {
            if (W.up[k]) cudaStreamDestroy(W.up[k]);
            if (W.up_ev[k]) cudaEventDestroy(W.up_ev[k]);
        }
        if (W.ready) cudaEventDestroy(W.ready);
        mdb_system_destroy(W.sys);
    }
    delete g;
}

int mdb_group_size(mdb_group *g) { return g ? g->D : 0; }

// Start the upload of one host frame (pageable or page-locked; it must stay valid until the next group call
// returns).  Returns once every copy is issued (page-locked input) or staged (pageable input).
int mdb_group_set_atoms(mdb_group *g, const double *x, const double *y, const double *z, int N, const double *box9,
                        const double *origin3, const int *boundary3)
{
    try {
        MDB_REQUIRE(g, MDB_ERR_VALUE, "group is NULL");
        MDB_REQUIRE(N > 0 && x && y && z, MDB_ERR_VALUE, "data must contain at least one atom.");
        MDB_REQUIRE(box9 && origin3 && boundary3, MDB_ERR_VALUE, "box, origin and boundary are required");
        DBox b;
        MDB_REQUIRE(dbox_make(b, box9, origin3, boundary3) == 0, MDB_ERR_BOX, "The volume of the box is zero.");
        for (int k = 0; k < 9; ++k) g->box9[k] = box9[k];
        for (int k = 0; k < 3; ++k) g->origin[k] = origin3[k], g->boundary[k] = boundary3[k];
        g->N = N;
        g->chunk = ((N + g->D - 1) / g->D + 255) / 256 * 256;
        g->t_upload0 = now_ms();
        for (int d = 0; d < g->D; ++d) {
            Member &W = g->m[d];
            W.start = (int)std::min<long long>((long long)d * g->chunk, N);
            W.count = std::min(g->chunk, N - W.start);
            CUDA_TRY(cudaSetDevice(W.device));
            double *p[3] = {W.cx.ensure<double>(std::max(W.count, 1)), W.cy.ensure<double>(std::max(W.count, 1)),
                            W.cz.ensure<double>(std::max(W.count, 1))};
            (void)p;
            // a recycled block may still be read by earlier work on the member's stream: the copy streams start after it
            CUDA_TRY(cudaEventRecord(W.ready, W.sys->stream));
            for (int k = 0; k < 3; ++k) CUDA_TRY(cudaStreamWaitEvent(W.up[k], W.ready, 0));
        }
        // one host thread per member; pageable input is staged by a share of the host cores each (staging.cu)
        const int threads = std::max(2, 2 * mdb_upload_threads() / g->D);
        for_members(*g, g->D, [&](int d) {
            Member &W = g->m[d];
            if (W.count) {
                void *dst[3] = {W.cx.as<double>(), W.cy.as<double>(), W.cz.as<double>()};
                const void *src[3] = {x + W.start, y + W.start, z + W.start};
                const size_t bytes[3] = {sizeof(double) * W.count, sizeof(double) * W.count, sizeof(double) * W.count};
                mdb_h2d(3, dst, src, bytes, W.up[0], threads);
            }
            for (int k = 0; k < 3; ++k) CUDA_TRY(cudaEventRecord(W.up_ev[k], W.up[0]));
        });
        g->has_atoms = true;
        g->times[0] = (float)(now_ms() - g->t_upload0);
    } catch (const MdbError &e) {
        return e.code;
    }
    return MDB_OK;
}

// FixedCNA labels (src/cna.cpp:429-506 on the list of src/neighbor.cpp:130-186) of the uploaded frame, original
// atom order, into pattern_host (N ints; page-locked memory makes the read-back asynchronous per member).
// *members = how many devices the frame ran on (1: gathered on the first member).
int mdb_group_fused_cna(mdb_group *g, double rc, int *pattern_host, int *members)
{
    try {
        MDB_REQUIRE(g && g->has_atoms, MDB_ERR_STATE, "no atoms uploaded");
        MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g.", rc);
        MDB_REQUIRE(pattern_host, MDB_ERR_VALUE, "pattern is required");
        const int D = g->D;
        DBox box;
        dbox_make(box, g->box9, g->origin, g->boundary);
        const CellGrid grid = cellgrid_make(box, rc);
        const double t0 = now_ms();
        for (int k = 1; k < 6; ++k) g->times[k] = 0.f;
        bool shard = D > 1 && grid.n[0] >= 3 * D;
        for (int d = 0; d < D && shard; ++d) shard = g->m[d].count > 0;
        if (shard) {
            RoutePlan P{};
            P.D = D;
            P.n0 = grid.n[0];
            for (int r = 0; r <= D; ++r) P.bounds[r] = (int)((long long)r * grid.n[0] / D);
            // ---- count
            for_members(*g, D, [&](int d) {
                Member &W = g->m[d];
                cudaStream_t st = W.sys->stream;
                wait_uploads(W);
                int *cnt = W.counts.ensure<int>(4 * MAX_DEV);
                CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int) * 4 * MAX_DEV, st));
                MDB_LAUNCH(k_route<false>, (W.count + 255) / 256, 256, 0, st, W.cx.as<double>(), W.cy.as<double>(),
                           W.cz.as<double>(), W.count, W.start, box, grid, P, RouteTable{}, cnt);
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaMemcpyAsync(W.h_counts, cnt, sizeof(int) * 2 * MAX_DEV, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
            });
            // ---- sizes: destination d holds [owned from source 0..D-1 | ghosts from source 0..D-1]
            for (int d = 0; d < D; ++d) {
                long long own = 0, gh = 0;
                for (int s = 0; s < D; ++s) own += g->m[s].h_counts[2 * d], gh += g->m[s].h_counts[2 * d + 1];
                MDB_REQUIRE(own + gh < 2147483647LL, MDB_ERR_VALUE, "slab of device %d has too many atoms", g->m[d].device);
                g->m[d].n_owned = (int)own;
                g->m[d].n_local = (int)(own + gh);
                shard = shard && own > 0;
            }
        }
        if (shard) {
            RoutePlan P{};
            P.D = D;
            P.n0 = grid.n[0];
            for (int r = 0; r <= D; ++r) P.bounds[r] = (int)((long long)r * grid.n[0] / D);
            std::vector<RouteTable> table(D);
            // ---- destination buffers (each on its own device; drained before any peer writes into them)
            for_members(*g, D, [&](int d) {
                Member &W = g->m[d];
                W.sx.ensure<double>(W.n_local);
                W.sy.ensure<double>(W.n_local);
                W.sz.ensure<double>(W.n_local);
                W.sgid.ensure<int>(W.n_local);
                W.lab.ensure<int>(std::max(W.count, 1));
                CUDA_TRY(cudaStreamSynchronize(W.sys->stream));
            });
            for (int d = 0; d < D; ++d) {
                int own = 0, gh = g->m[d].n_owned;
                for (int s = 0; s < D; ++s) {
                    RouteDest &R = table[s].d[d];
                    R.x = g->m[d].sx.as<double>();
                    R.y = g->m[d].sy.as<double>();
                    R.z = g->m[d].sz.as<double>();
                    R.gid = g->m[d].sgid.as<int>();
                    R.base[0] = own;
                    R.base[1] = gh;
                    own += g->m[s].h_counts[2 * d];
                    gh += g->m[s].h_counts[2 * d + 1];
                }
            }
            // ---- push
            for_members(*g, D, [&](int s) {
                Member &W = g->m[s];
                cudaStream_t st = W.sys->stream;
                int *cur = W.counts.as<int>() + 2 * MAX_DEV;   // zeroed with the counters above
                MDB_LAUNCH(k_route<true>, (W.count + 255) / 256, 256, 0, st, W.cx.as<double>(), W.cy.as<double>(),
                           W.cz.as<double>(), W.count, W.start, box, grid, P, table[s], cur);
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaStreamSynchronize(st));
            });
            g->times[1] = (float)(now_ms() - t0);
            // ---- slab compute
            const double t1 = now_ms();
            for_members(*g, D, [&](int d) {
                Member &W = g->m[d];
                const int lo = P.bounds[d], hi = P.bounds[d + 1];
                const int plane0 = (lo - 1 + P.n0) % P.n0, nplanes = hi - lo + 2;
                check_call(mdb_system_set_slab_device(W.sys, W.sx.as<double>(), W.sy.as<double>(), W.sz.as<double>(),
                                                      W.sgid.as<int>(), W.n_local, W.n_owned, plane0, nplanes, g->box9,
                                                      g->origin, g->boundary));
                check_call(mdb_system_set_local_fraction(W.sys, std::min(1.0, (double)nplanes / P.n0)));
                W.used = 0;
                check_call(mdb_system_fused_cna(W.sys, rc, nullptr, &W.used));
                CUDA_TRY(cudaStreamSynchronize(W.sys->stream));
            });
            g->times[2] = (float)(now_ms() - t1);
            for (int d = 0; d < D; ++d) shard = shard && g->m[d].used;
            if (shard) {
                const double t2 = now_ms();
                LabelTable LT{};
                for (int d = 0; d < D; ++d) LT.lab[d] = g->m[d].lab.as<int>();
                const int chunk = g->chunk;
                for_members(*g, D, [&](int d) {
                    Member &W = g->m[d];
                    int *pat = nullptr;
                    check_call(mdb_system_result_device(W.sys, &pat, nullptr));
                    MDB_LAUNCH(k_route_labels, (W.n_owned + 255) / 256, 256, 0, W.sys->stream, pat, W.sgid.as<int>(),
                               W.n_owned, chunk, LT);
                    CUDA_TRY(cudaGetLastError());
                    CUDA_TRY(cudaStreamSynchronize(W.sys->stream));
                });
                g->times[3] = (float)(now_ms() - t2);
                const double t3 = now_ms();
                for_members(*g, D, [&](int d) {
                    Member &W = g->m[d];
                    CUDA_TRY(cudaMemcpyAsync(pattern_host + W.start, W.lab.as<int>(), sizeof(int) * W.count,
                                             cudaMemcpyDeviceToHost, W.sys->stream));
                    CUDA_TRY(cudaStreamSynchronize(W.sys->stream));
                });
                g->times[4] = (float)(now_ms() - t3);
                g->sharded = D;
            }
        }
        if (!shard) {
            const double t1 = now_ms();
            gathered_cna(*g, rc, pattern_host);
            g->times[2] = (float)(now_ms() - t1);
        }
        if (members) *members = g->sharded;
        g->times[5] = (float)(now_ms() - g->t_upload0);
    } catch (const MdbError &e) {
        return e.code;
    }
    return MDB_OK;
}

// host-clock phase times of the last frame (ms): upload issue, route (count + push), slab compute, label push,
// label download, set_atoms-to-labels total
int mdb_group_last_times(mdb_group *g, float *ms6)
{
    if (!g || !ms6) return MDB_ERR_VALUE;
    for (int k = 0; k < 6; ++k) ms6[k] = g->times[k];
    return MDB_OK;
}

// atoms (owned, local = owned + ghosts) member d held in the last sharded frame
int mdb_group_member_atoms(mdb_group *g, int d, int *n_owned, int *n_local)
{
    if (!g || d < 0 || d >= g->D) return MDB_ERR_VALUE;
    if (n_owned) *n_owned = g->m[d].n_owned;
    if (n_local) *n_local = g->m[d].n_local;
    return MDB_OK;
}

}  // extern "C"
