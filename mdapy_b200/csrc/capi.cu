// mdapy_b200/csrc/capi.cu -- the C ABI declared in include/mdapy_b200.h.
#include "internal.cuh"
#include "../../include/mdapy_b200.h"
#include <cstdarg>
#include <new>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

std::atomic<long long> g_mdb_launches{0};
static thread_local char g_err[1024] = "";

void mdb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define API_BEGIN try {
#define API_END                                            \
    }                                                      \
    catch (const MdbError &e) { return e.code; }           \
    catch (const std::bad_alloc &)                         \
    {                                                      \
        mdb_set_error("host allocation failed");           \
        return MDB_ERR_CUDA;                               \
    }                                                      \
    return MDB_OK;

static void set_box(MdbSystem &s, const double *box9, const double *origin3, const int *boundary3)
{
    MDB_REQUIRE(box9 && origin3 && boundary3, MDB_ERR_VALUE, "box, origin and boundary are required");
    const int rc = dbox_make(s.box, box9, origin3, boundary3);
    MDB_REQUIRE(rc == 0, MDB_ERR_BOX, "The volume of the box is zero.");
    s.has_box = true;
}

static void invalidate(MdbSystem &s)
{
    s.bin_rc = -1.0;
    s.list_kind = LIST_NONE;
    s.list_rc = -1.0;
    s.M = 0;
    s.max_count = 0;
}

static void upload_atoms(MdbSystem &s, const double *x, const double *y, const double *z, int N)
{
    MDB_REQUIRE(N > 0, MDB_ERR_VALUE, "data must contain at least one atom.");
    MDB_REQUIRE(x && y && z, MDB_ERR_VALUE, "x, y, z are required");
    double *dx = s.bx.ensure<double>(N), *dy = s.by.ensure<double>(N), *dz = s.bz.ensure<double>(N);
    void *dst[3] = {dx, dy, dz};
    const void *src[3] = {x, y, z};
    const size_t bytes[3] = {sizeof(double) * N, sizeof(double) * N, sizeof(double) * N};
    mdb_h2d(3, dst, src, bytes, s.stream, mdb_upload_threads());
    s.x = dx;
    s.y = dy;
    s.z = dz;
    s.N = N;
    s.n_rows = N;
    s.gid = nullptr;
    s.slab_x0 = s.slab_nx = 0;
    s.local_frac = 1.0;
    invalidate(s);
}

static void prof_mark(MdbSystem &s, int k)
{
    if (s.profile) CUDA_TRY(cudaEventRecord(s.ev[k], s.stream));
}

// cut-off list into s.verlet/s.dist/s.nn.  max_neigh <= 0 -> automatic width M = max(count, 1)
// (neighbor.cpp:189-349).  Orthogonal frames with enough cells run the cell-tile kernel
// (neighbor_tiled.cu), everything else the direct kernel (neighbor.cu); both produce identical rows.
static void build_neighbor(MdbSystem &s, double rc, int max_neigh)
{
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g.", rc);
    prof_mark(s, 0);
    if (s.bin_rc != rc) launch_binning(s, rc);
    prof_mark(s, 1);
    int T = 0;
    const bool tiled = tiled_neighbor_plan(s, T);
    auto fill = [&](int M) {
        prof_mark(s, 1);
        if (tiled) launch_neighbor_tiled(s, rc, M, T, false, 1);
        else launch_neighbor(s, rc, M, false);
        prof_mark(s, 2);
        return tiled ? neighbor_tiled_max(s) : device_max_int(s, s.nn.as<int>(), s.n_rows);
    };
    if (max_neigh > 0) {
        s.max_count = fill(max_neigh);
        s.M = max_neigh;
    } else {
        // Width: exact count pass (direct kernel), or a count-only pass over every 16th tile (tiled).
        // A sample with min == max is a uniform frame (perfect lattice): the fill uses exactly that
        // width.  Otherwise the true maximum is an extreme value the sample probably missed: the fill
        // uses the next multiple of 4 above estimate + 1 (rows then leave in 16-byte stores) and the
        // rows are compacted to the true maximum afterwards -- one streaming pass instead of a second
        // search.  Only when even that margin was too small is the fill repeated.
        int est, est_min = 0;
        bool exact = !tiled;
        if (tiled && s.hint_rc == rc && s.hint_M > 0) {
            est = s.hint_M;             // previous frame on this handle, same cut-off
            exact = s.hint_uniform;
        } else if (tiled) {
            launch_neighbor_tiled(s, rc, 0, T, true, 16);
            est = neighbor_tiled_max(s, &est_min);
            if (est == est_min) exact = true;
        } else {
            launch_neighbor(s, rc, 0, true);
            est = device_max_int(s, s.nn.as<int>(), s.n_rows);
        }
        if (est < 1) est = 1;
        int width = exact ? est : (est + 1 + 3) / 4 * 4;
        int mx = fill(width);
        bool refilled = false;
        if (mx > width) {
            width = (mx + 3) / 4 * 4;
            mx = fill(width);
            refilled = true;
        }
        const int M = mx < 1 ? 1 : mx;
        if (M < width) launch_compact_rows(s, width, M);
        s.M = M;
        s.max_count = mx;
        s.hint_rc = rc;
        s.hint_M = M;
        s.hint_uniform = exact && !refilled && M == width;
    }
    s.list_kind = LIST_CUTOFF;
    s.list_rc = rc;
    s.has_dist = true;
    if (s.profile) {
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        CUDA_TRY(cudaEventElapsedTime(&s.t_bin, s.ev[0], s.ev[1]));
        CUDA_TRY(cudaEventElapsedTime(&s.t_neigh, s.ev[1], s.ev[2]));
    }
}

template <class T> static void d2h(MdbSystem &s, T *host, const T *dev, size_t n)
{
    if (host && n) CUDA_TRY(cudaMemcpyAsync(host, dev, sizeof(T) * n, cudaMemcpyDeviceToHost, s.stream));
}

template <class T> static T *h2d(MdbSystem &s, DevBuf &buf, const T *host, size_t n)
{
    T *d = buf.ensure<T>(n ? n : 1);
    if (n) CUDA_TRY(cudaMemcpyAsync(d, host, sizeof(T) * n, cudaMemcpyHostToDevice, s.stream));
    return d;
}

// ---- lists handed in by the caller (mdb_system_put_neighbor) may come without counts or distances
// (the reference's L3 classes accept a bare verlet_list): counts follow from the -1 padding, distances are
// recomputed on first use with the list builder's arithmetic (xi wrapped, x[j] raw, minimum image;
// neighbor.cpp:130-186).
__global__ void __launch_bounds__(256) k_counts_from_padding(const int *__restrict__ verlet, int rows, int M,
                                                             int *__restrict__ nn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int *row = verlet + (size_t)i * M;
    int c = 0;
    while (c < M && row[c] >= 0) ++c;
    nn[i] = c;
}

__global__ void __launch_bounds__(256) k_dist_from_verlet(const double *__restrict__ x, const double *__restrict__ y,
                                                          const double *__restrict__ z, DBox box,
                                                          const int *__restrict__ verlet, int rows, int M, double pad,
                                                          double *__restrict__ dist)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)rows * M) return;
    const int i = (int)(e / M), j = verlet[e];
    if (j < 0) {
        dist[e] = pad;
        return;
    }
    double xi = x[i], yi = y[i], zi = z[i];
    if (box.any_pbc) wrap_into_box(box, xi, yi, zi);
    double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
    min_image(box, dx, dy, dz);
    dist[e] = sqrt(dx * dx + dy * dy + dz * dz);
}

static double *list_dist(MdbSystem &s)
{
    if (!s.has_dist && s.list_kind != LIST_NONE) {
        const size_t n = (size_t)s.n_rows * s.M;
        double *d = s.dist.ensure<double>(n ? n : 1);
        if (n) {
            const double pad = s.list_rc > 0 ? s.list_rc + 1.0 : 1.0e300;
            MDB_LAUNCH(k_dist_from_verlet, (unsigned)((n + 255) / 256), 256, 0, s.stream, s.x, s.y, s.z, s.box,
                       s.verlet.as<int>(), s.n_rows, s.M, pad, d);
            CUDA_TRY(cudaGetLastError());
        }
        s.has_dist = true;
    }
    return s.dist.as<double>();
}

static void require_list(MdbSystem &s)
{
    MDB_REQUIRE(s.list_kind != LIST_NONE, MDB_ERR_STATE, "no neighbour list on the device; build one first");
}

// ---------------------------------------------------------------- block caches
namespace {
std::mutex g_pool_mu;
struct CachedBlock {
    void *p;
    cudaEvent_t ready;  // recorded on the last user's stream at release (nullptr: released after a host sync)
};
std::map<std::pair<int, size_t>, std::vector<CachedBlock>> g_dev_free;  // (device, bytes) -> blocks
size_t g_dev_cached = 0;
std::multimap<size_t, void *> g_host_free;                         // bytes -> pinned block
std::unordered_map<void *, size_t> g_host_live;
size_t g_host_cached = 0;
constexpr size_t HOST_CACHE_LIMIT = (size_t)12 << 30;

size_t round_block(size_t bytes)
{   // 2 MiB granules below 1 GiB, 64 MiB above: keeps successive frames of similar size on the same blocks
    const size_t g = bytes < ((size_t)1 << 30) ? ((size_t)2 << 20) : ((size_t)64 << 20);
    return bytes < 65536 ? ((bytes + 511) & ~(size_t)511) : (bytes + g - 1) / g * g;
}

void trim_device_locked(int device)
{
    for (auto it = g_dev_free.begin(); it != g_dev_free.end();) {
        if (device < 0 || it->first.first == device) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(it->first.first);
            for (CachedBlock &q : it->second) {
                if (q.ready) {
                    cudaEventSynchronize(q.ready);
                    cudaEventDestroy(q.ready);
                }
                cudaFree(q.p);
                g_dev_cached -= it->first.second;
            }
            cudaSetDevice(cur);
            it = g_dev_free.erase(it);
        } else ++it;
    }
}
}  // namespace

void *mdb_pool_alloc(size_t bytes, size_t *got, cudaStream_t user)
{
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    const size_t want = round_block(bytes);
    {
        CachedBlock hit{nullptr, nullptr};
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            // smallest cached block that fits without wasting more than half of it
            auto it = g_dev_free.lower_bound({device, want});
            while (it != g_dev_free.end() && it->first.first == device && it->first.second <= want + want / 2 + (1 << 20)) {
                if (!it->second.empty()) {
                    hit = it->second.back();
                    it->second.pop_back();
                    g_dev_cached -= it->first.second;
                    *got = it->first.second;
                    break;
                }
                ++it;
            }
        }
        if (hit.p) {
            if (hit.ready) {   // the previous owner's queued work comes first
                if (user) CUDA_TRY(cudaStreamWaitEvent(user, hit.ready, 0));
                else CUDA_TRY(cudaEventSynchronize(hit.ready));
                cudaEventDestroy(hit.ready);
            }
            return hit.p;
        }
    }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, want);
    if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            trim_device_locked(device);
        }
        e = cudaMalloc(&q, want);
    }
    CUDA_TRY(e);
    *got = want;
    return q;
}

void mdb_pool_free(void *p, size_t bytes, cudaStream_t last_user)
{
    if (!p) return;
    int device = 0;
    cudaGetDevice(&device);
    static const bool off = getenv("MDB_NO_CACHE") != nullptr;
    if (off) {
        cudaFree(p);   // synchronises with all outstanding work by itself
        return;
    }
    cudaEvent_t ready = nullptr;
    if (last_user) {
        if (cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(ready, last_user) != cudaSuccess) {
            if (ready) cudaEventDestroy(ready);
            ready = nullptr;
            cudaStreamSynchronize(last_user);   // no event: fall back to a host-side wait
        }
    } else {
        cudaDeviceSynchronize();   // unknown owner: be safe
    }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_dev_free[{device, bytes}].push_back(CachedBlock{p, ready});
    g_dev_cached += bytes;
}

int mdb_trim_cache(void)
{
    API_BEGIN
    std::lock_guard<std::mutex> lk(g_pool_mu);
    trim_device_locked(-1);
    for (auto &kv : g_host_free) cudaFreeHost(kv.second);
    g_host_free.clear();
    g_host_cached = 0;
    API_END
}

int mdb_host_alloc(size_t bytes, void **ptr)
{
    API_BEGIN
    MDB_REQUIRE(ptr, MDB_ERR_VALUE, "ptr is required");
    const size_t want = round_block(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_host_free.lower_bound(want);
        if (it != g_host_free.end() && it->first <= want + want / 2 + (1 << 20)) {
            *ptr = it->second;
            g_host_live[it->second] = it->first;
            g_host_cached -= it->first;
            g_host_free.erase(it);
            return MDB_OK;
        }
    }
    void *q = nullptr;
    CUDA_TRY(cudaHostAlloc(&q, want, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_host_live[q] = want;
    *ptr = q;
    API_END
}

void mdb_host_free(void *ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_host_live.find(ptr);
    if (it == g_host_live.end()) return;
    const size_t bytes = it->second;
    g_host_live.erase(it);
    if (g_host_cached + bytes > HOST_CACHE_LIMIT) {
        cudaFreeHost(ptr);
        return;
    }
    g_host_free.emplace(bytes, ptr);
    g_host_cached += bytes;
}

extern "C" {

const char *mdb_last_error(void) { return g_err; }
const char *mdb_version(void) { return "mdapy_b200 0.1 (sm_100a)"; }
long long mdb_launch_count(void) { return g_mdb_launches.load(); }

int mdb_device_count(int *count)
{
    API_BEGIN
    CUDA_TRY(cudaGetDeviceCount(count));
    API_END
}

int mdb_system_create(int device, mdb_system **out)
{
    API_BEGIN
    MDB_REQUIRE(out, MDB_ERR_VALUE, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        mdb_set_error("no CUDA device available (%s); mdapy_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MDB_ERR_CUDA;
    }
    MDB_REQUIRE(device >= 0 && device < ndev, MDB_ERR_VALUE, "device %d out of range [0,%d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    MdbSystem *s = new MdbSystem();
    s->bind_buffers();
    s->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
    for (int k = 0; k < 4; ++k) CUDA_TRY(cudaEventCreate(&s->ev[k]));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) CUDA_TRY(cudaEventCreateWithFlags(&s->chunk_ev[k], cudaEventDisableTiming));
    *out = s;
    API_END
}

void mdb_system_destroy(mdb_system *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    // both streams drain BEFORE any block goes back to the process-wide cache
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    cudaStreamSynchronize(s->stream);
    s->for_each_buffer([](DevBuf &b) { b.release(); });
    for (int k = 0; k < 4; ++k)
        if (s->ev[k]) cudaEventDestroy(s->ev[k]);
    for (int k = 0; k < 2; ++k)
        if (s->chunk_ev[k]) cudaEventDestroy(s->chunk_ev[k]);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int mdb_system_set_stream(mdb_system *s, void *cuda_stream)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) CUDA_TRY(cudaStreamDestroy(s->stream));
    s->stream = static_cast<cudaStream_t>(cuda_stream);
    s->own_stream = false;
    API_END
}

int mdb_system_synchronize(mdb_system *s)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_set_atoms(mdb_system *s, const double *x, const double *y, const double *z, int N,
                         const double *box9, const double *origin3, const int *boundary3)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    API_END
}

int mdb_system_set_atoms_device(mdb_system *s, const double *dx, const double *dy, const double *dz, int N,
                                const double *box9, const double *origin3, const int *boundary3)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(N > 0 && dx && dy && dz, MDB_ERR_VALUE, "data must contain at least one atom.");
    set_box(*s, box9, origin3, boundary3);
    s->x = dx;
    s->y = dy;
    s->z = dz;
    s->N = N;
    s->n_rows = N;
    s->gid = nullptr;
    s->slab_x0 = s->slab_nx = 0;
    invalidate(*s);
    API_END
}

int mdb_system_set_slab_device(mdb_system *s, const double *dx, const double *dy, const double *dz,
                               const int *dgid, int n_local, int n_owned, int plane0, int nplanes,
                               const double *box9, const double *origin3, const int *boundary3)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(n_local > 0 && dx && dy && dz && dgid, MDB_ERR_VALUE, "slab needs coordinates and global ids");
    MDB_REQUIRE(n_owned > 0 && n_owned <= n_local, MDB_ERR_VALUE, "n_owned=%d must be in (0, n_local=%d]", n_owned,
                n_local);
    MDB_REQUIRE(nplanes >= 3, MDB_ERR_VALUE, "a slab window needs >= 3 planes (ghost, owned, ghost), got %d", nplanes);
    set_box(*s, box9, origin3, boundary3);
    s->x = dx;
    s->y = dy;
    s->z = dz;
    s->gid = dgid;
    s->N = n_local;
    s->n_rows = n_owned;
    s->slab_x0 = plane0;
    s->slab_nx = nplanes;
    s->local_frac = 1.0;
    invalidate(*s);
    API_END
}

int mdb_system_set_local_fraction(mdb_system *s, double fraction)
{
    API_BEGIN
    MDB_REQUIRE(fraction > 0.0 && fraction <= 1.0, MDB_ERR_VALUE, "fraction must be in (0, 1], got %g", fraction);
    s->local_frac = fraction;
    API_END
}

int mdb_cell_grid(const double *box9, const double *origin3, const int *boundary3, double rc, int *n3)
{
    API_BEGIN
    MDB_REQUIRE(rc > 0 && n3, MDB_ERR_VALUE, "rc must be positive");
    DBox b;
    MDB_REQUIRE(dbox_make(b, box9, origin3, boundary3) == 0, MDB_ERR_BOX, "The volume of the box is zero.");
    const CellGrid g = cellgrid_make(b, rc);
    n3[0] = g.n[0];
    n3[1] = g.n[1];
    n3[2] = g.n[2];
    API_END
}

int mdb_cell_planes_device(const double *dx, const double *dy, const double *dz, int N, const double *box9,
                           const double *origin3, const int *boundary3, double rc, int *dplane, void *cuda_stream)
{
    API_BEGIN
    MDB_REQUIRE(rc > 0 && dplane, MDB_ERR_VALUE, "rc must be positive");
    DBox b;
    MDB_REQUIRE(dbox_make(b, box9, origin3, boundary3) == 0, MDB_ERR_BOX, "The volume of the box is zero.");
    launch_cell_planes(dx, dy, dz, N, b, cellgrid_make(b, rc), dplane, static_cast<cudaStream_t>(cuda_stream));
    API_END
}

int mdb_slab_pack_device(const double *dx, const double *dy, const double *dz, const int *dgid, int N,
                         const double *box9, const double *origin3, const int *boundary3, double rc, int lo, int hi,
                         int halo, double *send_left, double *send_right, int cap, int *dcounts, void *cuda_stream)
{
    API_BEGIN
    MDB_REQUIRE(rc > 0 && send_left && send_right && dcounts && cap > 1, MDB_ERR_VALUE, "bad slab pack arguments");
    DBox b;
    MDB_REQUIRE(dbox_make(b, box9, origin3, boundary3) == 0, MDB_ERR_BOX, "The volume of the box is zero.");
    launch_slab_pack(dx, dy, dz, dgid, N, b, cellgrid_make(b, rc), lo, hi, halo, send_left, send_right, cap, dcounts,
                     static_cast<cudaStream_t>(cuda_stream));
    API_END
}

int mdb_slab_unpack_device(const double *recv_a, const double *recv_b, int cap, double *dx, double *dy, double *dz,
                           int *dgid, int n_owned, int room, int *dtotal, void *cuda_stream)
{
    API_BEGIN
    MDB_REQUIRE(recv_a && dtotal && cap > 1, MDB_ERR_VALUE, "bad slab unpack arguments");
    launch_slab_unpack(recv_a, recv_b, cap, dx, dy, dz, dgid, n_owned, room, dtotal, static_cast<cudaStream_t>(cuda_stream));
    API_END
}

int mdb_system_build_neighbor(mdb_system *s, double rc, int max_neigh, int *M, int *max_count)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box, MDB_ERR_STATE, "no atoms uploaded");
    build_neighbor(*s, rc, max_neigh);
    if (M) *M = s->M;
    if (max_count) *max_count = s->max_count;
    API_END
}

int mdb_system_build_knn(mdb_system *s, int k)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box, MDB_ERR_STATE, "no atoms uploaded");
    launch_knn(*s, k);
    API_END
}

int mdb_system_sort_neighbor(mdb_system *s, int k)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    launch_sort_rows(*s, s->verlet.as<int>(), list_dist(*s), s->n_rows, s->M, k);
    API_END
}

int mdb_system_neighbor_min_count(mdb_system *s, int *min_count)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    *min_count = device_min_int(*s, s->nn.as<int>(), s->n_rows);
    API_END
}

int mdb_system_fetch_neighbor(mdb_system *s, int *verlet, double *dist, int *nn)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    const size_t n = (size_t)s->n_rows * s->M;
    if (verlet && s->gid) {  // decomposed frame: rows hold local indices, export global ids
        int *tmp = s->verlet_tmp.ensure<int>(n);
        launch_translate_ids(*s, s->verlet.as<int>(), tmp, n);
        d2h(*s, verlet, tmp, n);
    } else {
        d2h(*s, verlet, s->verlet.as<int>(), n);
    }
    d2h(*s, dist, list_dist(*s), n);
    d2h(*s, nn, s->nn.as<int>(), (size_t)s->n_rows);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_put_neighbor(mdb_system *s, const int *verlet, const double *dist, const int *nn, int M, double rc,
                            int kind)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0, MDB_ERR_STATE, "no atoms uploaded");
    MDB_REQUIRE(M > 0 && verlet, MDB_ERR_VALUE, "verlet list with M > 0 required");
    const size_t n = (size_t)s->n_rows * M;
    h2d(*s, s->verlet, verlet, n);
    if (dist) h2d(*s, s->dist, dist, n);
    s->has_dist = dist != nullptr;   // recomputed on first use otherwise (list_dist)
    if (nn) h2d(*s, s->nn, nn, (size_t)s->n_rows);
    else {
        MDB_LAUNCH(k_counts_from_padding, (s->n_rows + 255) / 256, 256, 0, s->stream, s->verlet.as<int>(), s->n_rows, M,
                   s->nn.ensure<int>(s->n_rows));
        CUDA_TRY(cudaGetLastError());
    }
    s->M = M;
    s->list_rc = rc;
    s->list_kind = kind ? kind : LIST_CUTOFF;
    API_END
}

int mdb_system_neighbor_device(mdb_system *s, int **verlet, double **dist, int **nn, int *M)
{
    API_BEGIN
    require_list(*s);
    if (verlet) *verlet = s->verlet.as<int>();
    if (dist) *dist = list_dist(*s);
    if (nn) *nn = s->nn.as<int>();
    if (M) *M = s->M;
    API_END
}

int mdb_system_fcna(mdb_system *s, double rc, int *pattern_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g.", rc);
    int *pat = s->out_i32.ensure<int>(s->n_rows);
    CUDA_TRY(cudaMemsetAsync(pat, 0, sizeof(int) * s->n_rows, s->stream));
    prof_mark(*s, 2);
    // Large frames with a host destination: the labels of one chunk travel to the host (copy stream) while the
    // next chunk is classified, so only the last chunk's copy is exposed.
    const int R = s->n_rows;
    const int nchunk = (pattern_host && R >= (1 << 22)) ? 8 : 1;
    if (nchunk == 1) {
        launch_fcna(*s, s->verlet.as<int>(), s->nn.as<int>(), s->M, rc, pat);
        prof_mark(*s, 3);
        d2h(*s, pattern_host, pat, (size_t)R);
    } else {
        const int step = ((R + nchunk - 1) / nchunk + 127) / 128 * 128;
        for (int c = 0, first = 0; first < R; ++c, first += step) {
            const int count = first + step <= R ? step : R - first;
            launch_fcna(*s, s->verlet.as<int>(), s->nn.as<int>(), s->M, rc, pat, first, count);
            cudaEvent_t ev = s->chunk_ev[c & 1];
            CUDA_TRY(cudaEventRecord(ev, s->stream));
            CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, ev, 0));
            CUDA_TRY(cudaMemcpyAsync(pattern_host + first, pat + first, sizeof(int) * (size_t)count,
                                     cudaMemcpyDeviceToHost, s->copy_stream));
        }
        prof_mark(*s, 3);
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    }
    if (pattern_host || s->profile) CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->profile) CUDA_TRY(cudaEventElapsedTime(&s->t_cna, s->ev[2], s->ev[3]));
    API_END
}

// Fused neighbour search + fixed-cutoff CNA: labels without a neighbour list in HBM (neighbor_tiled.cu,
// k_fused_cna).  *used = 1 when the fused kernel produced the labels, 0 when the frame is not eligible
// (triclinic / tiny box / overflow tiles) -- nothing was computed then and the caller takes the list path.
int mdb_system_fused_cna(mdb_system *s, double rc, int *pattern_host, int *used)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box, MDB_ERR_STATE, "no atoms uploaded");
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g.", rc);
    MDB_REQUIRE(used, MDB_ERR_VALUE, "used is required");
    *used = 0;
    prof_mark(*s, 0);
    if (s->bin_rc != rc) launch_binning(*s, rc);
    prof_mark(*s, 1);
    int *pat = s->out_i32.ensure<int>(s->n_rows);
    int left = 0;
    const bool ok = launch_fused_cna(*s, rc, pat, &left);
    prof_mark(*s, 2);
    if (ok && left == 0) {
        *used = 1;
        d2h(*s, pattern_host, pat, (size_t)s->n_rows);
        if (pattern_host || s->profile) CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->profile) {
            CUDA_TRY(cudaEventElapsedTime(&s->t_bin, s->ev[0], s->ev[1]));
            CUDA_TRY(cudaEventElapsedTime(&s->t_neigh, s->ev[1], s->ev[2]));
            s->t_cna = 0.f;
        }
    }
    API_END
}

int mdb_system_acna(mdb_system *s, int *pattern_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    int *pat = s->out_i32.ensure<int>(s->n_rows);
    CUDA_TRY(cudaMemsetAsync(pat, 0, sizeof(int) * s->n_rows, s->stream));
    launch_acna(*s, s->verlet.as<int>(), s->M, pat);
    d2h(*s, pattern_host, pat, (size_t)s->n_rows);
    if (pattern_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_ids(mdb_system *s, int *pattern_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(s->n_rows == s->N || s->list_kind == LIST_KNN, MDB_ERR_STATE,
                "diamond identification on a decomposed frame needs the k-nearest list (rows for ghosts too)");
    int *pat = s->out_i32.ensure<int>(s->N);
    CUDA_TRY(cudaMemsetAsync(pat, 0, sizeof(int) * s->N, s->stream));
    launch_ids(*s, s->verlet.as<int>(), s->M, nullptr, pat);
    d2h(*s, pattern_host, pat, (size_t)s->n_rows);
    if (pattern_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_csp(mdb_system *s, int nnei, double *csp_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    double *out = s->out_f64.ensure<double>(s->n_rows);
    launch_csp(*s, s->verlet.as<int>(), s->M, nnei, out);
    d2h(*s, csp_host, out, (size_t)s->n_rows);
    if (csp_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_aja(mdb_system *s, int *aja_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    int *out = s->out_i32.ensure<int>(s->n_rows);
    launch_aja(*s, s->verlet.as<int>(), s->M, list_dist(*s), s->M, out);
    d2h(*s, aja_host, out, (size_t)s->n_rows);
    if (aja_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// qlm_r/qlm_i/qnarray hosts may be NULL (results stay on the device).  weight_host: (n_rows, M) or NULL.
int mdb_system_steinhardt(mdb_system *s, const int *llist, int ndeg, int nnn, double rc, int average, int wl,
                          int wlhat, int use_voronoi, const double *weight_host, double *qnarray_host,
                          double *qlm_r_host, double *qlm_i_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(llist && ndeg > 0, MDB_ERR_VALUE, "llist is required");
    int lmax = 0;
    for (int i = 0; i < ndeg; ++i) lmax = llist[i] > lmax ? llist[i] : lmax;
    const int nz = 2 * lmax + 1, R = s->n_rows;
    const int ncol = ndeg * (1 + (wl ? 1 : 0) + (wlhat ? 1 : 0));
    const size_t nq = (size_t)R * ndeg * nz;
    // arrays indexed by NEIGHBOUR ids cover every local atom (s->N >= n_rows: ghosts of a decomposed
    // frame keep zero rows), results are read back for the n_rows row atoms
    const size_t nq_all = (size_t)s->N * ndeg * nz;
    double *qr = s->qlm_r.ensure<double>(nq_all), *qi = s->qlm_i.ensure<double>(nq_all);
    double *qn = s->qn.ensure<double>((size_t)s->N * ncol);
    CUDA_TRY(cudaMemsetAsync(qr, 0, sizeof(double) * nq_all, s->stream));
    CUDA_TRY(cudaMemsetAsync(qi, 0, sizeof(double) * nq_all, s->stream));
    CUDA_TRY(cudaMemsetAsync(qn, 0, sizeof(double) * (size_t)s->N * ncol, s->stream));
    const double *w = nullptr;
    if (weight_host) w = h2d(*s, s->weight, weight_host, (size_t)R * s->M);
    // steinhardt_bond_orientation.py:238-245: a huge rc for the voronoi / nnn neighbour sources
    double rc_eff = rc;
    if (use_voronoi) rc_eff = 10000000000.0;
    else if (nnn > 0) rc_eff = 1000000000.0;
    else MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "At least use voronoi, or set positive nnn, or positive rc.");
    launch_steinhardt(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, w, llist, ndeg, nnn, lmax,
                      wl != 0, wlhat != 0, average != 0, use_voronoi != 0, rc_eff, weight_host != nullptr, qr, qi, qn);
    s->sbo_ndeg = ndeg;
    s->sbo_nz = nz;
    s->sbo_ncol = ncol;
    d2h(*s, qnarray_host, qn, (size_t)R * ncol);
    d2h(*s, qlm_r_host, qr, nq);
    d2h(*s, qlm_i_host, qi, nq);
    if (qnarray_host || qlm_r_host || qlm_i_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// uses q_lm / q_l of the latest mdb_system_steinhardt call
int mdb_system_solid_liquid(mdb_system *s, int q6index, double threshold, int n_bond, int use_voronoi, int nnn,
                            double rc, int *solidliquid_host, int *nbond_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(s->sbo_ndeg > 0, MDB_ERR_STATE, "run mdb_system_steinhardt first");
    MDB_REQUIRE(q6index >= 0 && q6index < s->sbo_ndeg, MDB_ERR_VALUE, "Q6index %d out of range", q6index);
    const int R = s->n_rows;
    double rc_eff = rc;
    if (use_voronoi) rc_eff = 10000000000.0;
    else if (nnn > 0) rc_eff = 1000000000.0;
    // contiguous copy of the q6 column (np.ascontiguousarray(qnarray[:, Q6index]) in the reference wrapper)
    const int NA = s->N;  // arrays indexed by neighbour ids cover the ghosts of a decomposed frame too
    double *q6 = s->out_f64.ensure<double>(NA);
    CUDA_TRY(cudaMemcpy2DAsync(q6, sizeof(double), s->qn.as<double>() + q6index, sizeof(double) * s->sbo_ncol,
                               sizeof(double), NA, cudaMemcpyDeviceToDevice, s->stream));
    int *solid = s->out_i32.ensure<int>((size_t)2 * NA);
    int *nbond = solid + NA;
    CUDA_TRY(cudaMemsetAsync(solid, 0, sizeof(int) * 2 * (size_t)NA, s->stream));
    launch_solid_liquid(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, q6index, q6,
                        s->qlm_r.as<double>(), s->qlm_i.as<double>(), s->sbo_ndeg, s->sbo_nz, threshold, n_bond,
                        use_voronoi != 0, nnn, rc_eff, solid, nbond);
    d2h(*s, solidliquid_host, solid, (size_t)R);
    d2h(*s, nbond_host, nbond, (size_t)R);
    if (solidliquid_host || nbond_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// g_host (ntype*ntype*nbin doubles, or nbin when types_host == NULL) is ACCUMULATED into, like the reference
int mdb_system_rdf(mdb_system *s, const int *types_host, int ntype, double rc, int nbin, int streaming,
                   double *g_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(g_host && nbin > 0, MDB_ERR_VALUE, "g and nbin are required");
    const int *types = nullptr;
    if (types_host) types = h2d(*s, s->types, types_host, (size_t)s->N);
    const int nslot = (types ? ntype * ntype : 1) * nbin;
    double *g = h2d(*s, s->out_f64b, g_host, (size_t)nslot);
    if (streaming) {
        MDB_REQUIRE(types, MDB_ERR_VALUE, "streaming RDF needs a type list");
        launch_rdf_streaming(*s, types, ntype, rc, nbin, g);
    } else {
        require_list(*s);
        launch_rdf_list(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->n_rows, s->M, types, ntype,
                        rc, nbin, g);
    }
    d2h(*s, g_host, g, (size_t)nslot);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// PTM on the cached list (>= 18 sorted neighbours per row, or fewer for open clusters).  output_host:
int mdb_system_cnp(mdb_system *s, double rc, double *cnp_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc must be positive, got %g.", rc);
    MDB_REQUIRE(s->n_rows == s->N, MDB_ERR_STATE, "the common neighbour parameter reads its neighbours' rows: use a halo >= 2 frame "
                                                   "through distributed.py (rows for the inner ghost layer)");
    double *out = s->out_f64.ensure<double>(s->n_rows);
    launch_cnp(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, rc, out);
    d2h(*s, cnp_host, out, (size_t)s->n_rows);
    if (cnp_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// WCP[i][j] = 1 - Z_ij / (alpha_j * Z_i) from integer counts, normalised on the host with the
// reference's expressions (warren_cowley_parameter.cpp:67-79)
int mdb_system_wcp(mdb_system *s, const int *types_host, int ntype, double *wcp_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(types_host && wcp_host && ntype >= 1, MDB_ERR_VALUE, "type list, Ntype and WCP are required");
    const int *types = h2d(*s, s->types, types_host, (size_t)s->N);
    const int nslot = ntype * ntype + 2 * ntype;
    unsigned long long *counts = s->scratch2.ensure<unsigned long long>(nslot);
    launch_wcp_counts(*s, s->verlet.as<int>(), s->nn.as<int>(), s->M, types, ntype, counts);
    std::vector<unsigned long long> h(nslot);
    CUDA_TRY(cudaMemcpyAsync(h.data(), counts, sizeof(unsigned long long) * nslot, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const unsigned long long *Zmn = h.data(), *Zm = Zmn + ntype * ntype, *pop = Zm + ntype;
    const int N = s->n_rows;
    for (int i = 0; i < ntype; ++i)
        for (int j = 0; j < ntype; ++j) {
            const double alpha_j = (double)pop[j] / N;
            const int zm_i = (int)Zm[i];
            if (alpha_j > 0 && zm_i > 0)
                wcp_host[i * ntype + j] = 1.0 - static_cast<double>((int)Zmn[i * ntype + j]) / (alpha_j * zm_i);
            else wcp_host[i * ntype + j] = 0.0;
        }
    API_END
}

int mdb_system_average_by_neighbor(mdb_system *s, double rc, const double *value_host, int include_self,
                                   double *value_ave_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(value_host, MDB_ERR_VALUE, "value is required");
    const double *val = h2d(*s, s->out_f64b, value_host, (size_t)s->N);
    double *out = s->out_f64.ensure<double>(s->n_rows);
    launch_average_by_neighbor(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, rc, val,
                               include_self != 0, out);
    d2h(*s, value_ave_host, out, (size_t)s->n_rows);
    if (value_ave_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_cluster(mdb_system *s, double rc, const int *types_host, const int *type1, const int *type2,
                       const double *r, int npair, int *cluster_host, int *cluster_number)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(s->n_rows == s->N, MDB_ERR_STATE, "cluster analysis needs the whole frame on one device");
    const int R = s->n_rows, M = s->M;
    int *out = s->out_i32.ensure<int>(R);
    int count;
    if (npair > 0) {
        MDB_REQUIRE(types_host && type1 && type2 && r, MDB_ERR_VALUE, "Need type_list for multi cutoff mode.");
        const int *types = h2d(*s, s->types, types_host, (size_t)s->N);
        // filtered copy of the list (the reference works on verlet_list.copy(), cluster_analysis.py:66-69)
        int *vcopy = s->verlet_tmp.ensure<int>((size_t)R * M);
        CUDA_TRY(cudaMemcpyAsync(vcopy, s->verlet.as<int>(), sizeof(int) * (size_t)R * M, cudaMemcpyDeviceToDevice,
                                 s->stream));
        int *pairs = s->scratch2.ensure<int>((size_t)2 * npair + 2);
        double *rr = s->out_f64c.ensure<double>(npair);
        CUDA_TRY(cudaMemcpyAsync(pairs, type1, sizeof(int) * npair, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(pairs + npair, type2, sizeof(int) * npair, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(rr, r, sizeof(double) * npair, cudaMemcpyHostToDevice, s->stream));
        launch_filter_by_type(*s, vcopy, list_dist(*s), s->nn.as<int>(), M, types, pairs, pairs + npair, rr, npair);
        count = launch_cluster(*s, vcopy, nullptr, s->nn.as<int>(), M, 0.0, out);
    } else {
        MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "rc should be a positive number, got %g.", rc);
        count = launch_cluster(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), M, rc, out);
    }
    if (cluster_number) *cluster_number = count;
    d2h(*s, cluster_host, out, (size_t)R);
    if (cluster_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_structure_entropy(mdb_system *s, double rc, double sigma, int use_local_density, double volume,
                                 double average_rc, double *entropy_host, double *entropy_ave_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    const int R = s->n_rows;
    MDB_REQUIRE(average_rc <= 0 || R == s->N, MDB_ERR_STATE, "the neighbour average needs rows for every listed atom");
    double *ent = s->out_f64.ensure<double>((size_t)2 * R);
    launch_structure_entropy(*s, list_dist(*s), s->nn.as<int>(), s->M, rc, sigma, use_local_density != 0, volume,
                             ent);
    d2h(*s, entropy_host, ent, (size_t)R);
    if (average_rc > 0) {
        MDB_REQUIRE(average_rc <= rc, MDB_ERR_VALUE, "average_rc should be smaller than rc.");
        launch_average_by_neighbor(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, average_rc, ent,
                                   true, ent + R);
        d2h(*s, entropy_ave_host, ent + R, (size_t)R);
    }
    if (entropy_host || entropy_ave_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_check_small_division(mdb_system *s, const double *a_host, int n, int d, long long *mismatches)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(a_host && n > 0 && mismatches, MDB_ERR_VALUE, "values and the output are required");
    const double *a = h2d(*s, s->out_f64c, a_host, (size_t)n);
    *mismatches = sbo_div_small_mismatches(*s, a, n, d);
    API_END
}

int mdb_system_atomic_temperature(mdb_system *s, const double *vx, const double *vy, const double *vz,
                                  const double *mass, double rc, double *T_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(vx && vy && vz && mass, MDB_ERR_VALUE, "No velocity information.");
    const size_t NA = (size_t)s->N;
    double *buf = s->out_f64b.ensure<double>(4 * NA);
    CUDA_TRY(cudaMemcpyAsync(buf, vx, sizeof(double) * NA, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(buf + NA, vy, sizeof(double) * NA, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(buf + 2 * NA, vz, sizeof(double) * NA, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(buf + 3 * NA, mass, sizeof(double) * NA, cudaMemcpyHostToDevice, s->stream));
    double *T = s->out_f64.ensure<double>(s->n_rows);
    launch_atomic_temperature(*s, s->verlet.as<int>(), list_dist(*s), s->M, buf, buf + NA, buf + 2 * NA, buf + 3 * NA,
                              rc, T);
    d2h(*s, T_host, T, (size_t)s->n_rows);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_bond_analysis(mdb_system *s, double delta_r, double delta_theta, double rc, int nbins,
                             int *bond_length_host, int *bond_angle_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(bond_length_host && bond_angle_host && nbins > 0, MDB_ERR_VALUE, "both histograms and nbins are required");
    unsigned long long *hist = s->scratch2.ensure<unsigned long long>((size_t)2 * nbins);
    launch_bond_hist(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, delta_r, delta_theta, rc, nbins,
                     hist);
    std::vector<unsigned long long> h((size_t)2 * nbins);
    CUDA_TRY(cudaMemcpyAsync(h.data(), hist, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (int t = 0; t < nbins; ++t) {
        bond_length_host[t] += (int)h[t];
        bond_angle_host[t] += (int)h[nbins + t];
    }
    API_END
}

int mdb_system_adf(mdb_system *s, double delta_theta, const double *rc_list, const int *pair_list, int npair,
                   const int *types_host, int nbins, int *bond_angle_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(rc_list && pair_list && types_host && bond_angle_host && npair > 0 && nbins > 0, MDB_ERR_VALUE,
                "rc_list, pair_list, type_list and the histogram are required");
    const int *types = h2d(*s, s->types, types_host, (size_t)s->N);
    double *rcs = s->out_f64c.ensure<double>((size_t)4 * npair);
    int *pairs = s->perm_tmp.ensure<int>((size_t)3 * npair);
    CUDA_TRY(cudaMemcpyAsync(rcs, rc_list, sizeof(double) * 4 * npair, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(pairs, pair_list, sizeof(int) * 3 * npair, cudaMemcpyHostToDevice, s->stream));
    unsigned long long *hist = s->scratch2.ensure<unsigned long long>((size_t)npair * nbins);
    launch_adf_hist(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, delta_theta, rcs, pairs, npair,
                    types, nbins, hist);
    std::vector<unsigned long long> h((size_t)npair * nbins);
    CUDA_TRY(cudaMemcpyAsync(h.data(), hist, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (size_t t = 0; t < h.size(); ++t) bond_angle_host[t] += (int)h[t];
    API_END
}

// (n_rows, 8) = type, ordering, rmsd, interatomic distance, qw, qx, qy, qz; indices_host: (n_rows, 18).
int mdb_system_ptm(mdb_system *s, const char *structure, const int *types_host, double rmsd_threshold,
                   double *output_host, int *indices_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    const int R = s->n_rows;
    const int *types = nullptr;
    if (types_host) types = h2d(*s, s->types, types_host, (size_t)s->N);
    double *out = s->ptm_out.ensure<double>((size_t)R * 8);
    int *idx = s->ptm_idx.ensure<int>((size_t)R * 18);
    launch_ptm(*s, ptm_parse_flags(structure), s->verlet.as<int>(), s->M, types, rmsd_threshold, out, 8, idx, 18);
    d2h(*s, output_host, out, (size_t)R * 8);
    d2h(*s, indices_host, idx, (size_t)R * 18);
    if (output_host || indices_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// ---- Voronoi cells (voronoi.cu; src/voronoi.cpp:16-71, 307-447) ----------------------------------------------
int mdb_system_voronoi_volume(mdb_system *s, double *volume_host, int *neighbor_number_host, double *cavity_radius_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box, MDB_ERR_STATE, "no atoms uploaded");
    double *vol = s->out_f64.ensure<double>(s->N), *rad = s->out_f64b.ensure<double>(s->N);
    int *nf = s->out_i32.ensure<int>(s->N);
    launch_voronoi(*s, false, vol, nf, rad);
    d2h(*s, volume_host, vol, (size_t)s->N);
    d2h(*s, neighbor_number_host, nf, (size_t)s->N);
    d2h(*s, cavity_radius_host, rad, (size_t)s->N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_voronoi_neighbor(mdb_system *s, double a_face_area_threshold, double r_face_area_threshold, int *M)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box, MDB_ERR_STATE, "no atoms uploaded");
    MDB_REQUIRE(M, MDB_ERR_VALUE, "M is required");
    double *vol = s->out_f64.ensure<double>(s->N), *rad = s->out_f64b.ensure<double>(s->N);
    int *nf = s->vor_nn.ensure<int>(s->N);
    const int width = launch_voronoi(*s, true, vol, nf, rad);
    s->vor_M = width;   // voronoi.cpp:368-376: the row width is the largest face count (walls included)
    const size_t n = (size_t)s->N * (width ? width : 1);
    launch_voronoi_rows(*s, nf, width, a_face_area_threshold, r_face_area_threshold, s->vor_verlet.ensure<int>(n),
                        s->vor_dist.ensure<double>(n), s->vor_farea.ensure<double>(n));
    *M = width;
    API_END
}

int mdb_system_voronoi_fetch(mdb_system *s, int *verlet_host, double *distance_host, double *face_area_host,
                             int *neighbor_number_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->vor_M > 0 && s->vor_verlet.p, MDB_ERR_STATE, "no Voronoi neighbours on the device; build them first");
    const size_t n = (size_t)s->N * s->vor_M;
    d2h(*s, verlet_host, s->vor_verlet.as<int>(), n);
    d2h(*s, distance_host, s->vor_dist.as<double>(), n);
    d2h(*s, face_area_host, s->vor_farea.as<double>(), n);
    d2h(*s, neighbor_number_host, s->vor_nn.as<int>(), (size_t)s->N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_result_device(mdb_system *s, int **i32, double **f64)
{
    API_BEGIN
    if (i32) *i32 = s->out_i32.as<int>();
    if (f64) *f64 = s->out_f64.as<double>();
    API_END
}

int mdb_system_set_profiling(mdb_system *s, int on)
{
    API_BEGIN
    s->profile = on != 0;
    API_END
}

int mdb_system_last_times(mdb_system *s, float *t_binning_ms, float *t_neighbor_ms, float *t_cna_ms)
{
    API_BEGIN
    if (t_binning_ms) *t_binning_ms = s->t_bin;
    if (t_neighbor_ms) *t_neighbor_ms = s->t_neigh;
    if (t_cna_ms) *t_cna_ms = s->t_cna;
    API_END
}

// ---------------------------------------------------------------------------
// Section A: host-pointer drop-ins.  One transient system per call.
// ---------------------------------------------------------------------------
struct ScopedSystem {
    mdb_system *s{nullptr};
    ScopedSystem()
    {
        const int rc = mdb_system_create(0, &s);
        if (rc != MDB_OK) throw MdbError{rc};
    }
    ~ScopedSystem() { mdb_system_destroy(s); }
    MdbSystem &operator*() { return *s; }
    MdbSystem *operator->() { return s; }
};

// ---------------------------------------------------------------- CHILL+ and build_bond
int mdb_system_chill_plus(mdb_system *s, double rc, int *pattern_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(rc > 0, MDB_ERR_VALUE, "cutoff must be positive, got %g.", rc);
    int *pat = s->out_i32.ensure<int>(s->n_rows);
    launch_chill_plus(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, rc, pat);
    d2h(*s, pattern_host, pat, (size_t)s->n_rows);
    if (pattern_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_compute_chill_plus(const double *x, const double *y, const double *z, int N, const double *box9,
                           const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                           const int *nn, double rc, int *pattern, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nn, M, rc, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_chill_plus(s.s, rc, pattern);
    if (rcode != MDB_OK) return rcode;
    API_END
}

// bonds: *nbond rows of (i, j), i < j, in (i, list slot) order; two-call protocol: bonds_host == NULL returns the
// count only, a second call with a buffer of 2 * count ints receives the rows.
int mdb_system_build_bond(mdb_system *s, const int *types, const double *cutoff_matrix, int ntype, int *bonds_host,
                          int *nbond)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    require_list(*s);
    MDB_REQUIRE(types && cutoff_matrix && ntype > 0 && nbond, MDB_ERR_VALUE, "build_bond: types, cutoff matrix required");
    int *dt = h2d(*s, s->types, types, (size_t)s->N);
    double *dc = h2d(*s, s->out_f64b, cutoff_matrix, (size_t)ntype * ntype);
    int *db = nullptr;
    const int n = launch_build_bond(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), s->M, dt, dc, ntype, &db);
    *nbond = n;
    if (bonds_host && n > 0) {
        d2h(*s, bonds_host, db, (size_t)n * 2);
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    API_END
}

int mdb_build_bond(const int *verlet, int N, int M, const double *dist, const int *nn, const int *types,
                   const double *cutoff_matrix, int ntype, int *bonds, int *nbond, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(verlet && dist && nn && N > 0 && M > 0, MDB_ERR_VALUE, "build_bond: lists are required");
    ScopedSystem s;
    s->N = s->n_rows = N;   // list-only consumer: no coordinates needed
    h2d(*s, s->verlet, verlet, (size_t)N * M);
    h2d(*s, s->dist, dist, (size_t)N * M);
    h2d(*s, s->nn, nn, (size_t)N);
    s->M = M;
    s->list_kind = LIST_CUTOFF;
    s->has_dist = true;
    int rcode = mdb_system_build_bond(s.s, types, cutoff_matrix, ntype, bonds, nbond);
    if (rcode != MDB_OK) return rcode;
    API_END
}

// ---------------------------------------------------------------- FCC planar faults
int mdb_system_planar_faults(mdb_system *s, int identify_esf, int *fault_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->ptm_out.p && s->ptm_idx.p, MDB_ERR_STATE, "run polyhedral template matching first");
    MDB_REQUIRE(s->n_rows == s->N, MDB_ERR_STATE, "planar faults need the whole frame on one device");
    const int N = s->N;
    int *type = s->out_i32.ensure<int>((size_t)2 * N);
    int *fault = type + N;
    launch_types_from_ptm_output(*s, s->ptm_out.as<double>(), 8, N, type);
    launch_planar_faults(*s, type, N, s->ptm_idx.as<int>(), 18, 1, /*order: this library's template*/ 0,
                         identify_esf != 0, fault);
    d2h(*s, fault_host, fault, (size_t)N);
    if (fault_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_identify_sftb_fcc(const int *hcp_indices, int n_hcp, int *hcp_neighbors, const int *ptm_indices,
                          const int *structure_types, int N, int *fault_types, int identify_esf, int index_order,
                          int /*num_t*/)
{
    API_BEGIN
    (void)hcp_indices;
    (void)n_hcp;
    (void)hcp_neighbors;   // scratch of the reference; the device builds its own
    MDB_REQUIRE(ptm_indices && structure_types && fault_types && N > 0, MDB_ERR_VALUE, "identify_sftb_fcc: arrays required");
    MDB_REQUIRE(index_order == 0 || index_order == 1, MDB_ERR_VALUE, "index_order: 0 (mdapy_b200) or 1 (reference)");
    ScopedSystem s;
    s->N = s->n_rows = N;
    int *type = h2d(*s, s->types, structure_types, (size_t)N);
    int *idx = h2d(*s, s->ptm_idx, ptm_indices, (size_t)N * 12);
    int *fault = s->out_i32.ensure<int>((size_t)N);
    launch_planar_faults(*s, type, N, idx, 12, 0, index_order, identify_esf != 0, fault);
    d2h(*s, fault_types, fault, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

// ---------------------------------------------------------------- builders
int mdb_repeat_cell(double *new_pos, const double *old_box9, const double *old_pos, int n_old, int nx, int ny, int nz,
                    int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(new_pos && old_box9 && old_pos && n_old > 0 && nx > 0 && ny > 0 && nz > 0, MDB_ERR_VALUE,
                "repeat_cell: positions, box and positive replication counts are required");
    ScopedSystem s;
    DBox b{};
    for (int i = 0; i < 9; ++i) b.h[i] = old_box9[i];
    const size_t total = (size_t)n_old * nx * ny * nz;
    double *dold = h2d(*s, s->scratch, old_pos, (size_t)n_old * 3);
    double *dnew = s->out_f64.ensure<double>(total * 3);
    launch_repeat_cell(*s, dold, n_old, b, nx, ny, nz, nullptr, nullptr, nullptr, dnew);
    d2h(*s, new_pos, dnew, total * 3);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_set_atoms_lattice(mdb_system *s, const double *cell9, const double *basis_pos, int n_basis, int nx, int ny,
                                 int nz, const double *origin3, const int *boundary3)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(cell9 && basis_pos && n_basis > 0 && nx > 0 && ny > 0 && nz > 0, MDB_ERR_VALUE,
                "lattice: cell, basis and positive replication counts are required");
    const size_t total = (size_t)n_basis * nx * ny * nz;
    MDB_REQUIRE(total < ((size_t)1 << 31), MDB_ERR_VALUE, "lattice: more than 2^31 atoms");
    DBox cell{};
    double super9[9];
    for (int i = 0; i < 9; ++i) cell.h[i] = cell9[i];
    const int rep[3] = {nx, ny, nz};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) super9[r * 3 + c] = cell9[r * 3 + c] * rep[r];   // build_crystal: rows scaled
    set_box(*s, super9, origin3, boundary3);
    double *dbasis = h2d(*s, s->scratch, basis_pos, (size_t)n_basis * 3);
    double *dx = s->bx.ensure<double>(total), *dy = s->by.ensure<double>(total), *dz = s->bz.ensure<double>(total);
    launch_repeat_cell(*s, dbasis, n_basis, cell, nx, ny, nz, dx, dy, dz, nullptr);
    s->x = dx;
    s->y = dy;
    s->z = dz;
    s->N = s->n_rows = (int)total;
    s->gid = nullptr;
    s->slab_x0 = s->slab_nx = 0;
    s->local_frac = 1.0;
    invalidate(*s);
    API_END
}

int mdb_system_positions_device(mdb_system *s, double **dx, double **dy, double **dz, int *N)
{
    API_BEGIN
    MDB_REQUIRE(s->N > 0, MDB_ERR_STATE, "no atoms uploaded");
    if (dx) *dx = const_cast<double *>(s->x);
    if (dy) *dy = const_cast<double *>(s->y);
    if (dz) *dz = const_cast<double *>(s->z);
    if (N) *N = s->N;
    API_END
}

int mdb_system_fetch_positions(mdb_system *s, double *x, double *y, double *z)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0, MDB_ERR_STATE, "no atoms uploaded");
    d2h(*s, x, s->x, (size_t)s->N);
    d2h(*s, y, s->y, (size_t)s->N);
    d2h(*s, z, s->z, (size_t)s->N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_transform_and_filter(const double *x, const double *y, const double *z, int N, const double *rotation9,
                             const double *center3, const double *target3, const double *coeffs, int nfaces,
                             double *out_pos, int *count, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(x && y && z && rotation9 && center3 && target3 && (coeffs || nfaces == 0) && out_pos && count,
                MDB_ERR_VALUE, "transform_and_filter: missing argument");
    *count = 0;
    if (N <= 0) return MDB_OK;
    ScopedSystem s;
    double *dx = h2d(*s, s->bx, x, (size_t)N), *dy = h2d(*s, s->by, y, (size_t)N), *dz = h2d(*s, s->bz, z, (size_t)N);
    double *dpl = h2d(*s, s->out_f64b, coeffs, (size_t)nfaces * 4);
    double *dout = s->out_f64.ensure<double>((size_t)N * 3);
    const int n = launch_transform_and_filter(*s, dx, dy, dz, N, rotation9, center3, target3, dpl, nfaces, dout);
    d2h(*s, out_pos, dout, (size_t)n * 3);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    *count = n;
    API_END
}

int mdb_system_transform_and_filter(mdb_system *s, const double *rotation9, const double *center3,
                                    const double *target3, const double *coeffs, int nfaces, double *out_pos, int *count)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && rotation9 && center3 && target3 && out_pos && count, MDB_ERR_VALUE,
                "transform_and_filter: atoms on the device and all arguments are required");
    double *dpl = h2d(*s, s->out_f64b, coeffs, (size_t)nfaces * 4);
    double *dout = s->out_f64.ensure<double>((size_t)s->N * 3);
    const int n = launch_transform_and_filter(*s, s->x, s->y, s->z, s->N, rotation9, center3, target3, dpl, nfaces, dout);
    d2h(*s, out_pos, dout, (size_t)n * 3);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    *count = n;
    API_END
}

int mdb_filter_overlap_atom(const double *x, const double *y, const double *z, int N, const double *box9,
                            const double *origin3, const int *boundary3, double rc, unsigned char *keep, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(rc > 0 && keep, MDB_ERR_VALUE, "filter_overlap_atom: rc must be positive");
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    unsigned char *dk = reinterpret_cast<unsigned char *>(s->out_i32.ensure<int>((size_t)N / 4 + 1));
    launch_filter_overlap(*s, rc, dk);
    d2h(*s, keep, dk, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_system_filter_overlap(mdb_system *s, double rc, unsigned char *keep_host)
{
    API_BEGIN
    CUDA_TRY(cudaSetDevice(s->device));
    MDB_REQUIRE(s->N > 0 && s->has_box && rc > 0, MDB_ERR_STATE, "no atoms uploaded / rc must be positive");
    unsigned char *dk = reinterpret_cast<unsigned char *>(s->out_i32.ensure<int>((size_t)s->N / 4 + 1));
    launch_filter_overlap(*s, rc, dk);
    d2h(*s, keep_host, dk, (size_t)s->N);
    if (keep_host) CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_build_neighbor(const double *x, const double *y, const double *z, int N, const double *box9,
                       const double *origin3, const int *boundary3, double rc, int *verlet, double *dist,
                       int *nn, int M, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(M > 0, MDB_ERR_VALUE, "max_neigh must be positive, got %d.", M);
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    build_neighbor(*s, rc, M);
    d2h(*s, verlet, s->verlet.as<int>(), (size_t)N * M);
    d2h(*s, dist, list_dist(*s), (size_t)N * M);
    d2h(*s, nn, s->nn.as<int>(), (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_build_neighbor_without_max_neigh(const double *x, const double *y, const double *z, int N,
                                         const double *box9, const double *origin3, const int *boundary3,
                                         double rc, int /*num_t*/, void **handle, int *M)
{
    API_BEGIN
    MDB_REQUIRE(handle && M, MDB_ERR_VALUE, "handle and M are required");
    mdb_system *s = nullptr;
    int rcode = mdb_system_create(0, &s);
    if (rcode != MDB_OK) return rcode;
    try {
        set_box(*s, box9, origin3, boundary3);
        upload_atoms(*s, x, y, z, N);
        build_neighbor(*s, rc, 0);
    } catch (...) {
        mdb_system_destroy(s);
        throw;
    }
    *handle = s;
    *M = s->M;
    API_END
}

int mdb_neighbor_auto_fetch(void *handle, int *verlet, double *dist, int *nn)
{
    mdb_system *s = static_cast<mdb_system *>(handle);
    if (!s) {
        mdb_set_error("NULL handle");
        return MDB_ERR_VALUE;
    }
    const int rc = mdb_system_fetch_neighbor(s, verlet, dist, nn);
    mdb_system_destroy(s);
    return rc;
}

int mdb_knn(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
            const int *boundary3, int k, int *indices, double *distances, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    launch_knn(*s, k);
    d2h(*s, indices, s->verlet.as<int>(), (size_t)N * k);
    d2h(*s, distances, list_dist(*s), (size_t)N * k);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_sort_verlet_by_distance(int *verlet, double *dist, int N, int M, int k, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(N > 0 && M > 0 && verlet && dist, MDB_ERR_VALUE, "verlet/dist arrays required");
    ScopedSystem s;
    const size_t n = (size_t)N * M;
    int *dv = h2d(*s, s->verlet, verlet, n);
    double *dd = h2d(*s, s->dist, dist, n);
    launch_sort_rows(*s, dv, dd, N, M, k);
    d2h(*s, verlet, dv, n);
    d2h(*s, dist, dd, n);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_fcna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
             const int *boundary3, const int *verlet, int M, const int *nn, int *pattern, double rc, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    // a caller-supplied list carries no trustworthy distance bound (rc = -1): exact bond tests
    int rcode = mdb_system_put_neighbor(s.s, verlet, nullptr, nn, M, -1.0, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_fcna(s.s, rc, pattern);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_acna(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
             const int *boundary3, const int *verlet, int M, int *pattern, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, nullptr, nullptr, M, -1.0, LIST_KNN);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_acna(s.s, pattern);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_ids(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
            const int *boundary3, const int *verlet, int M, int *new_verlet, int *pattern, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, nullptr, nullptr, M, -1.0, LIST_KNN);
    if (rcode != MDB_OK) return rcode;
    int *pat = s->out_i32.ensure<int>(N);
    CUDA_TRY(cudaMemsetAsync(pat, 0, sizeof(int) * N, s->stream));
    int *second = new_verlet ? s->scratch2.ensure<int>((size_t)N * 12) : nullptr;
    launch_ids(*s, s->verlet.as<int>(), M, second, pat);
    d2h(*s, pattern, pat, (size_t)N);
    d2h(*s, new_verlet, second, (size_t)N * 12);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_get_csp(const double *x, const double *y, const double *z, int N, const double *box9,
                const double *origin3, const int *boundary3, const int *verlet, int M, int nnei, double *csp,
                int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, nullptr, nullptr, M, -1.0, LIST_KNN);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_csp(s.s, nnei, csp);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_compute_cnp(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    const int *nn, double *cnp, double rc, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nn, M, rc, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_cnp(s.s, rc, cnp);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_get_voronoi_volume_number_radius(const double *x, const double *y, const double *z, int N, const double *box9,
                                         const double *origin3, const int *boundary3, double *volume,
                                         int *neighbor_number, double *cavity_radius, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    int rcode = mdb_system_set_atoms(s.s, x, y, z, N, box9, origin3, boundary3);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_voronoi_volume(s.s, volume, neighbor_number, cavity_radius);
    if (rcode != MDB_OK) return rcode;
    API_END
}

// no coordinates on this path: the list alone (a placeholder frame carries the row count)
static int list_only_system(MdbSystem &s, int N, const int *verlet, const double *dist, const int *nn, int M, double rc)
{
    MDB_REQUIRE(N > 0 && verlet && nn, MDB_ERR_VALUE, "verlet_list and neighbor_number are required");
    s.N = s.n_rows = N;
    s.x = s.y = s.z = nullptr;
    return mdb_system_put_neighbor(&s, verlet, dist, nn, M, rc, LIST_CUTOFF);
}

int mdb_get_wcp(const int *verlet, int N, int M, const int *nn, const int *type_list, int Ntype, double *WCP,
                int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    int rcode = list_only_system(*s, N, verlet, nullptr, nn, M, -1.0);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_wcp(s.s, type_list, Ntype, WCP);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_average_by_neighbor(double rc, const int *verlet, int N, int M, const double *dist, const int *nn,
                            const double *value, double *value_ave, int include_self, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(dist && value_ave, MDB_ERR_VALUE, "distance_list and value_ave are required");
    ScopedSystem s;
    int rcode = list_only_system(*s, N, verlet, dist, nn, M, rc);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_average_by_neighbor(s.s, rc, value, include_self, value_ave);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_get_cluster(const int *verlet, int N, int M, const double *dist, const int *nn, double rc,
                    int *particle_clusters, int *cluster_number)
{
    API_BEGIN
    MDB_REQUIRE(dist && particle_clusters, MDB_ERR_VALUE, "distance_list and particleClusters are required");
    ScopedSystem s;
    int rcode = list_only_system(*s, N, verlet, dist, nn, M, rc);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_cluster(s.s, rc, nullptr, nullptr, nullptr, nullptr, 0, particle_clusters, cluster_number);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_get_cluster_by_bond(const int *verlet, int N, int M, const int *nn, int *particle_clusters,
                            int *cluster_number)
{
    API_BEGIN
    MDB_REQUIRE(particle_clusters, MDB_ERR_VALUE, "particleClusters is required");
    ScopedSystem s;
    int rcode = list_only_system(*s, N, verlet, nullptr, nn, M, -1.0);
    if (rcode != MDB_OK) return rcode;
    int *out = s->out_i32.ensure<int>(N);
    const int count = launch_cluster(*s, s->verlet.as<int>(), nullptr, s->nn.as<int>(), M, 0.0, out);
    if (cluster_number) *cluster_number = count;
    d2h(*s, particle_clusters, out, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_filter_by_type(int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                       const int *type1, const int *type2, const double *r, int npair, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(dist && type_list && type1 && type2 && r && npair > 0, MDB_ERR_VALUE,
                "distance_list, type_list and the type-pair table are required");
    ScopedSystem s;
    int rcode = list_only_system(*s, N, verlet, dist, nn, M, -1.0);
    if (rcode != MDB_OK) return rcode;
    const int *types = h2d(*s, s->types, type_list, (size_t)N);
    int *pairs = s->scratch2.ensure<int>((size_t)2 * npair + 2);
    double *rr = s->out_f64c.ensure<double>(npair);
    CUDA_TRY(cudaMemcpyAsync(pairs, type1, sizeof(int) * npair, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(pairs + npair, type2, sizeof(int) * npair, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(rr, r, sizeof(double) * npair, cudaMemcpyHostToDevice, s->stream));
    launch_filter_by_type(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), M, types, pairs, pairs + npair,
                          rr, npair);
    d2h(*s, verlet, s->verlet.as<int>(), (size_t)N * M);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_calculate_structure_entropy(double rc, double sigma, int use_local_density, double volume, const double *dist,
                                    int N, int M, const int *nn, double *entropy, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(dist && nn && entropy && N > 0 && M > 0, MDB_ERR_VALUE, "distance_list, neighbor_number and entropy are required");
    ScopedSystem s;
    s->N = s->n_rows = N;
    const double *dd = h2d(*s, s->dist, dist, (size_t)N * M);
    const int *dn = h2d(*s, s->nn, nn, (size_t)N);
    double *ent = s->out_f64.ensure<double>(N);
    launch_structure_entropy(*s, dd, dn, M, rc, sigma, use_local_density != 0, volume, ent);
    d2h(*s, entropy, ent, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_compute_temp(const int *verlet, int N, int M, const double *dist, const double *vx, const double *vy,
                     const double *vz, const double *mass, double *T, double rc, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(dist && T, MDB_ERR_VALUE, "distance_list and T are required");
    ScopedSystem s;
    s->N = s->n_rows = N;
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nullptr, M, rc, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_atomic_temperature(s.s, vx, vy, vz, mass, rc, T);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_compute_bond(const double *x, const double *y, const double *z, int N, const double *box9,
                     const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                     const int *nn, int *bond_length_distribution, int *bond_angle_distribution, double delta_r,
                     double delta_theta, double rc, int nbins, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nn, M, rc, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_bond_analysis(s.s, delta_r, delta_theta, rc, nbins, bond_length_distribution, bond_angle_distribution);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_compute_adf(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    const int *nn, double delta_theta, const double *rc_list, const int *pair_list, int npair,
                    const int *type_list, int nbins, int *bond_angle_distribution, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nn, M, -1.0, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_adf(s.s, delta_theta, rc_list, pair_list, npair, type_list, nbins, bond_angle_distribution);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_wrap_positions(double *x, double *y, double *z, int N, const double *box9, const double *origin3,
                       const int *boundary3, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    double *dx = s->bx.as<double>(), *dy = s->by.as<double>(), *dz = s->bz.as<double>();
    launch_wrap_positions(*s, dx, dy, dz, N);
    d2h(*s, x, dx, (size_t)N);
    d2h(*s, y, dy, (size_t)N);
    d2h(*s, z, dz, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_compute_aja(const double *x, const double *y, const double *z, int N, const double *box9,
                    const double *origin3, const int *boundary3, const int *verlet, int M, const double *dist,
                    int Md, int *aja, int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(M == Md, MDB_ERR_VALUE, "verlet_list and distance_list must have the same row width (%d vs %d)", M, Md);
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nullptr, M, -1.0, LIST_KNN);
    if (rcode != MDB_OK) return rcode;
    rcode = mdb_system_aja(s.s, aja);
    if (rcode != MDB_OK) return rcode;
    API_END
}

int mdb_get_ptm(const char *structure, const double *x, const double *y, const double *z, int N, const double *box9,
                const double *origin3, const int *boundary3, const int *verlet, int M, const int *atom_types,
                int ntypes, double rmsd_threshold, double *output, int ocols, int *ptm_indices, int icols,
                int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(output && ocols >= 1, MDB_ERR_VALUE, "output array required");
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, nullptr, nullptr, M, -1.0, LIST_KNN);
    if (rcode != MDB_OK) return rcode;
    // atom types are used only when one per atom is given (polyhedral_template_matching.cpp:160-162)
    const int *types = (atom_types && ntypes == N) ? h2d(*s, s->types, atom_types, (size_t)N) : nullptr;
    double *out = s->ptm_out.ensure<double>((size_t)N * ocols);
    int *idx = ptm_indices ? s->ptm_idx.ensure<int>((size_t)N * icols) : nullptr;
    launch_ptm(*s, ptm_parse_flags(structure), s->verlet.as<int>(), M, types, rmsd_threshold, out, ocols, idx, icols);
    d2h(*s, output, out, (size_t)N * ocols);
    if (ptm_indices) d2h(*s, ptm_indices, idx, (size_t)N * icols);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_get_sq(const double *x, const double *y, const double *z, int N, const double *box9, const double *origin3,
               const int *boundary3, const int *verlet, int M, const double *dist, const int *nn,
               const double *weight, const int *llist, int ndeg, int nnn, int lmax, int wl, int wlhat, int average,
               int use_voronoi, double rc, int use_weight, double *qlm_r, double *qlm_i, double *qnarray, int ncol,
               int /*num_t*/)
{
    API_BEGIN
    MDB_REQUIRE(qlm_r && qlm_i && qnarray, MDB_ERR_VALUE, "output arrays are required");
    MDB_REQUIRE(ncol == ndeg * (1 + (wl ? 1 : 0) + (wlhat ? 1 : 0)), MDB_ERR_VALUE, "qnarray has %d columns", ncol);
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    int rcode = mdb_system_put_neighbor(s.s, verlet, dist, nn, M, rc, LIST_CUTOFF);
    if (rcode != MDB_OK) return rcode;
    const int nz = 2 * lmax + 1;
    const size_t nq = (size_t)N * ndeg * nz;
    // inout semantics of the reference: start from the caller's arrays (zeros in the wrapper)
    double *qr = h2d(*s, s->qlm_r, qlm_r, nq), *qi = h2d(*s, s->qlm_i, qlm_i, nq);
    double *qn = h2d(*s, s->qn, qnarray, (size_t)N * ncol);
    const double *w = use_weight ? h2d(*s, s->weight, weight, (size_t)N * M) : nullptr;
    launch_steinhardt(*s, s->verlet.as<int>(), list_dist(*s), s->nn.as<int>(), M, w, llist, ndeg, nnn, lmax,
                      wl != 0, wlhat != 0, average != 0, use_voronoi != 0, rc, use_weight != 0, qr, qi, qn);
    d2h(*s, qlm_r, qr, nq);
    d2h(*s, qlm_i, qi, nq);
    d2h(*s, qnarray, qn, (size_t)N * ncol);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_identify_solid_liquid(int Q6index, const double *Q6, const int *verlet, int N, int M, const double *dist,
                              const int *nn, const double *qlm_r, const double *qlm_i, int ndeg, int nz,
                              double threshold, int n_bond, int *solidliquid, int *nbond, int use_voronoi, int nnn,
                              double rc, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    s->N = s->n_rows = N;
    const size_t nq = (size_t)N * ndeg * nz;
    int *dv = h2d(*s, s->verlet, verlet, (size_t)N * M);
    double *dd = h2d(*s, s->dist, dist, (size_t)N * M);
    int *dn = h2d(*s, s->nn, nn, (size_t)N);
    double *qr = h2d(*s, s->qlm_r, qlm_r, nq), *qi = h2d(*s, s->qlm_i, qlm_i, nq);
    double *q6 = h2d(*s, s->out_f64, Q6, (size_t)N);
    int *solid = h2d(*s, s->out_i32, solidliquid, (size_t)N);  // only 1s are written: keep the caller's zeros
    int *nb = s->scratch2.ensure<int>(N);
    launch_solid_liquid(*s, dv, dd, dn, M, Q6index, q6, qr, qi, ndeg, nz, threshold, n_bond, use_voronoi != 0, nnn, rc,
                        solid, nb);
    d2h(*s, solidliquid, solid, (size_t)N);
    d2h(*s, nbond, nb, (size_t)N);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

static int rdf_list_host(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list,
                         double *g, int ntype, double rc, int nbin)
{
    API_BEGIN
    MDB_REQUIRE(N > 0 && M > 0 && verlet && dist && nn && g, MDB_ERR_VALUE, "lists and g are required");
    ScopedSystem s;
    s->N = s->n_rows = N;
    int *dv = h2d(*s, s->verlet, verlet, (size_t)N * M);
    double *dd = h2d(*s, s->dist, dist, (size_t)N * M);
    int *dn = h2d(*s, s->nn, nn, (size_t)N);
    const int *dt = type_list ? h2d(*s, s->types, type_list, (size_t)N) : nullptr;
    const int nslot = (type_list ? ntype * ntype : 1) * nbin;
    double *dg = h2d(*s, s->out_f64b, g, (size_t)nslot);
    launch_rdf_list(*s, dv, dd, dn, N, M, dt, ntype, rc, nbin, dg);
    d2h(*s, g, dg, (size_t)nslot);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    API_END
}

int mdb_rdf(const int *verlet, int N, int M, const double *dist, const int *nn, const int *type_list, double *g,
            int ntype, double rc, int nbin)
{
    if (!type_list) {
        mdb_set_error("type_list is required");
        return MDB_ERR_VALUE;
    }
    return rdf_list_host(verlet, N, M, dist, nn, type_list, g, ntype, rc, nbin);
}

int mdb_rdf_single_species(const int *verlet, int N, int M, const double *dist, const int *nn, double *g, double rc,
                           int nbin)
{
    return rdf_list_host(verlet, N, M, dist, nn, nullptr, g, 1, rc, nbin);
}

int mdb_rdf_streaming(const double *x, const double *y, const double *z, int N, const int *type_list,
                      const double *box9, const double *origin3, const int *boundary3, double *g, int ntype,
                      double rc, int nbin, int /*num_t*/)
{
    API_BEGIN
    ScopedSystem s;
    set_box(*s, box9, origin3, boundary3);
    upload_atoms(*s, x, y, z, N);
    const int rcode = mdb_system_rdf(s.s, type_list, ntype, rc, nbin, 1, g);
    if (rcode != MDB_OK) return rcode;
    API_END
}

}  // extern "C"
