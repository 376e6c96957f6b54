// mdapy_b200/csrc/staging.cu -- host -> device upload of the caller's coordinate columns.
//
// The reference reads x, y, z straight from NumPy memory (src/neighbor.cpp:189-205 takes nanobind views); here
// they cross PCIe first, and NumPy memory is PAGEABLE: cudaMemcpyAsync then stages through one driver thread at
// ~10 GB/s, five times slower than the link.  mdb_h2d keeps the link busy instead: worker threads copy 4 MiB
// slices into page-locked ring buffers (two per worker) and queue the DMA from there, so the host-side copy of
// slice k+1 overlaps the transfer of slice k and several cores share the host copy.  Page-locked sources
// (mdb_host_alloc, torch pin_memory) skip the staging and go down as one asynchronous copy per column.
#include "internal.cuh"

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace {
constexpr size_t SLICE = (size_t)4 << 20;
constexpr int MAX_WORKERS = 8;

struct Stager {
    int device{0};
    std::mutex mu;   // one staged upload per device at a time
    cudaStream_t st[MAX_WORKERS]{};
    cudaEvent_t ev[MAX_WORKERS][2]{};
    cudaEvent_t done[MAX_WORKERS]{};
    cudaEvent_t start{};
    void *buf[MAX_WORKERS][2]{};
    int ready{0};    // workers with streams / buffers created
};

std::mutex g_stagers_mu;
std::map<int, Stager *> g_stagers;

Stager &stager_for(int device)
{
    std::lock_guard<std::mutex> lk(g_stagers_mu);
    Stager *&p = g_stagers[device];
    if (!p) {
        p = new Stager();
        p->device = device;
    }
    return *p;
}

bool is_pagelocked(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct Slice {
    char *dst;
    const char *src;
    size_t len;
};
}  // namespace

int mdb_upload_threads()
{
    if (const char *e = getenv("MDB_UPLOAD_THREADS")) {
        const int v = atoi(e);
        if (v >= 1) return std::min(v, MAX_WORKERS);
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(MAX_WORKERS, std::max(2u, hw ? hw / 2 : 2u));
}

// n columns: dst[k] (device) <- src[k] (host), bytes[k]; the copies are ordered after what `consumer` holds at the
// call and `consumer` waits for them before anything queued later.  On return the SOURCE arrays of pageable columns
// are no longer needed; page-locked ones must stay valid until the stream reaches the copies.
void mdb_h2d(int n, void *const *dst, const void *const *src, const size_t *bytes, cudaStream_t consumer, int threads)
{
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    std::vector<Slice> jobs;
    for (int k = 0; k < n; ++k) {
        if (!bytes[k]) continue;
        if (bytes[k] < ((size_t)1 << 20) || is_pagelocked(src[k])) {
            CUDA_TRY(cudaMemcpyAsync(dst[k], src[k], bytes[k], cudaMemcpyHostToDevice, consumer));
            continue;
        }
        for (size_t off = 0; off < bytes[k]; off += SLICE)
            jobs.push_back({(char *)dst[k] + off, (const char *)src[k] + off, std::min(SLICE, bytes[k] - off)});
    }
    if (jobs.empty()) return;
    Stager &S = stager_for(device);
    std::lock_guard<std::mutex> lk(S.mu);
    const int T = std::max(1, std::min({threads, MAX_WORKERS, (int)jobs.size()}));
    if (!S.start) CUDA_TRY(cudaEventCreateWithFlags(&S.start, cudaEventDisableTiming));
    for (; S.ready < T; ++S.ready) {
        const int w = S.ready;
        CUDA_TRY(cudaStreamCreateWithFlags(&S.st[w], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&S.done[w], cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&S.ev[w][b], cudaEventDisableTiming | cudaEventBlockingSync));
            MDB_REQUIRE(mdb_host_alloc(SLICE, &S.buf[w][b]) == MDB_OK, MDB_ERR_CUDA, "%s", mdb_last_error());
        }
    }
    CUDA_TRY(cudaEventRecord(S.start, consumer));
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    char errbuf[512] = "";
    std::mutex err_mu;
    auto work = [&](int w) {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(S.st[w], S.start, 0);
        int it = 0;
        while (e == cudaSuccess && !failed.load(std::memory_order_relaxed)) {
            const size_t j = next.fetch_add(1);
            if (j >= jobs.size()) break;
            const int b = it++ & 1;
            e = cudaEventSynchronize(S.ev[w][b]);          // the DMA that last read this buffer (this or an earlier call)
            if (e != cudaSuccess) break;
            memcpy(S.buf[w][b], jobs[j].src, jobs[j].len);
            e = cudaMemcpyAsync(jobs[j].dst, S.buf[w][b], jobs[j].len, cudaMemcpyHostToDevice, S.st[w]);
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[w][b], S.st[w]);
        }
        if (e == cudaSuccess) e = cudaEventRecord(S.done[w], S.st[w]);
        if (e != cudaSuccess) {
            failed.store(1);
            std::lock_guard<std::mutex> g(err_mu);
            snprintf(errbuf, sizeof errbuf, "staged upload: %s", cudaGetErrorString(e));
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < T; ++w) th.emplace_back(work, w);
    work(0);
    for (auto &t : th) t.join();
    MDB_REQUIRE(!failed.load(), MDB_ERR_CUDA, "%s", errbuf);
    for (int w = 0; w < T; ++w) CUDA_TRY(cudaStreamWaitEvent(consumer, S.done[w], 0));
}
