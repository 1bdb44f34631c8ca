// capi.cu -- the extern "C" boundary declared in include/pgeof_b200.h.
// Host flavours stage numpy buffers through stream-ordered device scratch; device flavours
// run directly on the caller's buffers and stream.  No entry point computes on the CPU.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <chrono>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace pgeof {

thread_local uint64_t g_launches = 0;
static thread_local char g_error[512] = "";

int sm_count()
{
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
    if (!cached[dev]) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
        cached[dev] = sms;
    }
    return cached[dev];
}

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------
// optional per-kernel event timing
// ---------------------------------------------------------------------------
static const char* const kTimerNames[] = {"knn_search", "radius_search", "features", "multiscale", "optimal", "selected", "grid_build", "row_order"};
constexpr int kNumTimers = sizeof(kTimerNames) / sizeof(kTimerNames[0]);
static bool g_profile = false;
static std::mutex g_profile_mutex;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events[kNumTimers];

KernelTimer::KernelTimer(const char* name, cudaStream_t s) : stream(s)
{
    if (!g_profile) return;
    for (int i = 0; i < kNumTimers; ++i) if (std::strcmp(name, kTimerNames[i]) == 0) slot = i;
    if (slot < 0) return;
    if (cudaEventCreate(&start) != cudaSuccess) { slot = -1; cudaGetLastError(); return; }
    cudaEventRecord(start, stream);
}

KernelTimer::~KernelTimer()
{
    if (slot < 0) return;
    cudaEvent_t stop;
    if (cudaEventCreate(&stop) != cudaSuccess) { cudaGetLastError(); cudaEventDestroy(start); return; }
    cudaEventRecord(stop, stream);
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_profile_events[slot].emplace_back(start, stop);
}

// ---------------------------------------------------------------------------
// device scratch: caching arena (see DeviceBuffer in common.cuh)
// ---------------------------------------------------------------------------
struct ArenaBlock { void* ptr; size_t bytes; cudaStream_t stream; cudaEvent_t event; uint64_t seq; };
struct DeviceArena {
    std::mutex m;
    std::multimap<size_t, ArenaBlock> free_blocks[64];
    size_t cached_bytes[64] = {};
    size_t budget[64] = {};          // cap on cached (idle) bytes per device, 0 = not initialised yet
    uint64_t seq = 0;
    static size_t round(size_t b) { const size_t g = b < ((size_t)1 << 20) ? 512 : ((size_t)2 << 20); return (b + g - 1) / g * g; }
    void trim_locked(int dev)
    {
        for (auto& kv : free_blocks[dev]) { cudaFree(kv.second.ptr); if (kv.second.event) cudaEventDestroy(kv.second.event); }
        free_blocks[dev].clear();
        cached_bytes[dev] = 0;
    }
    // Idle blocks are kept for reuse up to a budget: PGEOF_ARENA_MAX_MB, default a quarter of the device's memory.  Beyond it
    // the least recently released blocks go back to the driver (cudaFree synchronises; steady-state callers never get here),
    // so a co-resident allocator (torch) is not starved by scratch of calls long past.
    size_t budget_of(int dev)
    {
        if (!budget[dev]) {
            size_t b = 0;
            if (const char* e = std::getenv("PGEOF_ARENA_MAX_MB")) b = (size_t)std::strtoull(e, nullptr, 10) << 20;
            else {
                size_t free_b = 0, total_b = 0;
                if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) b = total_b / 4; else cudaGetLastError();
            }
            budget[dev] = std::max<size_t>(b, (size_t)64 << 20);
        }
        return budget[dev];
    }
    void enforce_budget_locked(int dev)
    {
        const size_t cap = budget_of(dev);
        while (cached_bytes[dev] > cap && !free_blocks[dev].empty()) {
            auto oldest = free_blocks[dev].begin();
            for (auto it = free_blocks[dev].begin(); it != free_blocks[dev].end(); ++it) if (it->second.seq < oldest->second.seq) oldest = it;
            cudaFree(oldest->second.ptr);
            if (oldest->second.event) cudaEventDestroy(oldest->second.event);
            cached_bytes[dev] -= oldest->second.bytes;
            free_blocks[dev].erase(oldest);
        }
    }
};
static DeviceArena& arena() { static DeviceArena* a = new DeviceArena(); return *a; }   // leaked on purpose (exit order)

int DeviceBuffer::alloc(size_t want_bytes, cudaStream_t s)
{
    release();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); set_error("no current CUDA device"); return PGEOF_ECUDA; }
    const size_t want = DeviceArena::round(want_bytes ? want_bytes : 16);
    DeviceArena& ar = arena();
    {
        std::lock_guard<std::mutex> lock(ar.m);
        auto& fl = ar.free_blocks[dev];
        auto it = fl.lower_bound(want);
        if (it != fl.end() && it->first <= want + want / 4 + ((size_t)1 << 20)) {
            const ArenaBlock blk = it->second;
            fl.erase(it);
            ar.cached_bytes[dev] -= blk.bytes;
            // stream-ordered reuse: work queued on another stream before the release must be over first
            if (blk.stream != s && blk.event && cudaStreamWaitEvent(s, blk.event, 0) != cudaSuccess) cudaGetLastError();
            ptr = blk.ptr; bytes = blk.bytes; event = blk.event; stream = s; device = dev;
            return PGEOF_OK;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {                       // give the cached blocks back and try once more
        cudaGetLastError();
        cudaDeviceSynchronize();
        { std::lock_guard<std::mutex> lock(ar.m); ar.trim_locked(dev); }
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? PGEOF_ENOMEM : PGEOF_ECUDA;
    }
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
    ptr = p; bytes = want; event = ev; stream = s; device = dev;
    return PGEOF_OK;
}

void DeviceBuffer::release()
{
    if (!ptr) return;
    cudaEvent_t ev = (cudaEvent_t)event;
    bool recorded = ev && cudaEventRecord(ev, stream) == cudaSuccess;
    if (!recorded) { cudaGetLastError(); cudaStreamSynchronize(stream); }   // no event: make the block safe for any stream
    DeviceArena& ar = arena();
    {
        std::lock_guard<std::mutex> lock(ar.m);
        ar.free_blocks[device].emplace(bytes, ArenaBlock{ptr, bytes, recorded ? stream : nullptr, ev, ++ar.seq});
        ar.cached_bytes[device] += bytes;
        int cur = -1;
        if (ar.cached_bytes[device] > ar.budget_of(device) && cudaGetDevice(&cur) == cudaSuccess) {
            if (cur != device) cudaSetDevice(device);
            ar.enforce_budget_locked(device);
            if (cur != device) cudaSetDevice(cur);
        }
    }
    ptr = nullptr; bytes = 0; event = nullptr;
}

// ---------------------------------------------------------------------------
// per-thread context: device check + the stream host entry points run on
// ---------------------------------------------------------------------------
struct HostCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy = nullptr;      // second stream of the chunked host pipelines (uploads run beside compute + downloads)
};
static thread_local HostCtx g_ctx;

static int ensure_device(int* device_out)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); pgeof_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        return PGEOF_ECUDA;
    }
    int dev = 0;
    PGEOF_CUDA(cudaGetDevice(&dev));
    if (device_out) *device_out = dev;
    return PGEOF_OK;
}

static int host_stream(cudaStream_t* s)
{
    int dev = 0;
    PGEOF_TRY(ensure_device(&dev));
    if (g_ctx.device != dev || !g_ctx.stream) {
        if (g_ctx.stream) {
            cudaSetDevice(g_ctx.device); cudaStreamDestroy(g_ctx.stream);
            if (g_ctx.copy) cudaStreamDestroy(g_ctx.copy);
            cudaSetDevice(dev); g_ctx.stream = nullptr; g_ctx.copy = nullptr;
        }
        PGEOF_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
        PGEOF_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy, cudaStreamNonBlocking));
        g_ctx.device = dev;
    }
    *s = g_ctx.stream;
    return PGEOF_OK;
}

// Switches to the device that owns `p` for the lifetime of the guard.
struct DeviceGuard {
    int prev = -1;
    int enter(const void* p)
    {
        int dev = 0;
        PGEOF_TRY(ensure_device(&dev));
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, p);
        if (e != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
            cudaGetLastError();
            set_error("expected a CUDA device pointer");
            return PGEOF_EINVAL;
        }
        if (attr.device != dev) { prev = dev; PGEOF_CUDA(cudaSetDevice(attr.device)); }
        return PGEOF_OK;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------
// pinned host pool (backs numpy result arrays created by the binding)
// ---------------------------------------------------------------------------
struct PinnedPool {
    std::mutex m;
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;
    static size_t round(size_t b) { const size_t g = b < ((size_t)1 << 20) ? 4096 : ((size_t)2 << 20); return (b + g - 1) / g * g; }
    void* get(size_t bytes)
    {
        const size_t want = round(bytes ? bytes : 1);
        std::lock_guard<std::mutex> lock(m);
        auto it = free_blocks.lower_bound(want);
        if (it != free_blocks.end() && it->first <= want + want / 4 + 4096) {
            void* p = it->second; live[p] = it->first; free_blocks.erase(it); return p;
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            trim_locked();
            if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        live[p] = want;
        return p;
    }
    void put(void* p)
    {
        std::lock_guard<std::mutex> lock(m);
        auto it = live.find(p);
        if (it == live.end()) return;
        free_blocks.emplace(it->second, p);
        live.erase(it);
    }
    void trim_locked()
    {
        for (auto& kv : free_blocks) cudaFreeHost(kv.second);
        free_blocks.clear();
    }
};
static PinnedPool& pinned_pool() { static PinnedPool* p = new PinnedPool(); return *p; }   // leaked on purpose (exit order)

// PGEOF_HOST_TRACE=1: wall-clock of the phases of a host-buffer call (synchronises between phases; diagnostics only)
struct HostTrace {
    bool on;
    cudaStream_t s;
    const char* what;
    std::chrono::steady_clock::time_point t0, t;
    std::string line;
    HostTrace(const char* w, cudaStream_t st) : on(std::getenv("PGEOF_HOST_TRACE") != nullptr), s(st), what(w)
    {
        if (on) t0 = t = std::chrono::steady_clock::now();
    }
    void mark(const char* phase)
    {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto now = std::chrono::steady_clock::now();
        char buf[64];
        std::snprintf(buf, sizeof(buf), " %s %.1f", phase, std::chrono::duration<double, std::milli>(now - t).count());
        line += buf;
        t = now;
    }
    ~HostTrace()
    {
        if (on) std::fprintf(stderr, "[pgeof host] %s:%s | total %.1f ms\n", what, line.c_str(),
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// host <-> device staging helpers
static int h2d(DeviceBuffer* d, const void* h, size_t bytes, cudaStream_t s)
{
    PGEOF_TRY(d->alloc(bytes, s));
    if (bytes) PGEOF_CUDA(cudaMemcpyAsync(d->ptr, h, bytes, cudaMemcpyHostToDevice, s));
    return PGEOF_OK;
}
static int d2h(void* h, const DeviceBuffer& d, size_t bytes, cudaStream_t s)
{
    if (bytes) PGEOF_CUDA(cudaMemcpyAsync(h, d.ptr, bytes, cudaMemcpyDeviceToHost, s));
    return PGEOF_OK;
}

static int check_scales(const uint32_t* k_scales, size_t n)
{
    uint32_t prev = 1;                                   // pgeof.hpp:123-132
    for (size_t i = 0; i < n; ++i) { if (k_scales[i] < prev) return 0; prev = k_scales[i]; }
    return 1;
}

__global__ void iota_scale_kernel(uint32_t* out, size_t n, uint32_t k)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(i * k);
}

}  // namespace pgeof

using namespace pgeof;

// parked state of a radius_search_csr call pair (see pgeof_radius_search_csr_dev / pgeof_radius_search_csr)
struct PendingCsr {
    const float* data = nullptr; const float* query = nullptr;
    size_t n_data = 0, n_query = 0;
    float radius = 0.f; uint32_t max_knn = 0;
    const uint32_t* nn_ptr = nullptr;
    int device = -1;
    DeviceBuffer idx;
    void clear() { idx.release(); data = query = nullptr; nn_ptr = nullptr; n_data = n_query = 0; }
};
static thread_local PendingCsr g_pending_csr;
struct PendingCsrHost {
    const float* data = nullptr; const float* query = nullptr; const uint32_t* nn_ptr = nullptr;
    size_t n_data = 0, n_query = 0;
    float radius = 0.f; uint32_t max_knn = 0;
    DeviceBuffer d_data, d_query, d_ptr;
    void clear() { d_data.release(); d_query.release(); d_ptr.release(); data = query = nullptr; nn_ptr = nullptr; }
};
static thread_local PendingCsrHost g_pending_csr_host;

#define PGEOF_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) { set_error(__VA_ARGS__); return PGEOF_EINVAL; } \
    } while (0)

extern "C" {

int pgeof_abi_version(void) { return 1; }
const char* pgeof_last_error(void) { return g_error; }

int pgeof_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int pgeof_set_device(int device)
{
    PGEOF_CUDA(cudaSetDevice(device));
    return PGEOF_OK;
}

int pgeof_get_device(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

uint64_t pgeof_launch_count(void) { return g_launches; }
void pgeof_reset_launch_count(void) { g_launches = 0; }

int pgeof_trim(void)
{
    g_pending_csr.clear();            // parked scratch of a radius_search_csr pair (this thread's)
    g_pending_csr_host.clear();
    {
        std::lock_guard<std::mutex> lock(pinned_pool().m);
        pinned_pool().trim_locked();
    }
    int cur = 0, count = 0;
    if (ensure_device(&cur) != PGEOF_OK) return PGEOF_OK;
    PGEOF_CUDA(cudaGetDeviceCount(&count));
    for (int dev = 0; dev < count && dev < 64; ++dev) {      // every device this process left scratch on
        {
            std::lock_guard<std::mutex> lock(arena().m);
            if (arena().free_blocks[dev].empty()) continue;
        }
        PGEOF_CUDA(cudaSetDevice(dev));
        PGEOF_CUDA(cudaDeviceSynchronize());
        std::lock_guard<std::mutex> lock(arena().m);
        arena().trim_locked(dev);
    }
    PGEOF_CUDA(cudaSetDevice(cur));
    return PGEOF_OK;
}

void pgeof_profile_enable(int on) { g_profile = on != 0; }

void pgeof_profile_reset(void)
{
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    for (auto& v : g_profile_events) {
        for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        v.clear();
    }
}

int pgeof_profile_read(const char* name, double* total_ms, uint64_t* launches)
{
    int slot = -1;
    for (int i = 0; i < kNumTimers; ++i) if (name && std::strcmp(name, kTimerNames[i]) == 0) slot = i;
    if (slot < 0) { set_error("unknown kernel timer '%s'", name ? name : "(null)"); return PGEOF_EINVAL; }
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    double sum = 0;
    for (auto& p : g_profile_events[slot]) {
        PGEOF_CUDA(cudaEventSynchronize(p.second));
        float ms = 0;
        PGEOF_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (launches) *launches = g_profile_events[slot].size();
    return PGEOF_OK;
}

void* pgeof_host_alloc(size_t bytes) { return pinned_pool().get(bytes); }
void pgeof_host_free(void* p) { if (p) pinned_pool().put(p); }

// ------------------------------- search ------------------------------------
int pgeof_knn_search_dev(const float* data, size_t n_data, const float* query, size_t n_query, uint32_t knn,
                         uint32_t* indices, float* sqr_dist, void* stream)
{
    PGEOF_REQUIRE(knn <= n_data, "knn size is greater than the data point cloud size");   // nn_search.hpp:37
    if (n_query == 0 || knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && indices && sqr_dist, "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(data));
    return search_run(SEARCH_KNN, data, n_data, query, n_query, knn, 0.f, indices, sqr_dist, nullptr, (cudaStream_t)stream);
}

int pgeof_knn_search(const float* data, size_t n_data, const float* query, size_t n_query, uint32_t knn,
                     uint32_t* indices, float* sqr_dist)
{
    PGEOF_REQUIRE(knn <= n_data, "knn size is greater than the data point cloud size");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_query == 0 || knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && indices && sqr_dist, "null pointer argument");
    HostTrace trace("knn_search", s);
    DeviceBuffer d_data, d_query, d_idx, d_d2;
    PGEOF_TRY(h2d(&d_data, data, n_data * 12, s));
    const bool self = (query == data && n_query == n_data);
    if (!self) PGEOF_TRY(h2d(&d_query, query, n_query * 12, s));
    const size_t out_elems = n_query * (size_t)knn;
    PGEOF_TRY(d_idx.alloc(out_elems * 4, s));
    PGEOF_TRY(d_d2.alloc(out_elems * 4, s));
    trace.mark("alloc+h2d");
    PGEOF_TRY(search_run(SEARCH_KNN, d_data.as<float>(), n_data, self ? d_data.as<float>() : d_query.as<float>(), n_query, knn, 0.f,
                         d_idx.ptr, d_d2.as<float>(), nullptr, s));
    trace.mark("search");
    PGEOF_TRY(d2h(indices, d_idx, out_elems * 4, s));
    trace.mark("d2h idx");
    PGEOF_TRY(d2h(sqr_dist, d_d2, out_elems * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    trace.mark("d2h d2");
    return PGEOF_OK;
}

// kNN straight into CSR (extension): nn = the (n_query, knn) index table flattened, nn_ptr = row * knn in
// uint32 or -- beyond 2^32-1 neighbours, the README's "known limitation" -- uint64.  The squared distances are not
// returned (2 GB less to write out / copy back at 10 M x 50).
__global__ void iota_scale64_kernel(unsigned long long* out, size_t n, uint32_t k)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)i * k;
}

int pgeof_knn_search_csr_dev(const float* data, size_t n_data, const float* query, size_t n_query, uint32_t knn, uint32_t* nn,
                             void* nn_ptr, int ptr_bits, void* stream)
{
    PGEOF_REQUIRE(knn <= n_data, "knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(ptr_bits == 32 || ptr_bits == 64, "ptr_bits must be 32 or 64");
    PGEOF_REQUIRE(nn_ptr, "null pointer argument");
    PGEOF_REQUIRE(ptr_bits == 64 || (uint64_t)n_query * knn <= 0xffffffffull, "n_query * knn exceeds the uint32 CSR limit: ask for 64-bit offsets or shard the queries");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(nn_ptr));
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n_query + 1 + 255) / 256);
    if (ptr_bits == 32) iota_scale_kernel<<<blocks, 256, 0, s>>>((uint32_t*)nn_ptr, n_query + 1, knn);
    else iota_scale64_kernel<<<blocks, 256, 0, s>>>((unsigned long long*)nn_ptr, n_query + 1, knn);
    PGEOF_LAUNCH_CHECK();
    if (n_query == 0 || knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && nn, "null pointer argument");
    DeviceBuffer d2;
    PGEOF_TRY(d2.alloc(n_query * (size_t)knn * 4, s));
    return search_run(SEARCH_KNN, data, n_data, query, n_query, knn, 0.f, nn, d2.as<float>(), nullptr, s);
}

int pgeof_knn_search_csr(const float* data, size_t n_data, const float* query, size_t n_query, uint32_t knn, uint32_t* nn, void* nn_ptr,
                         int ptr_bits)
{
    PGEOF_REQUIRE(knn <= n_data, "knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(ptr_bits == 32 || ptr_bits == 64, "ptr_bits must be 32 or 64");
    PGEOF_REQUIRE(nn_ptr, "null pointer argument");
    PGEOF_REQUIRE(ptr_bits == 64 || (uint64_t)n_query * knn <= 0xffffffffull, "n_query * knn exceeds the uint32 CSR limit: ask for 64-bit offsets or shard the queries");
    // the offsets are arithmetic: written on the host
    if (ptr_bits == 32) for (size_t i = 0; i <= n_query; ++i) ((uint32_t*)nn_ptr)[i] = (uint32_t)(i * knn);
    else for (size_t i = 0; i <= n_query; ++i) ((uint64_t*)nn_ptr)[i] = (uint64_t)i * knn;
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_query == 0 || knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && nn, "null pointer argument");
    DeviceBuffer d_data, d_query, d_idx, d_d2;
    PGEOF_TRY(h2d(&d_data, data, n_data * 12, s));
    const bool self = (query == data && n_query == n_data);
    if (!self) PGEOF_TRY(h2d(&d_query, query, n_query * 12, s));
    const size_t out_elems = n_query * (size_t)knn;
    PGEOF_TRY(d_idx.alloc(out_elems * 4, s));
    PGEOF_TRY(d_d2.alloc(out_elems * 4, s));
    PGEOF_TRY(search_run(SEARCH_KNN, d_data.as<float>(), n_data, self ? d_data.as<float>() : d_query.as<float>(), n_query, knn, 0.f,
                         d_idx.ptr, d_d2.as<float>(), nullptr, s));
    PGEOF_TRY(d2h(nn, d_idx, out_elems * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

int pgeof_radius_search_dev(const float* data, size_t n_data, const float* query, size_t n_query, float search_radius,
                            uint32_t max_knn, int32_t* indices, float* sqr_dist, void* stream)
{
    PGEOF_REQUIRE(max_knn <= n_data, "max knn size is greater than the data point cloud size");   // nn_search.hpp:92-95
    if (n_query == 0 || max_knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && indices && sqr_dist, "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(data));
    return search_run(SEARCH_RADIUS, data, n_data, query, n_query, max_knn, search_radius, indices, sqr_dist, nullptr, (cudaStream_t)stream);
}

int pgeof_radius_search(const float* data, size_t n_data, const float* query, size_t n_query, float search_radius,
                        uint32_t max_knn, int32_t* indices, float* sqr_dist)
{
    PGEOF_REQUIRE(max_knn <= n_data, "max knn size is greater than the data point cloud size");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_query == 0 || max_knn == 0) return PGEOF_OK;
    PGEOF_REQUIRE(data && query && indices && sqr_dist, "null pointer argument");
    DeviceBuffer d_data, d_query, d_idx, d_d2;
    PGEOF_TRY(h2d(&d_data, data, n_data * 12, s));
    const bool self = (query == data && n_query == n_data);
    if (!self) PGEOF_TRY(h2d(&d_query, query, n_query * 12, s));
    const size_t out_elems = n_query * (size_t)max_knn;
    PGEOF_TRY(d_idx.alloc(out_elems * 4, s));
    PGEOF_TRY(d_d2.alloc(out_elems * 4, s));
    PGEOF_TRY(search_run(SEARCH_RADIUS, d_data.as<float>(), n_data, self ? d_data.as<float>() : d_query.as<float>(), n_query, max_knn,
                         search_radius, d_idx.ptr, d_d2.as<float>(), nullptr, s));
    PGEOF_TRY(d2h(indices, d_idx, out_elems * 4, s));
    PGEOF_TRY(d2h(sqr_dist, d_d2, out_elems * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

// Radius search straight into CSR.  Called twice (the caller allocates nn once the total is known): with nn == NULL it
// writes the row offsets and returns the total, with nn it writes the neighbours.  ONE search serves both calls: the first
// runs the search into a (n_query, max_knn) scratch table -- every row writes its count straight into nn_ptr and its hits
// only, no padding, no distances -- and scans the counts; the scratch stays parked (per host thread) and the second
// call only compacts it (8 rows per warp).  If the second call does not match the parked search (other arguments, other thread) it searches
// again -- same result, one search slower.

__global__ void padded_compact_kernel(const int32_t* __restrict__ idx, size_t n_rows, uint32_t max_knn, const uint32_t* __restrict__ nn_ptr,
                                      uint32_t* __restrict__ nn)
{
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const uint32_t b = __ldg(nn_ptr + row), len = __ldg(nn_ptr + row + 1) - b;
    for (uint32_t j = lane; j < len; j += 32) nn[(size_t)b + j] = (uint32_t)__ldg(idx + row * max_knn + j);
}

// max_knn <= 64: a warp moves 8 rows, their (up to) 16 loads in flight together (one row per warp ran at a quarter of the
// memory rate: one load, one store, exit)
__global__ void __launch_bounds__(256) padded_compact8_kernel(const int32_t* __restrict__ idx, size_t n_rows, uint32_t max_knn,
                                                              const uint32_t* __restrict__ nn_ptr, uint32_t* __restrict__ nn)
{
    const size_t r0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8;
    if (r0 >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const uint32_t p = (lane < 9 && r0 + lane <= n_rows) ? __ldg(nn_ptr + r0 + lane) : 0u;
    uint32_t b[8], len[8], lo[8], hi[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        b[u] = __shfl_sync(0xffffffffu, p, u);
        len[u] = r0 + u < n_rows ? __shfl_sync(0xffffffffu, p, u + 1) - b[u] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int32_t* src = idx + (r0 + u) * max_knn;
        lo[u] = (uint32_t)lane < len[u] ? (uint32_t)__ldg(src + lane) : 0u;
        hi[u] = (uint32_t)lane + 32u < len[u] ? (uint32_t)__ldg(src + lane + 32) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if ((uint32_t)lane < len[u]) nn[(size_t)b[u] + lane] = lo[u];
        if ((uint32_t)lane + 32u < len[u]) nn[(size_t)b[u] + lane + 32] = hi[u];
    }
}

int pgeof_radius_search_csr_dev(const float* data, size_t n_data, const float* query, size_t n_query, float search_radius,
                                uint32_t max_knn, uint32_t* nn_ptr, uint32_t* nn, uint64_t* nnz, void* stream)
{
    PGEOF_REQUIRE(max_knn <= n_data, "max knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(nn_ptr && nnz, "null pointer argument");
    // uint32 offsets cap a CSR at 2^32-1 neighbours: shard the queries beyond that
    PGEOF_REQUIRE((uint64_t)n_query * max_knn <= 0xffffffffull, "n_query * max_knn exceeds the uint32 CSR limit; shard the queries");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(nn_ptr));
    cudaStream_t s = (cudaStream_t)stream;
    PendingCsr& pend = g_pending_csr;
    int dev = -1;
    PGEOF_CUDA(cudaGetDevice(&dev));
    // scratch of the single-search path: the padded (n_query, max_knn) table of indices and distances (8 B per slot)
    const bool single = std::getenv("PGEOF_RADIUS_CSR_TWO_PASS") == nullptr && (uint64_t)n_query * max_knn * 8ull <= (16ull << 30);
    if (!nn) {   // call 1: offsets + total
        pend.clear();
        PGEOF_CUDA(cudaMemsetAsync(nn_ptr, 0, (n_query + 1) * sizeof(uint32_t), s));
        if (n_query && max_knn) {
            if (single) {
                // the search writes every row's count into nn_ptr and its hits into the parked table (SearchArgs::nn_ptr)
                PGEOF_TRY(pend.idx.alloc(n_query * (size_t)max_knn * 4, s));
                PGEOF_TRY(search_run(SEARCH_RADIUS, data, n_data, query, n_query, max_knn, search_radius, pend.idx.ptr, nullptr, nn_ptr, s));
                pend.data = data; pend.query = query; pend.n_data = n_data; pend.n_query = n_query; pend.radius = search_radius;
                pend.max_knn = max_knn; pend.nn_ptr = nn_ptr; pend.device = dev;
            } else {
                PGEOF_TRY(search_run(SEARCH_RADIUS_COUNT, data, n_data, query, n_query, max_knn, search_radius, nullptr, nullptr, nn_ptr, s));
            }
        }
        PGEOF_TRY(exclusive_scan_u32(nn_ptr, n_query, s));
        uint32_t total = 0;
        PGEOF_CUDA(cudaMemcpyAsync(&total, nn_ptr + n_query, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        PGEOF_CUDA(cudaStreamSynchronize(s));
        *nnz = total;
        if (total == 0) pend.clear();
        return PGEOF_OK;
    }
    if (n_query == 0 || max_knn == 0) { pend.clear(); return PGEOF_OK; }
    const bool parked = pend.idx.ptr && pend.data == data && pend.query == query && pend.n_data == n_data && pend.n_query == n_query &&
                        pend.radius == search_radius && pend.max_knn == max_knn && pend.nn_ptr == nn_ptr && pend.device == dev &&
                        pend.idx.stream == s;
    if (parked) {   // call 2 of the pair: compact the parked table
        if (max_knn <= 64) padded_compact8_kernel<<<(unsigned)((n_query + 63) / 64), 256, 0, s>>>(pend.idx.as<int32_t>(), n_query, max_knn, nn_ptr, nn);
        else padded_compact_kernel<<<(unsigned)((n_query + 7) / 8), 256, 0, s>>>(pend.idx.as<int32_t>(), n_query, max_knn, nn_ptr, nn);
        PGEOF_LAUNCH_CHECK();
        pend.clear();
        return PGEOF_OK;
    }
    pend.clear();
    return search_run(SEARCH_RADIUS_CSR, data, n_data, query, n_query, max_knn, search_radius, nn, nullptr, nn_ptr, s);
}

// host flavour: the device copies of the cloud / queries / offsets stay parked between the two calls as well, so the pair
// uploads the cloud once and searches once

int pgeof_radius_search_csr(const float* data, size_t n_data, const float* query, size_t n_query, float search_radius,
                            uint32_t max_knn, uint32_t* nn_ptr, uint32_t* nn, uint64_t* nnz)
{
    PGEOF_REQUIRE(max_knn <= n_data, "max knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(nn_ptr && nnz, "null pointer argument");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    PendingCsrHost& pend = g_pending_csr_host;
    const bool self = (query == data && n_query == n_data);
    const bool parked = nn && pend.d_ptr.ptr && pend.data == data && pend.query == query && pend.n_data == n_data && pend.n_query == n_query &&
                        pend.radius == search_radius && pend.max_knn == max_knn && pend.nn_ptr == nn_ptr && pend.d_ptr.stream == s;
    if (!parked) {
        pend.clear();
        g_pending_csr.clear();
        PGEOF_TRY(h2d(&pend.d_data, data, n_data * 12, s));
        if (!self) PGEOF_TRY(h2d(&pend.d_query, query, n_query * 12, s));
        PGEOF_TRY(pend.d_ptr.alloc((n_query + 1) * 4, s));
    }
    const float* dq = self ? pend.d_data.as<float>() : pend.d_query.as<float>();
    if (!nn) {
        const int st = pgeof_radius_search_csr_dev(pend.d_data.as<float>(), n_data, dq, n_query, search_radius, max_knn, pend.d_ptr.as<uint32_t>(), nullptr, nnz, s);
        if (st != PGEOF_OK) { pend.clear(); return st; }
        PGEOF_TRY(d2h(nn_ptr, pend.d_ptr, (n_query + 1) * 4, s));
        PGEOF_CUDA(cudaStreamSynchronize(s));
        pend.data = data; pend.query = query; pend.n_data = n_data; pend.n_query = n_query; pend.radius = search_radius; pend.max_knn = max_knn;
        pend.nn_ptr = nn_ptr;
        return PGEOF_OK;
    }
    if (!parked) PGEOF_CUDA(cudaMemcpyAsync(pend.d_ptr.ptr, nn_ptr, (n_query + 1) * 4, cudaMemcpyHostToDevice, s));
    const size_t total = nn_ptr[n_query];
    DeviceBuffer d_nn;
    int st = d_nn.alloc(std::max<size_t>(total, 1) * 4, s);
    if (st == PGEOF_OK) st = pgeof_radius_search_csr_dev(pend.d_data.as<float>(), n_data, dq, n_query, search_radius, max_knn, pend.d_ptr.as<uint32_t>(), d_nn.as<uint32_t>(), nnz, s);
    if (st == PGEOF_OK) st = d2h(nn, d_nn, total * 4, s);
    if (st == PGEOF_OK && cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); set_error("stream synchronisation failed"); st = PGEOF_ECUDA; }
    pend.clear();
    g_pending_csr.clear();
    if (st != PGEOF_OK) return st;
    *nnz = total;
    return PGEOF_OK;
}

// ------------------------------- features ----------------------------------
// Every CSR feature function exists in four flavours: {host, device} buffers x {uint32 (reference dtype), uint64} row offsets.
static int features_dev_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr ptr, size_t n_rows, uint32_t k_min,
                             int eig_order, float* out, void* stream)
{
    PGEOF_REQUIRE(k_min >= 1, "k_min should be > 1");                                   // pgeof.hpp:81
    if (n_rows == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && (ptr.p32 || ptr.p64) && out && (nn || nnz == 0), "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(ptr.p32 ? (const void*)ptr.p32 : (const void*)ptr.p64));
    return features_run(xyz, n_xyz, nn, nnz, ptr, n_rows, k_min, eig_order, out, (cudaStream_t)stream);
}

static int multiscale_dev_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr ptr, size_t n_rows,
                               const uint32_t* k_scales, size_t n_scales, int eig_order, float* out, void* stream)
{
    PGEOF_REQUIRE(n_scales == 0 || k_scales, "null pointer argument");
    PGEOF_REQUIRE(check_scales(k_scales, n_scales), "k_scales should be > 1 and sorted in ascending order");   // pgeof.hpp:165-168
    if (n_rows == 0 || n_scales == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && (ptr.p32 || ptr.p64) && out && (nn || nnz == 0), "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(ptr.p32 ? (const void*)ptr.p32 : (const void*)ptr.p64));
    return features_multiscale_run(xyz, n_xyz, nn, nnz, ptr, n_rows, k_scales, n_scales, eig_order, out, (cudaStream_t)stream);
}

static int optimal_dev_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, RowPtr ptr, size_t n_rows, uint32_t k_min,
                            uint32_t k_step, uint32_t k_min_search, int eig_order, float* out, void* stream)
{
    PGEOF_REQUIRE(!(k_min < 1 && k_min_search < 1), "k_min and k_min_search should be > 1");   // pgeof.hpp:250 (sic)
    PGEOF_REQUIRE(k_step >= 1, "k_step should be >= 1");                                      // reference: modulo by zero
    if (n_rows == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && (ptr.p32 || ptr.p64) && out && (nn || nnz == 0), "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(ptr.p32 ? (const void*)ptr.p32 : (const void*)ptr.p64));
    return features_optimal_run(xyz, n_xyz, nn, nnz, ptr, n_rows, k_min, k_step, k_min_search, eig_order, out, (cudaStream_t)stream);
}

int pgeof_compute_features_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                               size_t n_rows, uint32_t k_min, int eig_order, float* out, void* stream)
{
    return features_dev_core(xyz, n_xyz, nn, nnz, RowPtr(nn_ptr), n_rows, k_min, eig_order, out, stream);
}
int pgeof_compute_features_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr,
                                   size_t n_rows, uint32_t k_min, int eig_order, float* out, void* stream)
{
    return features_dev_core(xyz, n_xyz, nn, nnz, RowPtr(reinterpret_cast<const unsigned long long*>(nn_ptr)), n_rows, k_min, eig_order, out, stream);
}

int pgeof_compute_features_multiscale_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                                          size_t n_rows, const uint32_t* k_scales, size_t n_scales, int eig_order, float* out,
                                          void* stream)
{
    return multiscale_dev_core(xyz, n_xyz, nn, nnz, RowPtr(nn_ptr), n_rows, k_scales, n_scales, eig_order, out, stream);
}
int pgeof_compute_features_multiscale_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr,
                                              size_t n_rows, const uint32_t* k_scales, size_t n_scales, int eig_order, float* out,
                                              void* stream)
{
    return multiscale_dev_core(xyz, n_xyz, nn, nnz, RowPtr(reinterpret_cast<const unsigned long long*>(nn_ptr)), n_rows, k_scales, n_scales, eig_order, out, stream);
}

int pgeof_compute_features_optimal_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                                       size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order,
                                       float* out, void* stream)
{
    return optimal_dev_core(xyz, n_xyz, nn, nnz, RowPtr(nn_ptr), n_rows, k_min, k_step, k_min_search, eig_order, out, stream);
}
int pgeof_compute_features_optimal_p64_dev(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr,
                                           size_t n_rows, uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order,
                                           float* out, void* stream)
{
    return optimal_dev_core(xyz, n_xyz, nn, nnz, RowPtr(reinterpret_cast<const unsigned long long*>(nn_ptr)), n_rows, k_min, k_step, k_min_search, eig_order, out, stream);
}

// host flavours of the CSR feature functions share the staging code
struct CsrOnDevice {
    DeviceBuffer xyz, nn, ptr, out;
    RowPtr dptr{(const uint32_t*)nullptr};
    int stage(const float* h_xyz, size_t n_xyz, const uint32_t* h_nn, size_t nnz, const void* h_ptr, int ptr_bytes, size_t n_rows, size_t out_floats,
              cudaStream_t s)
    {
        PGEOF_TRY(h2d(&xyz, h_xyz, n_xyz * 12, s));
        PGEOF_TRY(h2d(&nn, h_nn, nnz * 4, s));
        PGEOF_TRY(h2d(&ptr, h_ptr, (n_rows + 1) * (size_t)ptr_bytes, s));
        PGEOF_TRY(out.alloc(out_floats * 4, s));
        dptr = ptr_bytes == 8 ? RowPtr(ptr.as<unsigned long long>()) : RowPtr(ptr.as<uint32_t>());
        return PGEOF_OK;
    }
};

}  // extern "C" (templates need C++ linkage)

// Chunked host pipeline of the CSR feature functions.  The serial flavour uploads all of nn (2 GB at 10 M x 50), computes,
// downloads; here the rows are cut into up to 16 chunks whose nn slices go up on a copy stream into two alternating buffers
// while the previous chunk is computed and its rows go down on the compute stream: both PCIe directions and the SMs are
// busy at once, and the device holds two slices instead of the whole list.  run(nn, nnz_end, row offsets, rows, out)
// sees a pointer `slice - lo`, so the caller's absolute offsets address the slice; rows that point outside [lo, nnz_end)
// (a corrupt, non-monotonic nn_ptr) are refused by the kernels (FeatArgs::nn_lo) exactly like offsets beyond nnz.
// Returns 1 in *done when the pipeline ran, 0 when the input is too small or its chunk boundaries are not monotonic
// (the serial flavour then reports the error).
struct EventPair {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int create() { for (auto& e : ev) PGEOF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return PGEOF_OK; }
    ~EventPair() { for (auto e : ev) if (e) cudaEventDestroy(e); }
};

template <typename Run>
static int csr_pipeline(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const void* nn_ptr, int ptr_bytes, size_t n_rows,
                        size_t floats_per_row, float* out, cudaStream_t s, Run run, int* done)
{
    *done = 0;
    const char* env = std::getenv("PGEOF_HOST_CHUNK_MB");
    const size_t chunk_bytes = (size_t)(env ? std::max(1, std::atoi(env)) : 128) << 20;
    if (env && std::atoi(env) <= 0) return PGEOF_OK;                                     // PGEOF_HOST_CHUNK_MB=0: serial flavour
    const size_t chunks = std::min(std::min((nnz * 4 + chunk_bytes - 1) / chunk_bytes, (size_t)16), n_rows);
    if (chunks < 2 || !nn) return PGEOF_OK;
    auto off = [&](size_t r) -> unsigned long long {
        return ptr_bytes == 8 ? reinterpret_cast<const unsigned long long*>(nn_ptr)[r] : reinterpret_cast<const uint32_t*>(nn_ptr)[r];
    };
    const size_t per = (n_rows + chunks - 1) / chunks;
    unsigned long long max_len = 0, prev = off(0);
    for (size_t c = 0; c < chunks; ++c) {
        const unsigned long long hi = off(std::min(n_rows, (c + 1) * per));
        if (hi < prev || hi > nnz) return PGEOF_OK;
        max_len = std::max(max_len, hi - prev);
        prev = hi;
    }
    cudaStream_t up = g_ctx.copy;
    DeviceBuffer d_xyz, d_ptr, d_out, d_nn[2];
    PGEOF_TRY(h2d(&d_xyz, xyz, n_xyz * 12, s));
    PGEOF_TRY(h2d(&d_ptr, nn_ptr, (n_rows + 1) * (size_t)ptr_bytes, s));
    PGEOF_TRY(d_out.alloc(n_rows * floats_per_row * 4, s));
    for (auto& b : d_nn) PGEOF_TRY(b.alloc(std::max<size_t>((size_t)max_len * 4, 16), s));
    EventPair uploaded, consumed, ready;
    PGEOF_TRY(uploaded.create()); PGEOF_TRY(consumed.create()); PGEOF_TRY(ready.create());
    PGEOF_CUDA(cudaEventRecord(ready.ev[0], s));                                         // the buffers exist (arena blocks are stream ordered)
    PGEOF_CUDA(cudaStreamWaitEvent(up, ready.ev[0], 0));
    auto upload = [&](size_t c) -> int {
        const size_t r0 = c * per, r1 = std::min(n_rows, (c + 1) * per);
        const unsigned long long lo = off(r0), hi = off(r1);
        if (c >= 2) PGEOF_CUDA(cudaStreamWaitEvent(up, consumed.ev[c & 1], 0));
        if (hi > lo) PGEOF_CUDA(cudaMemcpyAsync(d_nn[c & 1].ptr, nn + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, up));
        PGEOF_CUDA(cudaEventRecord(uploaded.ev[c & 1], up));
        return PGEOF_OK;
    };
    PGEOF_TRY(upload(0));
    int status = PGEOF_OK;
    for (size_t c = 0; c < chunks && status == PGEOF_OK; ++c) {
        const size_t r0 = c * per, r1 = std::min(n_rows, (c + 1) * per);
        if (r1 <= r0) break;
        if (c + 1 < chunks) PGEOF_TRY(upload(c + 1));
        const unsigned long long lo = off(r0), hi = off(r1);
        PGEOF_CUDA(cudaStreamWaitEvent(s, uploaded.ev[c & 1], 0));
        g_nn_window_lo = lo;
        const RowPtr rows_ptr = ptr_bytes == 8 ? RowPtr(d_ptr.as<unsigned long long>() + r0) : RowPtr(d_ptr.as<uint32_t>() + r0);
        status = run(d_xyz.as<float>(), d_nn[c & 1].as<uint32_t>() - lo, (size_t)hi, rows_ptr, r1 - r0, d_out.as<float>() + r0 * floats_per_row);
        g_nn_window_lo = 0;
        PGEOF_CUDA(cudaEventRecord(consumed.ev[c & 1], s));
        if (status == PGEOF_OK)
            PGEOF_CUDA(cudaMemcpyAsync(out + r0 * floats_per_row, d_out.as<float>() + r0 * floats_per_row, (r1 - r0) * floats_per_row * 4, cudaMemcpyDeviceToHost, s));
    }
    PGEOF_CUDA(cudaStreamSynchronize(up));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    if (status != PGEOF_OK) return status;
    *done = 1;
    return PGEOF_OK;
}

extern "C" {

static int features_host_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const void* nn_ptr, int ptr_bytes, size_t n_rows,
                              uint32_t k_min, int eig_order, float* out)
{
    PGEOF_REQUIRE(k_min >= 1, "k_min should be > 1");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_rows == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && nn_ptr && out && (nn || nnz == 0), "null pointer argument");
    HostTrace trace("compute_features", s);
    int piped = 0;
    PGEOF_TRY(csr_pipeline(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, 11, out, s,
                           [&](const float* dx, const uint32_t* dnn, size_t end, RowPtr rp, size_t rows, float* dout) {
                               return features_run(dx, n_xyz, dnn, end, rp, rows, k_min, eig_order, dout, s);
                           }, &piped));
    if (piped) { trace.mark("chunked h2d | features | d2h"); return PGEOF_OK; }
    CsrOnDevice d;
    PGEOF_TRY(d.stage(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, n_rows * 11, s));
    trace.mark("alloc+h2d");
    PGEOF_TRY(features_run(d.xyz.as<float>(), n_xyz, d.nn.as<uint32_t>(), nnz, d.dptr, n_rows, k_min, eig_order, d.out.as<float>(), s));
    trace.mark("features");
    PGEOF_TRY(d2h(out, d.out, n_rows * 11 * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    trace.mark("d2h");
    return PGEOF_OK;
}

static int multiscale_host_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const void* nn_ptr, int ptr_bytes, size_t n_rows,
                                const uint32_t* k_scales, size_t n_scales, int eig_order, float* out)
{
    PGEOF_REQUIRE(n_scales == 0 || k_scales, "null pointer argument");
    PGEOF_REQUIRE(check_scales(k_scales, n_scales), "k_scales should be > 1 and sorted in ascending order");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_rows == 0 || n_scales == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && nn_ptr && out && (nn || nnz == 0), "null pointer argument");
    int piped = 0;
    PGEOF_TRY(csr_pipeline(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, n_scales * 11, out, s,
                           [&](const float* dx, const uint32_t* dnn, size_t end, RowPtr rp, size_t rows, float* dout) {
                               return features_multiscale_run(dx, n_xyz, dnn, end, rp, rows, k_scales, n_scales, eig_order, dout, s);
                           }, &piped));
    if (piped) return PGEOF_OK;
    CsrOnDevice d;
    PGEOF_TRY(d.stage(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, n_rows * n_scales * 11, s));
    PGEOF_TRY(features_multiscale_run(d.xyz.as<float>(), n_xyz, d.nn.as<uint32_t>(), nnz, d.dptr, n_rows, k_scales, n_scales,
                                      eig_order, d.out.as<float>(), s));
    PGEOF_TRY(d2h(out, d.out, n_rows * n_scales * 11 * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

static int optimal_host_core(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const void* nn_ptr, int ptr_bytes, size_t n_rows,
                             uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out)
{
    PGEOF_REQUIRE(!(k_min < 1 && k_min_search < 1), "k_min and k_min_search should be > 1");
    PGEOF_REQUIRE(k_step >= 1, "k_step should be >= 1");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n_rows == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && nn_ptr && out && (nn || nnz == 0), "null pointer argument");
    int piped = 0;
    PGEOF_TRY(csr_pipeline(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, 12, out, s,
                           [&](const float* dx, const uint32_t* dnn, size_t end, RowPtr rp, size_t rows, float* dout) {
                               return features_optimal_run(dx, n_xyz, dnn, end, rp, rows, k_min, k_step, k_min_search, eig_order, dout, s);
                           }, &piped));
    if (piped) return PGEOF_OK;
    CsrOnDevice d;
    PGEOF_TRY(d.stage(xyz, n_xyz, nn, nnz, nn_ptr, ptr_bytes, n_rows, n_rows * 12, s));
    PGEOF_TRY(features_optimal_run(d.xyz.as<float>(), n_xyz, d.nn.as<uint32_t>(), nnz, d.dptr, n_rows, k_min, k_step,
                                   k_min_search, eig_order, d.out.as<float>(), s));
    PGEOF_TRY(d2h(out, d.out, n_rows * 12 * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

int pgeof_compute_features(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr, size_t n_rows,
                           uint32_t k_min, int eig_order, float* out)
{
    return features_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 4, n_rows, k_min, eig_order, out);
}
int pgeof_compute_features_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr, size_t n_rows,
                               uint32_t k_min, int eig_order, float* out)
{
    return features_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 8, n_rows, k_min, eig_order, out);
}

int pgeof_compute_features_multiscale(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr,
                                      size_t n_rows, const uint32_t* k_scales, size_t n_scales, int eig_order, float* out)
{
    return multiscale_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 4, n_rows, k_scales, n_scales, eig_order, out);
}
int pgeof_compute_features_multiscale_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr,
                                          size_t n_rows, const uint32_t* k_scales, size_t n_scales, int eig_order, float* out)
{
    return multiscale_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 8, n_rows, k_scales, n_scales, eig_order, out);
}

int pgeof_compute_features_optimal(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint32_t* nn_ptr, size_t n_rows,
                                   uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out)
{
    return optimal_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 4, n_rows, k_min, k_step, k_min_search, eig_order, out);
}
int pgeof_compute_features_optimal_p64(const float* xyz, size_t n_xyz, const uint32_t* nn, size_t nnz, const uint64_t* nn_ptr, size_t n_rows,
                                       uint32_t k_min, uint32_t k_step, uint32_t k_min_search, int eig_order, float* out)
{
    return optimal_host_core(xyz, n_xyz, nn, nnz, nn_ptr, 8, n_rows, k_min, k_step, k_min_search, eig_order, out);
}

// ------------------------------- selected ----------------------------------
int pgeof_compute_features_selected_f32_dev(const float* xyz, size_t n, float search_radius, uint32_t max_knn, const int32_t* feature_ids,
                                            size_t n_features, int eig_order, float* out, void* stream)
{
    if (n == 0 || n_features == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && feature_ids && out, "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(xyz));
    return selected_run_f32(xyz, n, search_radius, max_knn, feature_ids, n_features, eig_order, out, (cudaStream_t)stream);
}

int pgeof_compute_features_selected_f64_dev(const double* xyz, size_t n, double search_radius, uint32_t max_knn, const int32_t* feature_ids,
                                            size_t n_features, int eig_order, double* out, void* stream)
{
    if (n == 0 || n_features == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && feature_ids && out, "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(xyz));
    return selected_run_f64(xyz, n, search_radius, max_knn, feature_ids, n_features, eig_order, out, (cudaStream_t)stream);
}

int pgeof_compute_features_selected_f32(const float* xyz, size_t n, float search_radius, uint32_t max_knn, const int32_t* feature_ids,
                                        size_t n_features, int eig_order, float* out)
{
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n == 0 || n_features == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && feature_ids && out, "null pointer argument");
    DeviceBuffer d_xyz, d_out;
    PGEOF_TRY(h2d(&d_xyz, xyz, n * 12, s));
    PGEOF_TRY(d_out.alloc(n * n_features * 4, s));
    PGEOF_TRY(selected_run_f32(d_xyz.as<float>(), n, search_radius, max_knn, feature_ids, n_features, eig_order, d_out.as<float>(), s));
    PGEOF_TRY(d2h(out, d_out, n * n_features * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

int pgeof_compute_features_selected_f64(const double* xyz, size_t n, double search_radius, uint32_t max_knn, const int32_t* feature_ids,
                                        size_t n_features, int eig_order, double* out)
{
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n == 0 || n_features == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && feature_ids && out, "null pointer argument");
    DeviceBuffer d_xyz, d_out;
    PGEOF_TRY(h2d(&d_xyz, xyz, n * 24, s));
    PGEOF_TRY(d_out.alloc(n * n_features * 8, s));
    PGEOF_TRY(selected_run_f64(d_xyz.as<double>(), n, search_radius, max_knn, feature_ids, n_features, eig_order, d_out.as<double>(), s));
    PGEOF_TRY(d2h(out, d_out, n * n_features * 8, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

// ------------------------------- fused pipeline ----------------------------
int pgeof_knn_features_dev(const float* xyz, size_t n, uint32_t knn, uint32_t k_min, int eig_order, uint32_t* indices, float* sqr_dist,
                           float* features, void* stream)
{
    PGEOF_REQUIRE(knn <= n, "knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(k_min >= 1, "k_min should be > 1");
    if (n == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && features, "null pointer argument");
    PGEOF_REQUIRE((uint64_t)n * knn <= 0xffffffffull, "n * knn exceeds the uint32 CSR limit; shard the queries");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(xyz));
    cudaStream_t s = (cudaStream_t)stream;
    if (!indices && !sqr_dist) {     // nobody wants the neighbour lists: never materialise them
        int done = 0;
        PGEOF_TRY(knn_features_fused_run(xyz, n, knn, k_min, eig_order, features, s, &done));
        if (done) return PGEOF_OK;
    }
    DeviceBuffer t_idx, t_d2, ptr;
    if (!indices) { PGEOF_TRY(t_idx.alloc(n * (size_t)knn * 4, s)); indices = t_idx.as<uint32_t>(); }
    if (!sqr_dist) { PGEOF_TRY(t_d2.alloc(n * (size_t)knn * 4, s)); sqr_dist = t_d2.as<float>(); }
    if (knn) PGEOF_TRY(search_run(SEARCH_KNN, xyz, n, xyz, n, knn, 0.f, indices, sqr_dist, nullptr, s));
    PGEOF_TRY(ptr.alloc((n + 1) * 4, s));
    iota_scale_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(ptr.as<uint32_t>(), n + 1, knn);
    PGEOF_LAUNCH_CHECK();
    return features_run(xyz, n, indices, n * (size_t)knn, ptr.as<uint32_t>(), n, k_min, eig_order, features, s);
}

int pgeof_knn_features(const float* xyz, size_t n, uint32_t knn, uint32_t k_min, int eig_order, uint32_t* indices, float* sqr_dist,
                       float* features)
{
    PGEOF_REQUIRE(knn <= n, "knn size is greater than the data point cloud size");
    PGEOF_REQUIRE(k_min >= 1, "k_min should be > 1");
    cudaStream_t s;
    PGEOF_TRY(host_stream(&s));
    if (n == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && features, "null pointer argument");
    DeviceBuffer d_xyz, d_idx, d_d2, d_feat;
    PGEOF_TRY(h2d(&d_xyz, xyz, n * 12, s));
    PGEOF_TRY(d_feat.alloc(n * 11 * 4, s));
    const size_t rows = n * (size_t)knn;
    if (indices) PGEOF_TRY(d_idx.alloc(std::max<size_t>(rows, 1) * 4, s));
    if (sqr_dist) PGEOF_TRY(d_d2.alloc(std::max<size_t>(rows, 1) * 4, s));
    PGEOF_TRY(pgeof_knn_features_dev(d_xyz.as<float>(), n, knn, k_min, eig_order, indices ? d_idx.as<uint32_t>() : nullptr,
                                     sqr_dist ? d_d2.as<float>() : nullptr, d_feat.as<float>(), s));
    PGEOF_TRY(d2h(features, d_feat, n * 11 * 4, s));
    if (indices) PGEOF_TRY(d2h(indices, d_idx, rows * 4, s));
    if (sqr_dist) PGEOF_TRY(d2h(sqr_dist, d_d2, rows * 4, s));
    PGEOF_CUDA(cudaStreamSynchronize(s));
    return PGEOF_OK;
}

// ------------------------------- query shards ------------------------------
int pgeof_slab_plan_dev(const float* xyz, size_t n, int rank, int world, int axis, pgeof_slab_plan* plan, void* stream)
{
    PGEOF_REQUIRE(plan, "null pointer argument");
    PGEOF_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    PGEOF_REQUIRE(axis >= 0 && axis < 3, "axis must be 0, 1 or 2");
    plan->lo = 0.f; plan->scale = 0.f; plan->bin_lo = 0; plan->bin_hi = 0; plan->count = 0; plan->axis = axis;
    if (n == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz, "null pointer argument");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(xyz));
    return slab_plan_run(xyz, n, rank, world, axis, plan, (cudaStream_t)stream);
}

int pgeof_slab_fill_dev(const float* xyz, size_t n, const pgeof_slab_plan* plan, int64_t* rows, float* query, void* stream)
{
    PGEOF_REQUIRE(plan, "null pointer argument");
    if (n == 0 || plan->count == 0) return PGEOF_OK;
    PGEOF_REQUIRE(xyz && rows && query, "null pointer argument");
    PGEOF_REQUIRE(plan->axis >= 0 && plan->axis < 3 && plan->bin_lo <= plan->bin_hi, "bad slab plan");
    DeviceGuard guard;
    PGEOF_TRY(guard.enter(xyz));
    return slab_fill_run(xyz, n, plan, reinterpret_cast<long long*>(rows), query, (cudaStream_t)stream);
}

}  // extern "C"
