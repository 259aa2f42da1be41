// Process-wide cache of device and pinned-host allocations (see common.hpp).
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.hpp"

namespace rambl {

namespace {
struct Cache
{
    std::mutex mu;
    std::multimap<size_t, void*> device, pinned;  // granted size -> free block
};
Cache& cache()
{
    static Cache* c = new Cache;  // never destroyed: blocks may be released after main() returns
    return *c;
}
int current_device()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    return d;
}
// device blocks are listed per device: granted sizes are powers of two below 2^48, the device goes above
size_t device_key(size_t granted, int device) { return granted | ((size_t)(device & 0xff) << 48); }
size_t round_up(size_t bytes)
{
    size_t g = 256;
    while (g < bytes) g <<= 1;
    return g;
}
}  // namespace

void* cached_device_alloc(size_t bytes, size_t* granted)
{
    const size_t g = round_up(bytes);
    *granted = g;
    {
        std::lock_guard<std::mutex> lk(cache().mu);
        auto it = cache().device.find(device_key(g, current_device()));
        if (it != cache().device.end())
        {
            void* p = it->second;
            cache().device.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, g);
    if (e != cudaSuccess)
    {   // give the driver back what we hold and try once more
        release_cached_memory();
        e = cudaMalloc(&p, g);
    }
    if (e != cudaSuccess) throw Error(RAMBL_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return p;
}

void cached_device_free(void* p, size_t granted)
{
    if (!p) return;
    // a block goes back to the list of the device it lives on (one process may drive several devices)
    int dev = current_device();
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
    else cudaGetLastError();
    std::lock_guard<std::mutex> lk(cache().mu);
    cache().device.insert({device_key(granted, dev), p});
}

void* cached_pinned_alloc(size_t bytes, size_t* granted)
{
    const size_t g = round_up(bytes);
    *granted = g;
    {
        std::lock_guard<std::mutex> lk(cache().mu);
        auto it = cache().pinned.find(g);
        if (it != cache().pinned.end())
        {
            void* p = it->second;
            cache().pinned.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, g);
    if (e != cudaSuccess) throw Error(RAMBL_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    return p;
}

void cached_pinned_free(void* p, size_t granted)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(cache().mu);
    cache().pinned.insert({granted, p});
}

// ---- pageable host blocks (common.hpp: cached_host_alloc).  A block carries its capacity in a 64-byte header; a
// request takes the smallest kept block that is large enough and at most half as large again.  When the kept blocks
// would exceed the limit, the ones released longest ago go back to the allocator first (a workload whose subgroup sizes
// change must not be left with a cache full of blocks that fit nothing).
namespace {
const size_t kHostHeader = 64, kHostSmall = 64 * 1024;
struct HostCache
{
    struct Kept { void* base; unsigned long long stamp; };
    std::mutex mu;
    std::multimap<size_t, Kept> by_size;                                            // capacity -> block
    std::map<unsigned long long, std::multimap<size_t, Kept>::iterator> by_age;     // release order -> its entry
    unsigned long long clock = 0;
    size_t held = 0, limit = 0;
};
HostCache& host_cache()
{
    static HostCache* c = [] {
        HostCache* h = new HostCache;
        const char* e = getenv("RAMBL_HOST_CACHE_MB");
        h->limit = (size_t)(e ? std::max(0L, atol(e)) : 16384L) << 20;
        return h;
    }();
    return *c;
}
}  // namespace

void* cached_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    if (bytes > (size_t)-1 - 0x20000) throw std::bad_alloc();  // the rounding below must not wrap
    size_t cap = bytes;
    void* base = nullptr;
    if (bytes >= kHostSmall)
    {
        cap = (bytes + 0xffff) & ~(size_t)0xffff;
        HostCache& hc = host_cache();
        std::lock_guard<std::mutex> lk(hc.mu);
        auto it = hc.by_size.lower_bound(cap);
        if (it != hc.by_size.end() && it->first <= cap + cap / 2)
        {
            base = it->second.base;
            cap = it->first;
            hc.held -= cap;
            hc.by_age.erase(it->second.stamp);
            hc.by_size.erase(it);
        }
    }
    if (!base)
    {
        base = malloc(kHostHeader + cap);
        if (!base) throw std::bad_alloc();
    }
    *static_cast<size_t*>(base) = cap;
    return static_cast<char*>(base) + kHostHeader;
}

void cached_host_free(void* p)
{
    if (!p) return;
    void* base = static_cast<char*>(p) - kHostHeader;
    const size_t cap = *static_cast<size_t*>(base);
    std::vector<void*> evicted;
    bool kept = false;
    if (cap >= kHostSmall)
    {
        HostCache& hc = host_cache();
        std::lock_guard<std::mutex> lk(hc.mu);
        if (cap <= hc.limit)
        {
            while (hc.held + cap > hc.limit && !hc.by_age.empty())
            {
                auto oldest = hc.by_age.begin();
                hc.held -= oldest->second->first;
                evicted.push_back(oldest->second->second.base);
                hc.by_size.erase(oldest->second);
                hc.by_age.erase(oldest);
            }
            const unsigned long long stamp = ++hc.clock;
            hc.by_age[stamp] = hc.by_size.insert({cap, HostCache::Kept{base, stamp}});
            hc.held += cap;
            kept = true;
        }
    }
    for (void* q : evicted) free(q);
    if (!kept) free(base);
}

size_t cached_host_bytes()
{
    std::lock_guard<std::mutex> lk(host_cache().mu);
    return host_cache().held;
}

void release_cached_memory()
{
    std::multimap<size_t, void*> d, h;
    {
        std::lock_guard<std::mutex> lk(cache().mu);
        d.swap(cache().device);
        h.swap(cache().pinned);
    }
    std::vector<void*> blocks;
    {
        std::lock_guard<std::mutex> lk(host_cache().mu);
        for (auto& kv : host_cache().by_size) blocks.push_back(kv.second.base);
        host_cache().by_size.clear();
        host_cache().by_age.clear();
        host_cache().held = 0;
    }
    for (auto& kv : d) cudaFree(kv.second);
    for (auto& kv : h) cudaFreeHost(kv.second);
    for (void* q : blocks) free(q);
}

}  // namespace rambl

namespace rambl {

static std::atomic<int> g_host_threads{0};

void set_host_threads(int n) { g_host_threads.store(n > 0 ? n : 0); }

unsigned host_threads()
{
    int n = g_host_threads.load();
    if (n <= 0)
    {
        const char* e = getenv("RAMBL_HOST_THREADS");
        if (e) n = atoi(e);
    }
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    return (unsigned)std::min(n, 256);
}

}  // namespace rambl
