// Shared helpers of the B200 StrainCall engine (host side).
#pragma once
#include <cuda_runtime.h>

#include "rambl_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rambl {

struct Error : std::runtime_error
{
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// status codes: include/rambl_b200.h (RAMBL_OK, RAMBL_ERR_*)

#define RAMBL_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            throw ::rambl::Error(RAMBL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// Device and pinned-host memory come from a process-wide cache: cudaMalloc / cudaFree / cudaMallocHost /
// cudaFreeHost synchronise the device and cost from tens of microseconds to a second (teardown of one
// strain search was measured at up to 1.3 s), so buffers released by one call are handed to the next
// instead of going back to the driver.  Sizes are rounded up to a power of two.  rambl_release_cached_memory()
// (C ABI) returns everything to the driver.
void* cached_device_alloc(size_t bytes, size_t* granted);
void cached_device_free(void* p, size_t granted);
void* cached_pinned_alloc(size_t bytes, size_t* granted);
void cached_pinned_free(void* p, size_t granted);
void release_cached_memory();

// Pageable host memory for the large per-subgroup arrays (the read pools of a flat graph: ~9 MB per 5 000-read subgroup).
// Fresh pages cost a fault each -- measured here at a fifth of graph construction, more when 16 workers fault at
// once -- so blocks of 64 KB and more are kept when released and handed to the next batch (up to $RAMBL_HOST_CACHE_MB,
// default 16 384; smaller requests are plain malloc).  release_cached_memory() frees them as well.
void* cached_host_alloc(size_t bytes);
void cached_host_free(void* p);
size_t cached_host_bytes();  // held by the cache right now

// std::vector allocator over the cache above; elements are default-initialised (resize() does not zero what the caller
// is about to overwrite).
template <typename T>
struct HostCacheAlloc
{
    typedef T value_type;
    HostCacheAlloc() {}
    template <typename U> HostCacheAlloc(const HostCacheAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(cached_host_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t) { cached_host_free(p); }
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename A0, typename... A> void construct(U* p, A0&& a0, A&&... a)
    {
        ::new (static_cast<void*>(p)) U(std::forward<A0>(a0), std::forward<A>(a)...);
    }
    template <typename U> bool operator==(const HostCacheAlloc<U>&) const { return true; }
    template <typename U> bool operator!=(const HostCacheAlloc<U>&) const { return false; }
};
template <typename T> using HostVec = std::vector<T, HostCacheAlloc<T>>;

// A device buffer that only grows; reused across launches so the level loop does not malloc.
template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0;      // elements
    size_t bytes = 0;    // granted by the cache
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap), bytes(o.bytes) { o.p = nullptr; o.cap = 0; o.bytes = 0; }
    ~DevBuf() { if (p) cached_device_free(p, bytes); }
    void swap(DevBuf& o) { std::swap(p, o.p); std::swap(cap, o.cap); std::swap(bytes, o.bytes); }
    // grows (contents are NOT kept unless keep=true)
    void reserve(size_t n, bool keep = false, cudaStream_t st = 0)
    {
        if (n <= cap) return;
        size_t granted = 0;
        T* q = static_cast<T*>(cached_device_alloc(std::max<size_t>(n, 64) * sizeof(T), &granted));
        if (keep && p && cap)
        {
            RAMBL_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
            RAMBL_CUDA(cudaStreamSynchronize(st));
        }
        if (p) cached_device_free(p, bytes);
        p = q;
        bytes = granted;
        cap = granted / sizeof(T);
    }
};

// Pinned host staging buffer (grow-only).
template <typename T>
struct PinBuf
{
    T* p = nullptr;
    size_t cap = 0;
    size_t bytes = 0;
    PinBuf() {}
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cached_pinned_free(p, bytes); }
    void reserve(size_t n)
    {
        if (n <= cap) return;
        size_t granted = 0;
        T* q = static_cast<T*>(cached_pinned_alloc(std::max<size_t>(n, 64) * sizeof(T), &granted));
        if (p) cached_pinned_free(p, bytes);
        p = q;
        bytes = granted;
        cap = granted / sizeof(T);
    }
};

void require_device();  // throws RAMBL_ERR_CUDA when no B200-class device is present

// Host worker threads this process may use for the per-subgroup host work (graph construction, staging):
// rambl_set_host_threads(), else $RAMBL_HOST_THREADS, else the hardware concurrency.  Several ranks on one box
// each take their share of the cores instead of all spawning one worker per core.
unsigned host_threads();
void set_host_threads(int n);

}  // namespace rambl
