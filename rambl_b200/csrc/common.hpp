// Shared helpers of the B200 StrainCall engine (host side).
#pragma once
#include <cuda_runtime.h>

#include "rambl_b200.h"

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
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

// A device buffer that only grows; reused across launches so the level loop does not malloc.
template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    ~DevBuf() { if (p) cudaFree(p); }
    // grows (contents are NOT kept unless keep=true)
    void reserve(size_t n, bool keep = false, cudaStream_t st = 0)
    {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap *= 2;
        T* q = nullptr;
        RAMBL_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
        if (keep && p && cap)
        {
            RAMBL_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
            RAMBL_CUDA(cudaStreamSynchronize(st));
        }
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
    }
};

// Pinned host staging buffer (grow-only).
template <typename T>
struct PinBuf
{
    T* p = nullptr;
    size_t cap = 0;
    PinBuf() {}
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    void reserve(size_t n)
    {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap *= 2;
        T* q = nullptr;
        RAMBL_CUDA(cudaMallocHost(&q, ncap * sizeof(T)));
        if (p) cudaFreeHost(p);
        p = q;
        cap = ncap;
    }
};

void require_device();  // throws RAMBL_ERR_CUDA when no B200-class device is present

}  // namespace rambl
