// Error-propagation macros of the C-ABI functions and an owning device allocation that only ever grows
// (engine.cu, metric.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/vitdet_b200.h"
#include "kernels.h"

#define CU_TRY(expr)                                                                               \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(VITDET_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define RC_TRY(expr)                   \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

namespace vitdet {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            if (p) cudaFree(p);
            p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0;
        }
        return *this;
    }
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        need = (need + 255) / 256 * 256;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) return fail(VITDET_E_CUDA, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        bytes = need;
        return 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace vitdet
