// capi_internal.h -- helpers shared by the translation units behind the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/fdr_b200.h"
#include "passes.h"

namespace fdr {

int set_error(int code, const char* fmt, ...);  // stores the calling thread's last-error text, returns code

#define FDR_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return fdr::set_error(FDR_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define FDR_TRY(call)                    \
    do {                                 \
        int rc__ = (call);               \
        if (rc__ != FDR_OK) return rc__; \
    } while (0)

#define FDR_API extern "C" __attribute__((visibility("default")))

inline int next_pow2(int n) {  // utils.hpp:27-31
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}
inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
inline int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return FDR_OK;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return set_error(FDR_E_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return FDR_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

template <typename T> struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return FDR_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMallocHost(&p, count * sizeof(T));
        if (e != cudaSuccess) return set_error(FDR_E_NOMEM, "cudaMallocHost(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return FDR_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
};

// Inverse rotation of motionBlurKernel (utils.hpp:16-22), as OpenCV derives it.
PsfAffine motion_affine(int size, double angle_deg);

}  // namespace fdr
