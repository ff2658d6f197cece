// passes_dispatch.cu -- run-time length -> instantiated kernel, and the per-device twiddle tables.
#include <map>
#include <mutex>
#include <utility>

#include "fft_mid.cuh"
#include "passes.h"

namespace fdr {
#define X(LOGN)                                                              \
    cudaError_t launch_row_pass_##LOGN(const RowPassArgs&, cudaStream_t);    \
    cudaError_t launch_col_pass_##LOGN(const ColPassArgs&, cudaStream_t);    \
    cudaError_t tw_fill_##LOGN(float2*, cudaStream_t);                       \
    int tw_total_##LOGN();
#define FDR_ALL_LOGNS X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)
FDR_ALL_LOGNS
#undef X

cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({func, dev});
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    done[{func, dev}] = bytes;
    return cudaSuccess;
}

int device_sm_count() {
    static std::mutex mu;
    static std::map<int, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    cache[dev] = n;
    return n;
}

static int ilog2_exact(int n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

cudaError_t launch_row_pass(const RowPassArgs& a, cudaStream_t s) {
    switch (ilog2_exact(a.n)) {
#define X(LOGN) \
    case LOGN: return launch_row_pass_##LOGN(a, s);
        FDR_ALL_LOGNS
#undef X
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_col_pass(const ColPassArgs& a, cudaStream_t s) {
    if (a.col_variant != 1 && col_tma_applicable(a)) return col_wide_applicable(a) ? launch_col_wiener_wide(a, s) : launch_col_wiener_tma(a, s);
    switch (ilog2_exact(a.n)) {
#define X(LOGN) \
    case LOGN: return launch_col_pass_##LOGN(a, s);
        FDR_ALL_LOGNS
#undef X
    }
    return cudaErrorInvalidValue;
}

// Twiddle table of length n on the current device: built once (double-precision sincospi on the
// device), cached for the life of the process.
cudaError_t get_twiddles(int n, const float2** out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2*> cache;
    *out = nullptr;
    const int l = ilog2_exact(n);
    if (l < 0 || l > 14) return cudaErrorInvalidValue;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, l});
    if (it != cache.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    int total = 0;
    switch (l) {
#define X(LOGN) \
    case LOGN: total = tw_total_##LOGN(); break;
        FDR_ALL_LOGNS
#undef X
    }
    float2* p = nullptr;
    e = cudaMalloc(&p, sizeof(float2) * (size_t)(total > 0 ? total : 1));
    if (e != cudaSuccess) return e;
    switch (l) {
#define X(LOGN) \
    case LOGN: e = tw_fill_##LOGN(p, 0); break;
        FDR_ALL_LOGNS
#undef X
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaFree(p);
        return e;
    }
    cache[{dev, l}] = p;
    *out = p;
    return cudaSuccess;
}

cudaError_t get_twiddles_mid(int n, const float2** out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2*> cache;
    *out = nullptr;
    if (n != 8192 && n != 16384) return cudaErrorInvalidValue;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, n});
    if (it != cache.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    const int total = (n == 16384) ? MidGeom<16384>::TW_ENTRIES : MidGeom<8192>::TW_ENTRIES;
    float2* p = nullptr;
    e = cudaMalloc(&p, sizeof(float2) * (size_t)total);
    if (e != cudaSuccess) return e;
    if (n == 16384)
        mid_tw_fill_kernel<16384><<<(total + 255) / 256, 256>>>(p);
    else
        mid_tw_fill_kernel<8192><<<(total + 255) / 256, 256>>>(p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaFree(p);
        return e;
    }
    cache[{dev, n}] = p;
    *out = p;
    return cudaSuccess;
}

int col_pass_tile_width(int n) {
    int l = ilog2_exact(n);
    return l <= 6 ? 32 : l <= 8 ? 16 : l <= 10 ? 8 : l == 11 ? 4 : l == 12 ? 2 : 1;
}
}  // namespace fdr
