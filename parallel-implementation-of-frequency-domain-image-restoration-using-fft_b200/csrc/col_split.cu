// col_split.cu -- host side of the long-column pass (see col_split.cuh).
#include <cstdlib>
#include <map>
#include <mutex>

#include "col_split.cuh"

namespace fdr {

__global__ void tw_full_fill_kernel(float2* tw, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    sincospi(2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_float2((float)c, (float)(-s));
}

cudaError_t get_full_twiddles(int n, const float2** out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2*> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, n});
    if (it != cache.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    float2* p = nullptr;
    e = cudaMalloc(&p, sizeof(float2) * (size_t)n);
    if (e != cudaSuccess) return e;
    tw_full_fill_kernel<<<(n + 255) / 256, 256>>>(p, n);
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

bool col_split_applicable(const ColPassArgs& a) {
    return (a.n == 8192 || a.n == 16384) && (a.mode == COL_WIENER || a.mode == COL_MAKE_WIENER) && a.pitch % SPLIT_CWC == 0 && !a.conj;
}

template <int LOGN1> static cudaError_t launch_strided(const ColSplitArgs& s, cudaStream_t st) {
    constexpr int N1 = 1 << LOGN1;
    constexpr int threads = N1 / 16 * SPLIT_CWC * SPLIT_NJ;
    constexpr size_t smem = (size_t)N1 * SPLIT_CWC * SPLIT_NJ * sizeof(float2);
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(col_split_strided_kernel<LOGN1>), smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(s.ncols / SPLIT_CWC, SPLIT_N2 / SPLIT_NJ, s.npairs);
    col_split_strided_kernel<LOGN1><<<grid, threads, smem, st>>>(s);
    return cudaGetLastError();
}

static cudaError_t launch_block(const ColSplitArgs& s, cudaStream_t st) {
    constexpr int threads = SPLIT_N2 / 16 * SPLIT_CWC * SPLIT_NJ;
    constexpr size_t smem = (size_t)SPLIT_N2 * SPLIT_CWC * SPLIT_NJ * sizeof(float2);
    dim3 grid(s.ncols / SPLIT_CWC, (s.n / SPLIT_N2) / SPLIT_NJ, s.npairs);
    if (s.mode == COL_MAKE_WIENER)
        col_split_block_kernel<COL_MAKE_WIENER><<<grid, threads, smem, st>>>(s);
    else
        col_split_block_kernel<COL_WIENER><<<grid, threads, smem, st>>>(s);
    return cudaGetLastError();
}

// Whole column pass, panel by panel; *launches receives the number of kernels launched.
int col_split_block_len(const ColPassArgs& a) { return col_blocks_applicable(a) ? 2048 : SPLIT_N2; }

cudaError_t launch_col_split(const ColPassArgs& a, cudaStream_t st, int* launches) {
    if (col_blocks_applicable(a)) return launch_col_blocks(a, st, launches);
    const float2* tw_sub1 = nullptr;
    const float2* tw_sub2 = nullptr;
    const float2* tw_full = nullptr;
    const int n1 = a.n / SPLIT_N2;
    cudaError_t e = get_twiddles(n1, &tw_sub1);
    if (e == cudaSuccess) e = get_twiddles(SPLIT_N2, &tw_sub2);
    if (e == cudaSuccess) e = get_full_twiddles(a.n, &tw_full);
    if (e != cudaSuccess) return e;
    // panel width: all pairs of one panel (npairs * n * width * 8 B) within the panel budget
    long long panel_mb = 256;
    if (const char* env = getenv("FDR_SPLIT_PANEL_MB"))
        if (atoi(env) > 0) panel_mb = atoi(env);
    int width = (int)((panel_mb << 20) / ((long long)a.n * 8));
    width = width / (a.npairs > 0 ? a.npairs : 1) / SPLIT_CWC * SPLIT_CWC;  // panel budget covers all pairs of the launch
    if (width < SPLIT_CWC) width = SPLIT_CWC;
    int count = 0;
    {
        for (int c0 = 0; c0 < a.pitch; c0 += width) {
            ColSplitArgs s{};
            s.n = a.n;
            s.pitch = a.pitch;
            s.col0 = c0;
            s.ncols = (a.pitch - c0 < width) ? (a.pitch - c0) : width;
            s.data = a.data;
            s.cplane = a.cplane;
            s.npairs = a.npairs;
            s.pair_base = a.pair_base;
            s.wiener = a.wiener;
            s.wiener_out = a.wiener_out;
            s.K = a.K;
            s.mode = a.mode;
            s.tw_full = tw_full;
            // A: strided forward + twiddle
            s.rows_valid = a.rows_valid;
            s.twiddle = 1;
            s.tw_sub = tw_sub1;
            e = (n1 == 128) ? launch_strided<7>(s, st) : launch_strided<6>(s, st);
            if (e != cudaSuccess) return e;
            // B: block forward (+ Wiener + inverse + twiddle)
            s.tw_sub = tw_sub2;
            e = launch_block(s, st);
            if (e != cudaSuccess) return e;
            count += 2;
            if (a.mode == COL_WIENER) {
                // A': strided, no twiddle, all rows
                s.rows_valid = a.n;
                s.twiddle = 0;
                s.tw_sub = tw_sub1;
                e = (n1 == 128) ? launch_strided<7>(s, st) : launch_strided<6>(s, st);
                if (e != cudaSuccess) return e;
                count += 1;
            }
        }
    }
    if (launches) *launches = count;
    return cudaSuccess;
}

}  // namespace fdr
