// col_split.cuh -- column pass for LONG columns (N >= 8192), four-step inside the pass.
//
// A column of N = N1 * 128 points does not fit a CTA's registers/shared memory at a useful tile
// width (16384 x 8 B = 128 KB per column), so the monolithic column kernel degenerates to one
// column per CTA and 8-byte strided accesses.  Here the transform is split as
//     n = n2 + 128*n1,  k = k1 + N1*k2      (n2, k2 < 128;  n1, k1 < N1 = N/128)
//   A  (strided):  for each n2: DFT_N1 over n1, times W_N^{n2*k1}, stored in place at row n2 + 128*k1
//   B  (block):    for each k1: DFT_128 over the CONTIGUOUS rows 128*k1 + n2  -> X[k1 + N1*k2] at
//                  row 128*k1 + k2 (digit-swapped order); times the Wiener factor (kept in the same
//                  order); conj; DFT_128 again (= inverse); times W_N^{n2*k1}; stored in place
//   A' (strided):  for each n2: DFT_N1 over k1 (no twiddle) -> conj of the inverse column transform,
//                  natural order, exactly what the monolithic COL_WIENER kernel leaves behind.
// Every access is 16 columns x 8 B = 128 contiguous bytes per row; tiles are 32 KB, 256 threads.
// The three kernels run panel by panel.  Measured on B200 (profiles/): 32 MB panels stay L2-resident
// (ncu: 0.2 MB of DRAM reads per strided launch) but the small grids are latency-bound (16 us per
// 64 MB of L2 traffic); 256 MB panels are 15 % faster end to end and are the default.
#pragma once
#include "fft_core.cuh"
#include "passes.h"

namespace fdr {

constexpr int SPLIT_N2 = 128;   // contiguous block length
constexpr int SPLIT_CWC = 16;   // columns per tile (128 bytes per row)
constexpr int SPLIT_NJ = 2;     // n2 (or k1) values per CTA -> 32 interleaved transforms

struct ColSplitArgs {
    int n;               // column length
    int pitch;           // elements per row
    int col0, ncols;     // panel [col0, col0 + ncols), multiples of 16
    int npairs, pair_base;
    int rows_valid;      // kernel A: rows >= rows_valid read as zero
    int twiddle;         // strided kernel: multiply by W_N^{n2*k1} after the transform
    int mode;            // block kernel: COL_WIENER or COL_MAKE_WIENER
    float2* data;        // pair p at data + p*cplane (in place); grid.z = pair
    long long cplane;
    const float2* wiener;  // digit-swapped order
    float2* wiener_out;
    float K;
    const float2* tw_sub;   // twiddle table of the sub-transform length (get_twiddles)
    const float2* tw_full;  // exp(-2 pi i j / N), j < N
};

// grid = (ncols/16, 128/NJ), block = (N1/16) * 32 threads
template <int LOGN1>
__global__ void __launch_bounds__((1 << LOGN1) / 16 * SPLIT_CWC * SPLIT_NJ) col_split_strided_kernel(const ColSplitArgs a) {
    constexpr int N1 = 1 << LOGN1, E = FftGeom<N1>::E, T = FftGeom<N1>::T, CW = SPLIT_CWC * SPLIT_NJ;
    extern __shared__ float2 smem2[];
    const int tid = threadIdx.x;
    const int cc = tid % CW, t = tid / CW;
    const int c = cc % SPLIT_CWC, j = cc / SPLIT_CWC;
    const int n2 = blockIdx.y * SPLIT_NJ + j;
    const int col = a.col0 + blockIdx.x * SPLIT_CWC + c;
    float2* base = a.data + (long long)(blockIdx.z + a.pair_base) * a.cplane + (long long)n2 * a.pitch + col;
    const long long rstride = (long long)SPLIT_N2 * a.pitch;  // rows n2 + 128*k
    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int k = t + T * m;
        v[m] = (n2 + SPLIT_N2 * k < a.rows_valid) ? base[k * rstride] : make_float2(0.f, 0.f);
    }
    fft_forward<N1, CW>(v, smem2, a.tw_sub, t, cc);
    if (a.twiddle) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cmul(v[m], __ldg(a.tw_full + n2 * (t + T * m)));
    }
#pragma unroll
    for (int m = 0; m < E; ++m) base[(t + T * m) * rstride] = v[m];
}

// grid = (ncols/16, N1/NJ), block = 8 * 32 = 256 threads
template <int MODE>
__global__ void __launch_bounds__(SPLIT_N2 / 16 * SPLIT_CWC * SPLIT_NJ, 4) col_split_block_kernel(const ColSplitArgs a) {
    constexpr int N2 = SPLIT_N2, E = FftGeom<N2>::E, T = FftGeom<N2>::T, CW = SPLIT_CWC * SPLIT_NJ;
    extern __shared__ float2 smem2[];
    const int tid = threadIdx.x;
    const int cc = tid % CW, t = tid / CW;
    const int c = cc % SPLIT_CWC, j = cc / SPLIT_CWC;
    const int k1 = blockIdx.y * SPLIT_NJ + j;
    const int col = a.col0 + blockIdx.x * SPLIT_CWC + c;
    const long long first = ((long long)k1 * N2 + t) * a.pitch + col;
    const long long rstride = (long long)T * a.pitch;
    float2* base = a.data + (long long)(blockIdx.z + a.pair_base) * a.cplane + first;
    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = base[m * rstride];
    fft_forward<N2, CW>(v, smem2, a.tw_sub, t, cc);
    if constexpr (MODE == COL_MAKE_WIENER) {
        float2* wo = a.wiener_out + first;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float hr = v[m].x, hi = v[m].y;
            const float denom = fmaf(hr, hr, hi * hi) + a.K;
            wo[m * rstride] = make_float2(hr / denom, -hi / denom);
        }
    } else {
        const float2* wf = a.wiener + first;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float2 y = cmul(v[m], __ldg(wf + m * rstride));
            v[m] = make_float2(y.x, -y.y);
        }
        fft_forward<N2, CW>(v, smem2, a.tw_sub, t, cc);
#pragma unroll
        for (int m = 0; m < E; ++m) base[m * rstride] = cmul(v[m], __ldg(a.tw_full + (t + T * m) * k1));
    }
}

}  // namespace fdr
