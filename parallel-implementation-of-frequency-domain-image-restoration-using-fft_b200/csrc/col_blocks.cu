// col_blocks.cu -- column pass for LONG columns (N = 8192, 16384) as K x 2048:  n = 2048*n1 + n2,
// k = k1 + K*k2  (K = N/2048 = 4 or 8).
//   S1  for every (n2, column): DFT_K over the K rows 2048*n1 + n2, times W_N^{n2 k1}, in place at row
//       2048*k1 + n2.  A streaming pass over full-width rows: every access is a contiguous row segment.
//   S2  every block of 2048 rows [2048*k1, 2048*k1 + 2048) is an ordinary 2048-point column problem: the
//       wide TMA kernel (col_wide.cu) does DFT_2048 over n2 -> X[k1 + K*k2] at row 2048*k1 + k2, times the
//       Wiener factor (stored in that same block order), conj, DFT_2048 again (= inverse over k2).
//   S3  for every (n2, column): times W_N^{n2 k1}, DFT_K over k1 -> conj of the inverse transform at
//       row 2048*n1 + n2: natural order, what the monolithic COL_WIENER kernel leaves behind.
// Replaces the 128 x 128 four-step pass (col_split.cu) as the default for these lengths: the same 56 B per
// complex pixel of traffic, but S1/S3 stream at HBM speed and S2 is the kernel that already runs 2048-row
// planes at 0.87-0.89 of the HBM roofline; col_split's three kernels ran at 1.4 TB/s-equivalent.
// The reference has no counterpart (its shared-memory kernel stops at N = 4096, fft_gpu.cu:219-221).
#include <cstdint>
#include <cstdlib>

#include "fft_core.cuh"
#include "passes.h"

namespace fdr {

constexpr int BLK_M = 2048;

// full-length twiddle table exp(-2 pi i j / n), j < n (col_split.cu)
cudaError_t get_full_twiddles(int n, const float2** out);

__device__ __forceinline__ float4 cmul4(float4 a, float2 w) {  // two complex numbers times w
    return make_float4(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.y, w.x, a.x * w.y), fmaf(-a.w, w.y, a.z * w.x), fmaf(a.w, w.x, a.z * w.y));
}

// grid = (pitch/2/128, 2048, npairs), block = 128: thread = two adjacent columns of one n2.
// PRE = false: DFT_K then twiddle (S1); PRE = true: twiddle then DFT_K (S3).
template <int K, bool PRE> __global__ void __launch_bounds__(128) radixk_rows_kernel(float2* data, long long cplane, int pitch, int pair_base, int rows_valid, const float2* __restrict__ tw_full) {
    const int x2 = blockIdx.x * 128 + threadIdx.x;  // column pair
    if (2 * x2 >= pitch) return;
    const int n2 = blockIdx.y;
    float4* base = reinterpret_cast<float4*>(data + (long long)(blockIdx.z + pair_base) * cplane + (long long)n2 * pitch) + x2;
    const long long rstride = (long long)BLK_M * pitch / 2;  // float4 elements between row blocks
    float2 a[K], b[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (BLK_M * j + n2 < rows_valid) v = base[j * rstride];
        if constexpr (PRE) {
            if (j > 0) v = cmul4(v, __ldg(tw_full + n2 * j));
        }
        a[j] = make_float2(v.x, v.y);
        b[j] = make_float2(v.z, v.w);
    }
    Dft<K>::run(a);
    Dft<K>::run(b);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        float4 v = make_float4(a[j].x, a[j].y, b[j].x, b[j].y);
        if constexpr (!PRE) {
            if (j > 0) v = cmul4(v, __ldg(tw_full + n2 * j));
        }
        base[j * rstride] = v;
    }
}

template <int K, bool PRE> static cudaError_t launch_radixk(const ColPassArgs& a, int rows_valid, const float2* tw_full, cudaStream_t s) {
    dim3 grid((a.pitch / 2 + 127) / 128, BLK_M, a.npairs);
    radixk_rows_kernel<K, PRE><<<grid, 128, 0, s>>>(a.data, a.cplane, a.pitch, a.pair_base, rows_valid, tw_full);
    return cudaGetLastError();
}

bool col_blocks_applicable(const ColPassArgs& a) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* env = getenv("FDR_COL_BLOCKS");
        enabled = (env && atoi(env) == 0) ? 0 : 1;
    }
    if (!enabled) return false;
    if (a.n != 8192 && a.n != 16384) return false;
    if (a.mode != COL_WIENER && a.mode != COL_MAKE_WIENER) return false;
    if (a.conj || a.pitch % 4 != 0) return false;
    // every 2048-row block must be a legal plane of the wide TMA kernel (pointer alignment is checked at launch)
    ColPassArgs sub{};
    sub.n = BLK_M;
    sub.pitch = a.pitch;
    sub.mode = COL_WIENER;
    return col_tma_geometry_ok(BLK_M, a.pitch, (long long)BLK_M * a.pitch) && col_wide_applicable(sub);
}

cudaError_t launch_col_blocks(const ColPassArgs& a, cudaStream_t s, int* launches) {
    const int K = a.n / BLK_M;
    const float2* tw_full = nullptr;
    cudaError_t e = get_full_twiddles(a.n, &tw_full);
    if (e != cudaSuccess) return e;
    if (a.cplane != (long long)a.n * a.pitch) return cudaErrorInvalidValue;
    if (reinterpret_cast<uintptr_t>(a.data) & 15) return cudaErrorInvalidValue;  // TMA tiles and float4 rows
    if (a.mode == COL_WIENER && (reinterpret_cast<uintptr_t>(a.wiener) & 15)) return cudaErrorInvalidValue;
    int count = 0;
    // S1
    e = (K == 8) ? launch_radixk<8, false>(a, a.rows_valid, tw_full, s) : launch_radixk<4, false>(a, a.rows_valid, tw_full, s);
    if (e != cudaSuccess) return e;
    ++count;
    if (a.mode == COL_MAKE_WIENER) {
        // one 2048-point column pass per row block; the factor lands in block order
        for (int kb = 0; kb < K; ++kb) {
            ColPassArgs c = a;
            c.n = BLK_M;
            c.npairs = 1;
            c.rows_valid = BLK_M;
            c.data = a.data + (long long)a.pair_base * a.cplane + (long long)kb * BLK_M * a.pitch;
            c.pair_base = 0;
            c.cplane = (long long)BLK_M * a.pitch;
            c.wiener_out = a.wiener_out + (long long)kb * BLK_M * a.pitch;
            e = get_twiddles(BLK_M, &c.tw);
            if (e != cudaSuccess) return e;
            e = launch_col_pass(c, s);
            if (e != cudaSuccess) return e;
            ++count;
        }
        if (launches) *launches = count;
        return cudaSuccess;
    }
    // S2: every (pair, row block) is one 2048-row plane of the wide kernel
    {
        ColPassArgs c = a;
        c.n = BLK_M;
        c.npairs = a.npairs * K;
        c.pair_base = a.pair_base * K;
        c.cplane = (long long)BLK_M * a.pitch;
        c.rows_valid = BLK_M;
        c.wiener_blocks = K;
        c.col_variant = 0;
        e = launch_col_wiener_wide(c, s);
        if (e != cudaSuccess) return e;
        ++count;
    }
    // S3
    e = (K == 8) ? launch_radixk<8, true>(a, a.n, tw_full, s) : launch_radixk<4, true>(a, a.n, tw_full, s);
    if (e != cudaSuccess) return e;
    ++count;
    if (launches) *launches = count;
    return cudaSuccess;
}

}  // namespace fdr
