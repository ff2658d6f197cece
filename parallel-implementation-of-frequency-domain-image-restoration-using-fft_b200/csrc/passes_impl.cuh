// passes_impl.cuh -- the two fused FFT pass kernels (rows, columns), templated on log2(N).
// Included by passes_g*.cu, each of which instantiates a group of sizes so the groups
// compile in parallel.
#pragma once
#include "fft_core.cuh"
#include "passes.h"

namespace fdr {

__device__ __forceinline__ unsigned int f32_ordered(float f) {
    unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---------------------------------------------------------------------------------
// Row pass.  One transform of length N per T = N/16 threads; RPC rows per CTA so that
// a CTA has at least 128 threads.  grid = (ceil(nrows/RPC), npairs).
// ---------------------------------------------------------------------------------
template <int LOGN> struct RowGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int E = FftGeom<N>::E;
    static constexpr int T = FftGeom<N>::T;
    static constexpr int RPC = (T >= 128) ? 1 : (128 / T);
    static constexpr int THREADS = T * RPC;
    static constexpr size_t SMEM = fft_smem_bytes<N>(RPC);
};

template <int LOGN>
__global__ void __launch_bounds__(RowGeom<LOGN>::THREADS) row_pass_kernel(const RowPassArgs a) {
    using Gm = RowGeom<LOGN>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T, RPC = Gm::RPC;
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    const int rl = (RPC > 1) ? (tid / T) : 0;
    const int t = (RPC > 1) ? (tid % T) : tid;
    const int row = blockIdx.x * RPC + rl;
    const int pair = blockIdx.y;
    const bool active = row < a.nrows;
    float* sre = smem + (size_t)rl * 2 * N;
    float* sim = sre + N;

    const long long u0 = 2LL * pair, u1 = u0 + 1;  // local units
    const bool has1 = (a.unit_base + u1) < a.units_total;

    float2 v[E];
    if (a.in_mode == ROW_IN_COMPLEX) {
        const float2* src = a.cin + (long long)pair * a.cplane + (long long)row * N;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 z = make_float2(0.f, 0.f);
            if (active) z = src[t + T * m];
            if (a.conj_in) z.y = -z.y;
            v[m] = z;
        }
    } else if (a.in_mode == ROW_IN_PAIR_F32) {
        const float* p0 = a.in_f32 + (a.unit_base + u0) * a.in_unit_stride + (long long)row * a.in_row_stride;
        const float* p1 = a.in_f32 + (a.unit_base + u1) * a.in_unit_stride + (long long)row * a.in_row_stride;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int x = t + T * m;
            float2 z = make_float2(0.f, 0.f);
            if (active && x < a.img_cols) {
                z.x = __ldg(p0 + x);
                if (has1) z.y = __ldg(p1 + x);
            }
            v[m] = z;
        }
    } else {  // ROW_IN_PAIR_U8 : x * (float)(1/255.)  (serial.cpp:24-25 convertTo + /= 255.0)
        const long long g0 = a.unit_base + u0, g1 = a.unit_base + u1;
        const int C = a.channels;
        const long long i0 = g0 / C, i1 = g1 / C;
        const int c0 = (int)(g0 - i0 * C), c1 = (int)(g1 - i1 * C);
        const uint8_t* p0 = a.in_u8 + ((i0 * a.img_rows + row) * (long long)a.img_cols) * C + c0;
        const uint8_t* p1 = a.in_u8 + ((i1 * a.img_rows + row) * (long long)a.img_cols) * C + c1;
        const float inv255 = (float)(1.0 / 255.0);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int x = t + T * m;
            float2 z = make_float2(0.f, 0.f);
            if (active && x < a.img_cols) {
                z.x = (float)__ldg(p0 + (long long)x * C) * inv255;
                if (has1) z.y = (float)__ldg(p1 + (long long)x * C) * inv255;
            }
            v[m] = z;
        }
    }

    fft_forward<N, 1>(v, sre, sim, t, 0);

    if (a.out_mode == ROW_OUT_COMPLEX) {
        if (active) {
            float2* dst = a.cout + (long long)pair * a.cplane + (long long)row * N;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                float2 z = v[m];
                if (a.conj_out) z.y = -z.y;
                dst[t + T * m] = z;
            }
        }
    } else {
        // inverse row pass of the restoration: input was conj(column result), so the restored
        // pair is conj(v): plane a = v.x, plane b = -v.y.
        float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
        if (active) {
            const bool store_row = row < a.raw_rows;
            float* d0 = a.raw + u0 * a.raw_unit_stride + (long long)row * a.raw_cols;
            float* d1 = a.raw + u1 * a.raw_unit_stride + (long long)row * a.raw_cols;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int x = t + T * m;
                const float ra = v[m].x, rb = -v[m].y;
                mn0 = fminf(mn0, ra);
                mx0 = fmaxf(mx0, ra);
                mn1 = fminf(mn1, rb);
                mx1 = fmaxf(mx1, rb);
                if (store_row && x < a.raw_cols) {
                    d0[x] = ra;
                    if (has1) d1[x] = rb;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
            mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        }
        __shared__ float red[32][4];
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int NW = (Gm::THREADS + 31) / 32;
        if (lane == 0) {
            red[warp][0] = mn0;
            red[warp][1] = mx0;
            red[warp][2] = mn1;
            red[warp][3] = mx1;
        }
        __syncthreads();
        if (warp == 0) {
            mn0 = (lane < NW) ? red[lane][0] : INFINITY;
            mx0 = (lane < NW) ? red[lane][1] : -INFINITY;
            mn1 = (lane < NW) ? red[lane][2] : INFINITY;
            mx1 = (lane < NW) ? red[lane][3] : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
                mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
            }
            if (lane == 0) {
                atomicMin(a.minmax + 2 * u0, f32_ordered(mn0));
                atomicMax(a.minmax + 2 * u0 + 1, f32_ordered(mx0));
                if (has1) {
                    atomicMin(a.minmax + 2 * u1, f32_ordered(mn1));
                    atomicMax(a.minmax + 2 * u1 + 1, f32_ordered(mx1));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// Column pass.  A CTA owns CW adjacent columns of one pair: thread (c, t) holds rows
// t + T*m of column c0 + c.  Lanes vary c fastest so a warp touches 32/CW rows x
// (CW*8) contiguous bytes.  In place.  grid = (pitch/CW, npairs).
//   COL_FFT          forward FFT (or inverse through conj_in/conj_out)
//   COL_WIENER       FFT, z = conj(X * Wf), FFT again: the stored value is
//                    conj(IFFT_y(X*Wf)), which pass 3 consumes without a load conjugate
//   COL_MAKE_WIENER  FFT of the PSF's row spectrum, store Wf = conj(H)/(|H|^2+K)
//   COL_FILTER       FFT, store X * Wf (the filtered spectrum F; parity gate only)
// ---------------------------------------------------------------------------------
template <int LOGN, int CW> struct ColGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int E = FftGeom<N>::E;
    static constexpr int T = FftGeom<N>::T;
    static constexpr int THREADS = T * CW;
    static constexpr size_t SMEM = fft_smem_bytes<N>(CW);
};

template <int LOGN, int CW>
__global__ void __launch_bounds__(ColGeom<LOGN, CW>::THREADS) col_pass_kernel(const ColPassArgs a) {
    using Gm = ColGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T;
    extern __shared__ float smem[];
    float* sre = smem;
    float* sim = smem + (size_t)N * CW;
    const int tid = threadIdx.x;
    const int c = tid % CW, t = tid / CW;
    const int col = blockIdx.x * CW + c;
    const bool active = col < a.pitch;
    float2* base = a.data + (long long)blockIdx.y * a.cplane + col;

    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int r = t + T * m;
        float2 z = make_float2(0.f, 0.f);
        if (active && r < a.rows_valid) z = base[(long long)r * a.pitch];
        if (a.conj_in) z.y = -z.y;
        v[m] = z;
    }

    fft_forward<N, CW>(v, sre, sim, t, c);

    if (a.mode == COL_WIENER) {
        const float2* wf = a.wiener + col;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int r = t + T * m;
            float2 w = make_float2(0.f, 0.f);
            if (active) w = __ldg(wf + (long long)r * a.pitch);
            const float2 y = cmul(v[m], w);
            v[m] = make_float2(y.x, -y.y);
        }
        fft_forward<N, CW>(v, sre, sim, t, c);
    }

    if (a.mode == COL_FILTER) {
        const float2* wf = a.wiener + col;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 w = make_float2(0.f, 0.f);
            if (active) w = __ldg(wf + (long long)(t + T * m) * a.pitch);
            v[m] = cmul(v[m], w);
        }
    }

    if (!active) return;
    if (a.mode == COL_MAKE_WIENER) {
        float2* wo = a.wiener_out + col;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int r = t + T * m;
            const float hr = v[m].x, hi = v[m].y;
            const float denom = fmaf(hr, hr, hi * hi) + a.K;  // fft_serial.cpp:195-197
            wo[(long long)r * a.pitch] = make_float2(hr / denom, -hi / denom);
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int r = t + T * m;
            float2 z = v[m];
            if (a.conj_out) z.y = -z.y;
            base[(long long)r * a.pitch] = z;
        }
    }
}

template <int LOGN> cudaError_t launch_row_pass_t(const RowPassArgs& a, cudaStream_t s) {
    using Gm = RowGeom<LOGN>;
    dim3 grid((a.nrows + Gm::RPC - 1) / Gm::RPC, a.npairs);
    row_pass_kernel<LOGN><<<grid, Gm::THREADS, Gm::SMEM, s>>>(a);
    return cudaGetLastError();
}
template <int LOGN> cudaError_t configure_row_pass_t() {
    return cudaFuncSetAttribute(row_pass_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)RowGeom<LOGN>::SMEM);
}
template <int LOGN, int CW> cudaError_t launch_col_pass_t(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColGeom<LOGN, CW>;
    dim3 grid((a.pitch + CW - 1) / CW, a.npairs);
    col_pass_kernel<LOGN, CW><<<grid, Gm::THREADS, Gm::SMEM, s>>>(a);
    return cudaGetLastError();
}
template <int LOGN, int CW> cudaError_t configure_col_pass_t() {
    return cudaFuncSetAttribute(col_pass_kernel<LOGN, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ColGeom<LOGN, CW>::SMEM);
}

// Default tile width per length: keep T*CW <= 512 threads, CW*8 B >= 32 B where possible.
constexpr int default_col_cw(int logn) {
    return logn <= 6 ? 32 : logn <= 8 ? 16 : logn <= 10 ? 8 : logn == 11 ? 4 : logn == 12 ? 2 : 1;
}

}  // namespace fdr
