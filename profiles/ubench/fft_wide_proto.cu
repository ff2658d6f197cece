// Prototype: 2048-point row FFT with 64 points per thread (ONE warp per row, ONE shared-memory
// exchange, __syncwarp only) against the production 16-points-per-thread core (4 warps per row,
// two exchanges, __syncthreads).  Complex in -> complex out, same buffers, checked against each other.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../<pkg>/csrc -o fft_wide_proto fft_wide_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <type_traits>
#include "fft_core.cuh"

namespace fdr {
__host__ __device__ constexpr double sin_t(double x) {  // Taylor, |x| <= 2 pi
    double term = x, sum = x;
    for (int i = 1; i < 30; ++i) { term *= -x * x / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
__host__ __device__ constexpr double cos_t(double x) {
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 30; ++i) { term *= -x * x / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}
// a * exp(-2 pi i J / M), compile-time J, M
template <int M, int J> __device__ __forceinline__ float2 cmul_root(float2 a) {
    constexpr int j = ((J % M) + M) % M;
    if constexpr (j == 0) return a;
    else if constexpr (4 * j == M) return make_float2(a.y, -a.x);
    else if constexpr (2 * j == M) return make_float2(-a.x, -a.y);
    else if constexpr (4 * j == 3 * M) return make_float2(-a.y, a.x);
    else {
        constexpr double ang = 6.283185307179586476925286766559 * j / M;
        constexpr float wr = (float)cos_t(ang), wi = (float)(-sin_t(ang));
        return cmulc(a, wr, wi);
    }
}
// DFT of R = R1*R2 points in registers, natural order in and out.
template <int R1, int R2> __device__ __forceinline__ void dft_composite(float2* x) {
    constexpr int R = R1 * R2;
    static_for<0, R2>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        float2 y[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) y[n1] = x[R2 * n1 + n2];
        Dft<R1>::run(y);
        static_for<0, R1>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            x[R2 * k1 + n2] = cmul_root<R, n2 * k1>(y[k1]);
        });
    });
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) Dft<R2>::run(x + R2 * k1);
    float2 y[R];
#pragma unroll
    for (int q = 0; q < R; ++q) y[q] = x[R2 * (q % R1) + q / R1];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = y[q];
}

constexpr int WN = 2048, WE = 64, WT = WN / WE;  // 32 threads = one warp per row
__host__ __device__ constexpr int skew64(int idx) { return idx + (idx >> 6); }
constexpr int WEX = WN + WN / 64;  // float2 words per row buffer

template <int RPC> __global__ void __launch_bounds__(32 * RPC) row_wide_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int nrows) {
    extern __shared__ float2 smem2[];
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = blockIdx.x * RPC + w;
    if (row >= nrows) return;
    float2* ex = smem2 + w * WEX;
    const float2* src = in + (size_t)row * WN + t;
    float2 v[WE];
#pragma unroll
    for (int m = 0; m < WE; ++m) v[m] = src[WT * m];
    dft_composite<8, 8>(v);
    // exchange: out index of stage 1 = 64*t + q
    {
        float2* w0 = ex + skew64(64 * t);
#pragma unroll
        for (int q = 0; q < WE; ++q) w0[q] = v[q];
        __syncwarp();
        const float2* r0 = ex + t;
#pragma unroll
        for (int m = 0; m < WE; ++m) v[m] = r0[skew64(WT * m)];
    }
    // stage 2: NS = 64, R = 32, NB = 2: butterfly j = t + 32 b, inputs v[b + 2 r], twiddle exp(-2 pi i r j / 2048)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float2 x[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) x[r] = v[b + 2 * r];
        const float2* twb = tw + b * 31 * 32 + t;
#pragma unroll
        for (int r = 1; r < 32; ++r) x[r] = cmul(x[r], __ldg(twb + (r - 1) * 32));
        dft_composite<8, 4>(x);
#pragma unroll
        for (int r = 0; r < 32; ++r) v[b + 2 * r] = x[r];
    }
    float2* dst = out + (size_t)row * WN + t;
#pragma unroll
    for (int m = 0; m < WE; ++m) dst[WT * m] = v[m];
}

// production core, complex -> complex
__global__ void __launch_bounds__(128, 9) row_ref_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int nrows) {
    extern __shared__ float2 smem2[];
    constexpr int N = 2048, E = 16, T = 128;
    const int t = threadIdx.x, row = blockIdx.x;
    const float2* src = in + (size_t)row * N + t;
    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = src[T * m];
    fft_forward<N, 1>(v, smem2, tw, t, 0);
    float2* dst = out + (size_t)row * N + t;
#pragma unroll
    for (int m = 0; m < E; ++m) dst[T * m] = v[m];
}
// pass-3-like: split Re/Im into two real planes, optional min/max atomics
template <int ATOM> __global__ void __launch_bounds__(128, 9) row_split_kernel(const float2* __restrict__ in, float* __restrict__ out, const float2* __restrict__ tw, int nrows, unsigned* mm) {
    extern __shared__ float2 smem2[];
    constexpr int N = 2048, E = 16, T = 128;
    const int t = threadIdx.x, row = blockIdx.x % 2048, pair = blockIdx.x / 2048;
    const float2* src = in + ((size_t)pair * 2048 + row) * N + t;
    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = src[T * m];
    fft_forward<N, 1>(v, smem2, tw, t, 0);
    float* d0 = out + ((size_t)(2 * pair) * 2048 + row) * N + t;
    float* d1 = out + ((size_t)(2 * pair + 1) * 2048 + row) * N + t;
    float mn = 1e30f, mx = -1e30f;
#pragma unroll
    for (int m = 0; m < E; ++m) { d0[T * m] = v[m].x; d1[T * m] = v[m].y; mn = fminf(mn, fminf(v[m].x, v[m].y)); mx = fmaxf(mx, fmaxf(v[m].x, v[m].y)); }
    if (ATOM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if ((t & 31) == 0) {
            if (ATOM == 1) { atomicMin(mm + 4 * pair, __float_as_uint(mn)); atomicMax(mm + 4 * pair + 1, __float_as_uint(mx)); }
            if (ATOM == 2) { atomicMin(mm + 64 * (blockIdx.x % 997), __float_as_uint(mn)); atomicMax(mm + 64 * (blockIdx.x % 997) + 1, __float_as_uint(mx)); }
        }
    }
}
}  // namespace fdr

using namespace fdr;
template <class F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main(int argc, char** argv) {
    const int nrows = (argc > 1 ? atoi(argv[1]) : 6) * 2048;
    const size_t n = (size_t)nrows * WN;
    std::vector<float2> h(n);
    unsigned s = 12345u;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i].x = (s >> 8) * (1.f / 16777216.f); s = s * 1664525u + 1013904223u; h[i].y = (s >> 8) * (1.f / 16777216.f); }
    float2 *din, *d1, *d2, *tw16, *tw64;
    cudaMalloc(&din, n * 8); cudaMalloc(&d1, n * 8); cudaMalloc(&d2, n * 8);
    cudaMemcpy(din, h.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&tw16, sizeof(float2) * TwTotal<2048>::value);
    tw_fill_kernel<2048><<<(15 * 128 + 255) / 256, 256>>>(tw16);
    std::vector<float2> htw(2 * 31 * 32);
    for (int b = 0; b < 2; ++b) for (int r = 1; r < 32; ++r) for (int t = 0; t < 32; ++t) {
        double a = -2.0 * M_PI * r * (t + 32 * b) / 2048.0;
        htw[(b * 31 + (r - 1)) * 32 + t] = make_float2((float)cos(a), (float)sin(a));
    }
    cudaMalloc(&tw64, htw.size() * 8); cudaMemcpy(tw64, htw.data(), htw.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem_ref = fft_smem_bytes<2048, 1>();
    auto ref = [&] { row_ref_kernel<<<nrows, 128, smem_ref>>>(din, d1, tw16, nrows); };
    constexpr int RPC = 4;
    const size_t smem_w = (size_t)RPC * WEX * 8;
    cudaFuncSetAttribute(row_wide_kernel<RPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
    auto wide = [&] { row_wide_kernel<RPC><<<(nrows + RPC - 1) / RPC, 32 * RPC, smem_w>>>(din, d2, tw64, nrows); };
    float t_ref = time_ms(ref, 20), t_w = time_ms(wide, 20);
    unsigned* mm; cudaMalloc(&mm, 1 << 20); cudaMemset(mm, 0x7f, 1 << 20);
    float* dsplit = (float*)d2;
    auto s0 = [&] { row_split_kernel<0><<<nrows, 128, smem_ref>>>(din, dsplit, tw16, nrows, mm); };
    auto s1 = [&] { row_split_kernel<1><<<nrows, 128, smem_ref>>>(din, dsplit, tw16, nrows, mm); };
    auto s2 = [&] { row_split_kernel<2><<<nrows, 128, smem_ref>>>(din, dsplit, tw16, nrows, mm); };
    float ts0 = time_ms(s0, 20), ts1 = time_ms(s1, 20), ts2 = time_ms(s2, 20);
    printf("split planes: no atomics %.2f us/pair | same-address atomics %.2f | spread atomics %.2f | c2c %.2f\n", ts0 * 1e3 / (nrows / 2048), ts1 * 1e3 / (nrows / 2048), ts2 * 1e3 / (nrows / 2048), t_ref * 1e3 / (nrows / 2048));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    std::vector<float2> o1(n), o2(n);
    cudaMemcpy(o1.data(), d1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(o2.data(), d2, n * 8, cudaMemcpyDeviceToHost);
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) { double dx = o1[i].x - o2[i].x, dy = o1[i].y - o2[i].y; num += dx * dx + dy * dy; den += (double)o1[i].x * o1[i].x + (double)o1[i].y * o1[i].y; }
    printf("rel L2 wide vs production: %.3e\n", sqrt(num / den));
    printf("production 16-pt core: %.1f us (%.0f GB/s)   wide 64-pt core: %.1f us (%.0f GB/s)   [%d rows of 2048]\n", t_ref * 1e3, n * 16 / (t_ref * 1e-3) / 1e9, t_w * 1e3, n * 16 / (t_w * 1e-3) / 1e9, nrows);
    return 0;
}
