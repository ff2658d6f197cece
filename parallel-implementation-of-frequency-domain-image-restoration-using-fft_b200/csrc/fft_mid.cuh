// fft_mid.cuh -- 32 points per thread for the LONG row transforms (N = 8192, 16384).
//
// A 16384-point complex row is half of an SM's register file, so exactly one row is in flight per SM and nothing but the
// row's own instruction-level parallelism hides its latencies.  With the 16-point core (fft_core.cuh) that row is 1024
// threads x 16 points, four stages (16.16.16.4), three shared-memory exchanges and six CTA-wide barriers: ncu showed 34 %
// issue activity, 28 % DRAM, all 32 warps waiting at the same barriers (profiles/r1b/ncu_kernels_16384_1gpu.txt).  Here the
// same row is 512 threads x 32 points: three stages (32.32.16; 8192 = 32.16.16 on 256 threads), TWO exchanges, four barriers,
// twice the independent work per thread between them, one third fewer shared-memory wavefronts and one twiddle stage less.
//
// Same conventions as the 16-point core: thread t owns points t + T*m (m < 32) before and after (natural order, coalesced
// global access), forward transform only, Stockham autosort.  Exchange layout: word(idx) = idx + (idx >> 5) -- one skew word
// per 32 points -- which keeps the strided stage writes and the consecutive reads free of bank conflicts (64-bit accesses,
// half-warp phases) and every access at thread base + compile-time offset.
//
// Replaces (does not port) /root/reference/fft/fft_gpu.cu:108-148, which cannot launch at all for rows of 8192 points or
// more (one thread per butterfly pair in a single block, fft_gpu.cu:219-221).
#pragma once
#include "fft_wide.cuh"

namespace fdr {

template <int N> struct MidGeom {
    static_assert(N == 8192 || N == 16384, "mid core: rows of 8192 or 16384 points");
    static constexpr int E = 32;
    static constexpr int T = N / E;                 // threads per transform: 256 or 512
    static constexpr int R1 = 32;                   // stage radices, N = R1 * R2 * R3
    static constexpr int R2 = (N == 16384) ? 32 : 16;
    static constexpr int R3 = 16;
    static constexpr int NB2 = E / R2, NB3 = E / R3;
    static constexpr int NS2 = R1, NS3 = R1 * R2;   // sub-transform lengths already done before stages 2 and 3
    static_assert(NB3 * T == NS3, "last stage: k = t + b T runs over exactly one sub-transform");
    // tables: stage 2 [r - 1][k], k = t mod 32 (independent of the butterfly);  stage 3 [r - 1][t], butterfly 0 only -- butterfly b
    // multiplies the same entry by the compile-time root W_32^{r b} (exp(-2 pi i r b T / N), T / N = 1/32)
    static constexpr int TW2 = (R2 - 1) * NS2;
    static constexpr int TW3 = (R3 - 1) * T;
    static constexpr int TW_ENTRIES = TW2 + TW3;
};
__host__ __device__ constexpr int mid_skew(int idx) { return idx + (idx >> 5); }
template <int N> __host__ __device__ constexpr int mid_ex_words() { return mid_skew(N); }

template <int N> __device__ __forceinline__ void fft_mid_forward(float2* v, float2* ex, const float2* __restrict__ tw, int t) {
    using G = MidGeom<N>;
    constexpr int E = G::E, T = G::T, R2 = G::R2, R3 = G::R3, NB2 = G::NB2, NB3 = G::NB3;
    // ---- stage 1: one radix-32 DFT over the thread's own 32 points (stride T), no twiddles ----
    dft_wide<32>(v);
    __syncthreads();  // whatever the caller last read from `ex`
    {
        float2* w0 = ex + 33 * t;  // outputs 32 t + q  ->  word 33 t + q
#pragma unroll
        for (int q = 0; q < E; ++q) w0[q] = v[q];
    }
    __syncthreads();
    {
        const float2* r0 = ex + mid_skew(t);  // T is a multiple of 32: word(t + T m) = word(t) + (T + T/32) m
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = r0[(T + T / 32) * m];
    }
    // ---- stage 2: radix R2, sub-length 32 ----
    static_for<0, NB2>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        float2 x[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) x[r] = v[b + r * NB2];
        const float2* twb = tw + (t & 31);
#pragma unroll
        for (int r = 1; r < R2; ++r) x[r] = cmul(x[r], __ldg(twb + (r - 1) * 32));
        dft_wide<R2>(x);
#pragma unroll
        for (int r = 0; r < R2; ++r) v[b + r * NB2] = x[r];
    });
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB2; ++b) {
        const int j = t + b * T;
        const int base = (j >> 5) * (32 * R2) + (j & 31);  // outputs base + 32 q  ->  word(base) + 33 q
        float2* w0 = ex + mid_skew(base);
#pragma unroll
        for (int q = 0; q < R2; ++q) w0[33 * q] = v[b + q * NB2];
    }
    __syncthreads();
    {
        const float2* r0 = ex + mid_skew(t);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = r0[(T + T / 32) * m];
    }
    // ---- stage 3: radix 16, sub-length 32 R2; k = t + b T ----
    static_for<0, NB3>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        float2 x[R3];
#pragma unroll
        for (int r = 0; r < R3; ++r) x[r] = v[b + r * NB3];
        const float2* twb = tw + G::TW2 + t;
        static_for<1, R3>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            x[r] = cmul(cmul_root<32, r * b>(x[r]), __ldg(twb + (r - 1) * T));
        });
        Dft<R3>::run(x);
#pragma unroll
        for (int r = 0; r < R3; ++r) v[b + r * NB3] = x[r];
    });
}

template <int N> __global__ void mid_tw_fill_kernel(float2* tw) {
    using G = MidGeom<N>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G::TW_ENTRIES) return;
    double s, c;
    if (i < G::TW2) {
        const int k = i % 32, r = i / 32 + 1;
        sincospi(2.0 * (double)(r * k) / (double)(G::NS2 * G::R2), &s, &c);
    } else {
        const int ii = i - G::TW2, t = ii % G::T, r = ii / G::T + 1;
        sincospi(2.0 * (double)(r * t) / (double)N, &s, &c);
    }
    tw[i] = make_float2((float)c, (float)(-s));
}

}  // namespace fdr
