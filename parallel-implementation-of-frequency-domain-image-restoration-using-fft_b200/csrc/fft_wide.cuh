// fft_wide.cuh -- 64 points per thread: the transform needs ONE shared-memory exchange.
//
// The 16-point core (fft_core.cuh) runs a 2048-point transform as 16 x 16 x 8 with two exchanges;
// ncu shows the column pass spends 65-70 % of the shared-memory / L1 data pipe on those exchanges
// (4 per tile) plus the three tile stagings.  Here a transform of N = 64 * R2 points (R2 = 16, 32, 64)
// is shared by T = N/64 threads: stage 1 is a radix-64 DFT in registers (8 x 8, compile-time
// twiddles), then one exchange, then 64/R2 radix-R2 butterflies with table twiddles -- half the
// exchange traffic, ~45 % fewer table loads and fewer FP instructions per point, paid for with
// ~170 registers per thread (12 warps per SM, but 64 independent points of ILP per thread).
// Used by the column pass (col_wide.cu), where the load/store pipe, not HBM, is the bound; the row
// passes run at the HBM roofline with the 16-point core and keep it (profiles/ubench/fft_wide_proto.cu).
//
// Thread t of a transform owns points t + T*m (m < 64) before and after, natural order, like the
// 16-point core.  Exchange layout: word(idx, c) = (idx + (idx >> 6)) * CW + c (one skew group per 64
// points): conflict-free for the 64-consecutive writes of a thread and the strided reads.
//
// Replaces (does not port) /root/reference/fft/fft_gpu.cu:108-148.
#pragma once
#include "fft_core.cuh"

namespace fdr {

// Forward DFT of R = R1 * R2 points in registers, natural order in and out:
// n = R2*n1 + n2, k = k1 + R1*k2; DFT_R1 over n1, twiddle W_R^{n2 k1}, DFT_R2 over n2.
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
template <int R> __device__ __forceinline__ void dft_wide(float2* x) {
    if constexpr (R == 64)
        dft_composite<8, 8>(x);
    else if constexpr (R == 32)
        dft_composite<8, 4>(x);
    else
        Dft<R>::run(x);
}

template <int N> struct WideGeom {
    static constexpr int E = 64;
    static constexpr int T = N / E;    // threads per transform
    static constexpr int R2 = N / E;   // radix of the second stage (16, 32 or 64)
    static constexpr int NB = E / R2;  // second-stage butterflies per thread
    static_assert(N >= 1024 && N <= 4096, "wide core: 1024, 2048 or 4096 points");
    static constexpr int TW_ENTRIES = NB * (R2 - 1) * T;  // table [b][r - 1][t] = exp(-2 pi i r (t + b T) / N)
};
__host__ __device__ constexpr int wide_skew(int idx) { return idx + (idx >> 6); }
template <int N, int CW> __host__ __device__ constexpr int wide_ex_words() { return wide_skew(N) * CW; }

// Forward FFT over the 64 points of thread t (points t + T*m) of transform c; all threads of the CTA
// that share `ex` call it together.  The first barrier protects whatever the caller last read from `ex`.
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `idle` runs right after the exchange reads, behind a barrier when it is not NoHook: from there on `ex` is
// free while the second-stage butterflies still compute -- the column kernel starts its next TMA load there.
template <int N, int CW, class Bar = CtaBarrier, class Idle = NoHook>
__device__ __forceinline__ void fft_wide_forward(float2* v, float2* ex, const float2* __restrict__ tw, int t, int c, const Bar& bar = Bar(),
                                                 const Idle& idle = Idle()) {
    using G = WideGeom<N>;
    constexpr int E = G::E, T = G::T, R2 = G::R2, NB = G::NB;
    dft_wide<64>(v);
    bar.sync();
    {
        float2* w0 = ex + wide_skew(64 * t) * CW + c;  // stage-1 outputs of thread t: positions 64 t + q
#pragma unroll
        for (int q = 0; q < E; ++q) w0[q * CW] = v[q];
    }
    bar.sync();
    {
        const float2* r0 = ex + t * CW + c;  // t < T <= 64
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = r0[wide_skew(T * m) * CW];
    }
    if constexpr (!std::is_same<Idle, NoHook>::value) {
        bar.sync();
        idle();
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float2 x[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) x[r] = v[b + NB * r];
        const float2* twb = tw + b * (R2 - 1) * T + t;
#pragma unroll
        for (int r = 1; r < R2; ++r) x[r] = cmul(x[r], __ldg(twb + (r - 1) * T));
        dft_wide<R2>(x);
#pragma unroll
        for (int r = 0; r < R2; ++r) v[b + NB * r] = x[r];
    }
}

template <int N> __global__ void wide_tw_fill_kernel(float2* tw) {
    using G = WideGeom<N>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G::TW_ENTRIES) return;
    const int t = i % G::T, rr = (i / G::T) % (G::R2 - 1), b = i / (G::T * (G::R2 - 1));
    double s, c;
    sincospi(2.0 * (double)((rr + 1) * (t + b * G::T)) / (double)N, &s, &c);
    tw[i] = make_float2((float)c, (float)(-s));
}

}  // namespace fdr
