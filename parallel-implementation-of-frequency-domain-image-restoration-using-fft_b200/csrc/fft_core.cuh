// fft_core.cuh -- register-resident power-of-two FFT for sm_100a.
//
// One transform of length N is shared by T = N/E threads (E = 16 points per thread
// for N >= 16).  Thread t owns the points  t + T*m  (m = 0..E-1)  at the start of
// every stage and again at the end, so global loads/stores are coalesced and the
// result is in natural order (Stockham autosort, no bit reversal).  Between stages
// the points are exchanged through shared memory; each stage is one radix-16 (or
// the remaining 8/4/2) DFT done entirely in registers with compile-time inner
// twiddles.  Outer twiddles come from a per-length table laid out [stage][b][r][t]
// (one coalesced 8-byte load each, L1-resident), computed once per device in double
// precision -- no sincos or power products in the hot loop.
//
// Shared-memory layout of an exchange: float2 words, CW interleaved transforms, one skew group of CW
// words after every 16 points:  word(idx, c) = (idx + (idx >> 4)) * CW + c   (struct Skew below).
// For E = 16 this is bank-conflict free (64-bit accesses, half-warp phases) for the scattered stage
// writes and the strided reads at every N and every CW (simulated offline), and every
// access is thread base + compile-time offset: LDS/STS immediate offsets, no address arithmetic.
//
// Only the FORWARD transform (e^{-2 pi i nk/N}) is implemented; callers obtain the
// inverse as conj(FFT(conj(x))), folding the conjugations into their load/store.
//
// Replaces (does not port) the reference's one-butterfly-per-thread radix-2 kernel
// with log2(N) __syncthreads stages and global twiddle loads
// (/root/reference/fft/fft_gpu.cu:108-148).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace fdr {

// Blackwell (sm_100) packed fp32x2 arithmetic: one FADD2 per complex add/subtract.  The kernels are
// issue-slot bound, so halving the FP instruction count of the butterflies pays
// even though the FLOP rate of the pipe is unchanged.  -DFDR_NO_F32X2 falls back to scalar ops.
#if !defined(FDR_NO_F32X2)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }
// a * (wr + i*wi) with compile-time constants
__device__ __forceinline__ float2 cmulc(float2 a, float wr, float wi) {
    return make_float2(fmaf(-a.y, wi, a.x * wr), fmaf(a.y, wr, a.x * wi));
}

#define FDR_C1 0.92387953251128674f  // cos(pi/8)
#define FDR_S1 0.38268343236508977f  // sin(pi/8)
#define FDR_C2 0.70710678118654752f  // cos(pi/4)

// ---------------------------------------------------------------------------------
// Small forward DFTs on registers, natural order in, natural order out.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void dft2(float2& a0, float2& a1) {
    float2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    float2 t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

template <int R> struct Dft;

template <> struct Dft<1> {
    __device__ __forceinline__ static void run(float2*) {}
};
template <> struct Dft<2> {
    __device__ __forceinline__ static void run(float2* x) { dft2(x[0], x[1]); }
};
template <> struct Dft<4> {
    __device__ __forceinline__ static void run(float2* x) { dft4(x[0], x[1], x[2], x[3]); }
};
template <> struct Dft<8> {
    __device__ __forceinline__ static void run(float2* x) {
        // even / odd halves
        dft4(x[0], x[2], x[4], x[6]);
        dft4(x[1], x[3], x[5], x[7]);
        // odd outputs times W8^q
        float2 o1 = make_float2((x[3].x + x[3].y) * FDR_C2, (x[3].y - x[3].x) * FDR_C2);
        float2 o2 = cmul_mi(x[5]);
        float2 o3 = make_float2((x[7].y - x[7].x) * FDR_C2, -(x[7].x + x[7].y) * FDR_C2);
        float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1];
        x[0] = cadd(e0, o0);
        x[4] = csub(e0, o0);
        x[1] = cadd(e1, o1);
        x[5] = csub(e1, o1);
        x[2] = cadd(e2, o2);
        x[6] = csub(e2, o2);
        x[3] = cadd(e3, o3);
        x[7] = csub(e3, o3);
    }
};
template <> struct Dft<16> {
    __device__ __forceinline__ static void run(float2* x) {
        // n = 4*n1 + n2 ; k = k1 + 4*k2.  Step 1: DFT4 over n1 for each n2 (in place:
        // Y_{n2}[k1] lands in x[4*k1 + n2]).
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);
        // Step 2: Y_{n2}[k1] *= W16^{n2*k1}
        x[5] = cmulc(x[5], FDR_C1, -FDR_S1);    // k1=1 n2=1 : W^1
        x[6] = cmulc(x[6], FDR_C2, -FDR_C2);    // k1=1 n2=2 : W^2
        x[7] = cmulc(x[7], FDR_S1, -FDR_C1);    // k1=1 n2=3 : W^3
        x[9] = cmulc(x[9], FDR_C2, -FDR_C2);    // k1=2 n2=1 : W^2
        x[10] = cmul_mi(x[10]);                 // k1=2 n2=2 : W^4
        x[11] = cmulc(x[11], -FDR_C2, -FDR_C2); // k1=2 n2=3 : W^6
        x[13] = cmulc(x[13], FDR_S1, -FDR_C1);  // k1=3 n2=1 : W^3
        x[14] = cmulc(x[14], -FDR_C2, -FDR_C2); // k1=3 n2=2 : W^6
        x[15] = cmulc(x[15], -FDR_C1, FDR_S1);  // k1=3 n2=3 : W^9
        // Step 3: DFT4 over n2 for each k1 (in place: X[k1 + 4*k2] lands in x[4*k1 + k2]).
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
        // Un-digit-reverse: natural X[q] = x[4*(q%4) + q/4].  Pure register renaming.
        float2 y[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) y[q] = x[4 * (q & 3) + (q >> 2)];
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = y[q];
    }
};

// ---- compile-time roots of unity (constant twiddles inside the wide radices and derived table twiddles) ----
__host__ __device__ constexpr double wide_sin(double x) {  // Taylor series, |x| <= 2 pi, compile time only
    double term = x, sum = x;
    for (int i = 1; i < 32; ++i) {
        term *= -x * x / ((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
__host__ __device__ constexpr double wide_cos(double x) {
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 32; ++i) {
        term *= -x * x / ((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}

template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// a * exp(-2 pi i J / M) with compile-time J, M
template <int M, int J> __device__ __forceinline__ float2 cmul_root(float2 a) {
    constexpr int j = ((J % M) + M) % M;
    if constexpr (j == 0)
        return a;
    else if constexpr (4 * j == M)
        return make_float2(a.y, -a.x);
    else if constexpr (2 * j == M)
        return make_float2(-a.x, -a.y);
    else if constexpr (4 * j == 3 * M)
        return make_float2(-a.y, a.x);
    else {
        constexpr double ang = 6.283185307179586476925286766559 * j / M;
        constexpr float wr = (float)wide_cos(ang), wi = (float)(-wide_sin(ang));
        return cmulc(a, wr, wi);
    }
}

template <int N> struct FftGeom {
    static constexpr int E = (N >= 16) ? 16 : N;  // points per thread
    static constexpr int T = N / E;               // threads per transform
};

// ---------------------------------------------------------------------------------
// Twiddle table of one length N: for every stage after the first and every r = 1..R-1 the factors
// exp(-2 pi i * r * k / (NS*R)), k = (t + b*T) mod NS.  While NS <= T the index k = t mod NS does not
// depend on the butterfly b and only NS distinct entries exist per r (layout [r][k]: the two half-warps
// of a load coincide, one 128-byte wavefront).  For NS > T (always the last stage) only butterfly 0 is
// tabulated ([r][t]); butterfly b uses the same entries times the compile-time constant W_E^{r b}
// (k = t + b T and NS R = N, so exp(-2 pi i r b T / N) = W_16^{r b}): a few FP instructions instead of a
// second set of loads on the L1 data pipe, which is what bounds pass 1.  Offsets are compile-time.
// ---------------------------------------------------------------------------------
template <int N, int NS> struct TwStage {
    static constexpr int E = FftGeom<N>::E, T = FftGeom<N>::T;
    static constexpr int REM = N / NS;
    static constexpr int R = (REM < E) ? REM : E;
    static constexpr int NB = E / R;
    static constexpr bool SHARED_B = (NS <= T);          // k independent of b
    static constexpr int PER_R = SHARED_B ? NS : T;      // entries per (b, r)
    static constexpr int ENTRIES = (NS > 1) ? (R - 1) * PER_R : 0;
    // table offset of this stage = entries of all earlier stages
};
// All stages before the last have radix 16 (greedy radices), so the stage with sub-length NS is
// preceded by the stages with sub-lengths NS/16, NS/256, ...
template <int N, int NS, bool FIRST = (NS <= 16)> struct TwOffset {
    static constexpr int value = TwOffset<N, NS / 16>::value + TwStage<N, NS / 16>::ENTRIES;
};
template <int N, int NS> struct TwOffset<N, NS, true> {
    static constexpr int value = 0;
};
// Total entries for length N (stages have NS = 1, E, E^2, ... while NS < N).
template <int N, int NS = 1> struct TwTotal {
    static constexpr int R = TwStage<N, NS>::R;
    static constexpr int value = TwStage<N, NS>::ENTRIES + ((NS * R < N) ? TwTotal<N, (NS * R < N) ? NS * R : N>::value : 0);
};
template <int N> struct TwTotal<N, N> {
    static constexpr int value = 0;
};

// Exchange-buffer layout: one skew group (CW words) is inserted after every 16 points when a point
// group is narrower than a 128-byte bank row (CW < 16):  word(idx, c) = (idx + (idx >> 4)) * CW + c.
// Every access of a stage is then  base(thread) + compile-time offset  (LDS/STS immediate offsets, no
// per-access address arithmetic) and free of bank conflicts for the scattered writes and the strided
// reads at every N and CW (64-bit accesses, half-warp phases; simulated offline).
template <int CW> struct Skew {
    static constexpr bool ON = (CW < 16);
    __host__ __device__ static constexpr int f(int idx) { return ON ? idx + (idx >> 4) : idx; }
};
// float2 words of one exchange buffer holding CW interleaved transforms of length N
template <int N, int CW> __host__ __device__ constexpr int ex_words() { return Skew<CW>::f(N) * CW; }

// Barrier policies of the exchanges: the whole CTA, or a named barrier shared by one group of
// `threads` threads (several independent groups per CTA, col_tma.cu's pipelined kernel).
struct CtaBarrier {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct GroupBarrier {
    int id, threads;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
};

// One Stockham stage of radix R with sub-transform length NS already done.
template <int N, int R, int NS> __device__ __forceinline__ void stage_butterflies(float2* v, const float2* __restrict__ tw, int t) {
    constexpr int E = FftGeom<N>::E, T = FftGeom<N>::T, NB = E / R;
    static_for<0, NB>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        float2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = v[b + r * NB];
        if constexpr (NS > 1) {
            using St = TwStage<N, NS>;
            const float2* twb = tw + TwOffset<N, NS>::value + (St::SHARED_B ? (t & (NS - 1)) : t);
            if constexpr (St::SHARED_B || b == 0) {
#pragma unroll
                for (int r = 1; r < R; ++r) x[r] = cmul(x[r], __ldg(twb + (r - 1) * St::PER_R));
            } else {
                static_assert(b == 0 || (NS * R) % T == 0, "derived twiddles need an integral root order");
                constexpr int Q = (b == 0) ? 1 : NS * R / T;  // W_Q^{r b} = exp(-2 pi i r b T / (NS R))
                static_for<1, R>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    x[r] = cmul(cmul_root<Q, r * b>(x[r]), __ldg(twb + (r - 1) * St::PER_R));
                });
            }
        }
        Dft<R>::run(x);
#pragma unroll
        for (int r = 0; r < R; ++r) v[b + r * NB] = x[r];
    });
}

// Scatter the stage outputs to shared memory and gather the next stage's inputs.
// ex: this CTA's exchange buffer (N*CW float2); c = transform index in the tile.
// LEAD_SYNC = false when the caller guarantees nobody still reads `ex` (double-buffered exchanges).
template <int N, int CW, int R, int NS, bool LEAD_SYNC = true, class Bar = CtaBarrier>
__device__ __forceinline__ void stage_exchange(float2* v, float2* ex, int t, int c, const Bar& bar = Bar()) {
    constexpr int E = FftGeom<N>::E, T = FftGeom<N>::T, NB = E / R;
    static_assert(NS == 1 || NS % 16 == 0, "stage offsets must not carry into the skew term");
    if constexpr (LEAD_SYNC) bar.sync();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
        const int base = ((j & ~(NS - 1)) * R) + (j & (NS - 1));  // NS == 1: a multiple of 16 (R == 16), or N < 32
        float2* w0 = ex + Skew<CW>::f(base) * CW + c;
#pragma unroll
        for (int q = 0; q < R; ++q) w0[((NS == 1) ? q : Skew<CW>::f(q * NS)) * CW] = v[b + q * NB];
    }
    bar.sync();
    if constexpr (T % 16 == 0) {
        const float2* r0 = ex + Skew<CW>::f(t) * CW + c;
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = r0[Skew<CW>::f(T * m) * CW];
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = ex[Skew<CW>::f(t + T * m) * CW + c];
    }
}

// DB = true: `ex` holds TWO exchange buffers of N*CW float2 used alternately, which removes the
// write-after-read barrier of every exchange (one __syncthreads per exchange instead of two).
template <int N, int CW, int NS, bool DB = false, int XI = 0> struct FftStages {
    template <class Bar> __device__ __forceinline__ static void run(float2* v, float2* ex, const float2* __restrict__ tw, int t, int c, const Bar& bar) {
        constexpr int E = FftGeom<N>::E;
        constexpr int REM = N / NS;
        constexpr int R = (REM < E) ? REM : E;
        stage_butterflies<N, R, NS>(v, tw, t);
        if constexpr (NS * R < N) {
            if constexpr (DB)
                stage_exchange<N, CW, R, NS, false>(v, ex + (size_t)(XI & 1) * ex_words<N, CW>(), t, c, bar);
            else
                stage_exchange<N, CW, R, NS, true>(v, ex, t, c, bar);
            FftStages<N, CW, NS * R, DB, XI + 1>::run(v, ex, tw, t, c, bar);
        }
    }
};

// Forward FFT of length N over the E points held by thread t (points t + T*m).
// All T*CW threads of all transforms in the CTA must call this together when N > E
// (it contains __syncthreads).
template <int N, int CW, bool DB = false, class Bar = CtaBarrier>
__device__ __forceinline__ void fft_forward(float2* v, float2* ex, const float2* __restrict__ tw, int t, int c, const Bar& bar = Bar()) {
    if constexpr (N > 1) FftStages<N, CW, 1, DB>::run(v, ex, tw, t, c, bar);
}

// Shared memory (bytes) one CTA needs for `ntransforms` interleaved transforms of length N.
template <int N, int CW> constexpr size_t fft_smem_bytes() {
    return (N > FftGeom<N>::E) ? (size_t)ex_words<N, CW>() * sizeof(float2) : 0;
}

// Fills the twiddle table of length N (TwTotal<N>::value entries), double precision.
template <int N, int NS> __device__ __forceinline__ void tw_fill_stage(float2* tw, int i) {
    using St = TwStage<N, NS>;
    constexpr int R = St::R, PER_R = St::PER_R;
    if constexpr (NS > 1) {
        if (i < St::ENTRIES) {
            const int t = i % PER_R, rr = i / PER_R;  // butterfly 0 only (header comment)
            const int k = t & (NS - 1);
            double s, c;
            sincospi(2.0 * (double)((rr + 1) * k) / (double)(NS * R), &s, &c);
            tw[TwOffset<N, NS>::value + i] = make_float2((float)c, (float)(-s));
        }
    }
    if constexpr (NS * R < N) tw_fill_stage<N, NS * R>(tw, i);
}
template <int N> __global__ void tw_fill_kernel(float2* tw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (N > 16) tw_fill_stage<N, 1>(tw, i);
}

}  // namespace fdr
