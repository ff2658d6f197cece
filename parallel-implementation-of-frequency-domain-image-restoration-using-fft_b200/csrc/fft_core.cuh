// fft_core.cuh -- register-resident power-of-two FFT for sm_100a.
//
// One transform of length N is shared by T = N/E threads (E = 16 points per thread
// for N >= 16).  Thread t owns the points  t + T*m  (m = 0..E-1)  at the start of
// every stage and again at the end, so global loads/stores are coalesced and the
// result is in natural order (Stockham autosort, no bit reversal).  Between stages
// the points are exchanged through shared memory; each stage is one radix-16 (or
// the remaining 8/4/2) DFT done entirely in registers with compile-time inner
// twiddles and one sincospi per thread for the outer twiddle.
//
// Shared-memory layout of an exchange (split re/im planes, CW interleaved
// transforms):  word(idx, c) = ((idx ^ ((idx >> 4) & (32/CW - 1))) * CW + c).
// For E = 16 this is bank-conflict free for the scattered stage writes and the
// strided reads at every N and every CW in {1,2,4,8,16,32} (simulated offline,
// see DESIGN.md).
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

namespace fdr {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ float2 csqr(float2 a) {
    return make_float2(fmaf(a.x, a.x, -(a.y * a.y)), (a.x + a.x) * a.y);
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

// x[r] *= w^r, r = 1..R-1, powers built by a product tree of depth <= 4.
template <int R> __device__ __forceinline__ void apply_twiddle_powers(float2* x, float2 w) {
    if constexpr (R >= 2) x[1] = cmul(x[1], w);
    if constexpr (R >= 4) {
        float2 w2 = csqr(w);
        float2 w3 = cmul(w2, w);
        x[2] = cmul(x[2], w2);
        x[3] = cmul(x[3], w3);
        if constexpr (R >= 8) {
            float2 w4 = csqr(w2);
            x[4] = cmul(x[4], w4);
            x[5] = cmul(x[5], cmul(w4, w));
            x[6] = cmul(x[6], cmul(w4, w2));
            x[7] = cmul(x[7], cmul(w4, w3));
            if constexpr (R >= 16) {
                float2 w8 = csqr(w4);
                float2 w12 = cmul(w8, w4);
                x[8] = cmul(x[8], w8);
                x[9] = cmul(x[9], cmul(w8, w));
                x[10] = cmul(x[10], cmul(w8, w2));
                x[11] = cmul(x[11], cmul(w8, w3));
                x[12] = cmul(x[12], w12);
                x[13] = cmul(x[13], cmul(w12, w));
                x[14] = cmul(x[14], cmul(w12, w2));
                x[15] = cmul(x[15], cmul(w12, w3));
            }
        }
    }
}

template <int N> struct FftGeom {
    static constexpr int E = (N >= 16) ? 16 : N;  // points per thread
    static constexpr int T = N / E;               // threads per transform
};

template <int CW> __device__ __forceinline__ int smem_word(int idx, int c) {
    constexpr int G = 32 / CW;
    return ((idx ^ ((idx >> 4) & (G - 1))) * CW) + c;
}

// One Stockham stage of radix R with sub-transform length NS already done.
template <int N, int R, int NS> __device__ __forceinline__ void stage_butterflies(float2* v, int t) {
    constexpr int E = FftGeom<N>::E, T = FftGeom<N>::T, NB = E / R, L = NS * R;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = v[b + r * NB];
        if constexpr (NS > 1) {
            const int k = (t + b * T) & (NS - 1);
            float s, c;
            sincospif((float)k * (2.0f / (float)L), &s, &c);
            apply_twiddle_powers<R>(x, make_float2(c, -s));
        }
        Dft<R>::run(x);
#pragma unroll
        for (int r = 0; r < R; ++r) v[b + r * NB] = x[r];
    }
}

// Scatter the stage outputs to shared memory and gather the next stage's inputs.
// sre/sim: this CTA's exchange planes (N*CW floats each); c = transform index in the tile.
template <int N, int CW, int R, int NS>
__device__ __forceinline__ void stage_exchange(float2* v, float* sre, float* sim, int t, int c) {
    constexpr int E = FftGeom<N>::E, T = FftGeom<N>::T, NB = E / R;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
        const int base = ((j & ~(NS - 1)) * R) + (j & (NS - 1));
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int w = smem_word<CW>(base + q * NS, c);
            sre[w] = v[b + q * NB].x;
            sim[w] = v[b + q * NB].y;
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int w = smem_word<CW>(t + T * m, c);
        v[m] = make_float2(sre[w], sim[w]);
    }
}

template <int N, int CW, int NS> struct FftStages {
    __device__ __forceinline__ static void run(float2* v, float* sre, float* sim, int t, int c) {
        constexpr int E = FftGeom<N>::E;
        constexpr int REM = N / NS;
        constexpr int R = (REM < E) ? REM : E;
        stage_butterflies<N, R, NS>(v, t);
        if constexpr (NS * R < N) {
            stage_exchange<N, CW, R, NS>(v, sre, sim, t, c);
            FftStages<N, CW, NS * R>::run(v, sre, sim, t, c);
        }
    }
};

// Forward FFT of length N over the E points held by thread t (points t + T*m).
// All T*CW threads of all transforms in the CTA must call this together when N > E
// (it contains __syncthreads).
template <int N, int CW> __device__ __forceinline__ void fft_forward(float2* v, float* sre, float* sim, int t, int c) {
    if constexpr (N > 1) FftStages<N, CW, 1>::run(v, sre, sim, t, c);
}

// Shared memory (bytes) one CTA needs for `ntransforms` interleaved transforms of length N.
template <int N> constexpr size_t fft_smem_bytes(int ntransforms) {
    return (N > FftGeom<N>::E) ? (size_t)2 * N * ntransforms * sizeof(float) : 0;
}

}  // namespace fdr
