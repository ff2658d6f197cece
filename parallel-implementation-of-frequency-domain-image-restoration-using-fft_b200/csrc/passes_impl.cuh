// passes_impl.cuh -- the two fused FFT pass kernels (rows, columns), templated on log2(N).
// Included by passes_g*.cu, each of which instantiates a group of sizes so the groups
// compile in parallel.
#pragma once
#include "fft_core.cuh"
#include "fft_mid.cuh"
#include "passes.h"

namespace fdr {

// u8 -> float * k without the conversion pipe: 0x4B000000 | b is the float 2^23 + b exactly, so
// fma(2^23 + b, k, -(2^23 * k)) = b*k with ONE rounding of b*k (2^23*k is exact for k = (float)(1/255): a
// power-of-two multiple), i.e. bit-identical to (float)b * k.
__device__ __forceinline__ float u8_scaled(uint8_t b, float k, float neg_bias) {
    return fmaf(__uint_as_float(0x4B000000u | (unsigned int)b), k, neg_bias);
}

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
    static constexpr bool MID = (LOGN >= 13);  // long rows: 32 points per thread (fft_mid.cuh)
    static constexpr int E = MID ? 32 : FftGeom<N>::E;
    static constexpr int T = N / E;
    static constexpr int RPC = (T >= 128) ? 1 : (128 / T);
    static constexpr int THREADS = T * RPC;
    static constexpr int MIN_BLOCKS = MID ? (LOGN == 13 ? 2 : 1) : ((THREADS == 128 && N >= 1024) ? 9 : 1);
    // double-buffered exchange (one barrier per exchange) while two buffers of all rows fit 32 KB
    static constexpr bool DB = false;  // measured slower on B200: the second buffer costs one resident CTA per SM
    static constexpr int EXW = MID ? mid_skew(N) : ex_words<N, 1>();  // float2 words of one row's exchange buffer
    static constexpr size_t SMEM = MID ? (size_t)EXW * sizeof(float2) : fft_smem_bytes<N, 1>() * RPC * (DB ? 2 : 1);
};
// the row transform of the geometry above
template <int LOGN> __device__ __forceinline__ void row_fft(float2* v, float2* ex, const float2* __restrict__ tw, int t) {
    if constexpr (RowGeom<LOGN>::MID)
        fft_mid_forward<(1 << LOGN)>(v, ex, tw, t);
    else
        fft_forward<(1 << LOGN), 1, RowGeom<LOGN>::DB>(v, ex, tw, t, 0);
}

// One block of RPC rows of pair (half-plane forms: plane) blockIdx.y; `row_block` is blockIdx.x in the one-shot kernel
// and the loop index in the persistent one.
template <int LOGN, int IN_MODE, int OUT_MODE, bool CONJ>
__device__ __forceinline__ void row_pass_body(const RowPassArgs& a, const int row_block) {
    using Gm = RowGeom<LOGN>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T, RPC = Gm::RPC;
    constexpr bool IN_ROWS2 = (IN_MODE == ROW_IN_ROWS2_F32 || IN_MODE == ROW_IN_ROWS2_U8);
    constexpr bool HALF = IN_ROWS2 || IN_MODE == ROW_IN_HALF || OUT_MODE == ROW_OUT_HALF || OUT_MODE == ROW_OUT_REAL_ROWS2;
    static_assert(!HALF || (E >= 16 && N >= FDR_HALF_MIN_N), "half-plane forms need at least 16 points per thread");
    constexpr int H8 = E / 2;  // points of the lower half spectrum per thread
    extern __shared__ float2 smem2[];
    const int tid = threadIdx.x;
    const int rl = (RPC > 1) ? (tid / T) : 0;
    const int t = (RPC > 1) ? (tid % T) : tid;
    const int row = row_block * RPC + rl;
    const int pair = blockIdx.y + a.pair_base;  // half-plane forms: the plane (local unit)
    const bool active = row < a.nrows;
    float2* ex = smem2 + (size_t)rl * Gm::EXW * (Gm::DB ? 2 : 1);

    const long long u0 = HALF ? (long long)pair : 2LL * pair, u1 = HALF ? u0 : u0 + 1;  // local units
    const bool has1 = HALF ? true : (a.unit_base + u1) < a.units_total;

    float2 v[E];
    if constexpr (IN_MODE == ROW_IN_COMPLEX) {
        const float2* src = a.cin + (long long)pair * a.cplane + (long long)row * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 z = make_float2(0.f, 0.f);
            if (active) z = src[T * m];
            if constexpr (CONJ) z.y = -z.y;
            v[m] = z;
        }
    } else if constexpr (IN_MODE == ROW_IN_GATHER) {
        const int cl_mask = (1 << a.peer_shift) - 1;
        const long long off = (long long)pair * a.peer_plane + ((long long)(a.row0 + row) << a.peer_shift);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int x = t + T * m;
            float2 z = make_float2(0.f, 0.f);
            if (active) {
                const float2* src = a.peers[x >> a.peer_shift];
                z = src[off + (x & cl_mask)];
            }
            v[m] = z;
        }
    } else if constexpr (IN_MODE == ROW_IN_HALF) {
        // The half plane holds conj(Y_y[k]) for k < N/2 (what the column pass leaves behind).  With A = stored row y, B = stored
        // row y + D, the packed inverse input conj(Z), Z = Y_y + i Y_{y+D} (Hermitian extension), is
        //   conj Z[k]   = A - i B            = (A.x + B.y,  A.y - B.x)
        //   conj Z[N-k] = conj(A) - i conj(B) = (A.x - B.y, -A.y - B.x)
        // the mirrored value belongs to thread T - t, register E-1-m: one shared-memory exchange.
        const int mask = (1 << a.hp_shift) - 1;
        const long long b1 = (long long)pair * a.hp_plane + ((long long)(a.row0 + row) << a.hp_shift);
        const long long b2 = b1 + ((long long)a.pair_dist << a.hp_shift);
        float2 mv[H8];
#pragma unroll
        for (int m = 0; m < H8; ++m) {
            const int k = t + T * m;
            float2 A = make_float2(0.f, 0.f), B = A;
            if (active) {
                const float2* src = a.hp_peers[k >> a.hp_shift];
                A = src[b1 + (k & mask)];
                B = src[b2 + (k & mask)];
            }
            v[m] = make_float2(A.x + B.y, A.y - B.x);
            mv[m] = make_float2(A.x - B.y, -(A.y + B.x));
        }
#pragma unroll
        for (int m = 0; m < H8; ++m) ex[(H8 - m) * T - t] = mv[m];  // (t = 0, m = 0 lands in the spare word H8*T)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < H8; ++j) v[H8 + j] = ex[j * T + t];
        if (t == 0) {  // Nyquist column: word 0 was written by nobody
            float2 A = make_float2(0.f, 0.f), B = A;
            if (active) {
                const float2* nq = a.nyq_peers[pair % a.nyq_world] + (long long)pair * a.nyq_plane + a.row0 + row;
                A = nq[0];
                B = nq[a.pair_dist];
            }
            v[H8] = make_float2(A.x + B.y, A.y - B.x);
        }
        // (the first exchange of the transform starts with a barrier: these reads are done before `ex` is reused)
    } else if constexpr (IN_MODE == ROW_IN_PAIR_F32 || IN_MODE == ROW_IN_ROWS2_F32) {
        const float* p0 = a.in_f32 + (a.unit_base + u0) * a.in_unit_stride + (long long)row * a.in_row_stride;
        const float* p1;
        float k1;
        if constexpr (IN_ROWS2) {
            const bool has2 = row + a.pair_dist < a.rows_in;
            p1 = has2 ? p0 + (long long)a.pair_dist * a.in_row_stride : p0;
            k1 = has2 ? 1.f : 0.f;
        } else {
            p1 = has1 ? a.in_f32 + (a.unit_base + u1) * a.in_unit_stride + (long long)row * a.in_row_stride : p0;
            k1 = has1 ? 1.f : 0.f;
        }
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int x = t + T * m;
            float2 z = make_float2(0.f, 0.f);
            if (active && x < a.img_cols) {
                z.x = __ldg(p0 + x);
                z.y = __ldg(p1 + x) * k1;
            }
            v[m] = z;
        }
    } else {  // ROW_IN_PAIR_U8 / ROW_IN_ROWS2_U8 : x * (float)(1/255.)  (serial.cpp:24-25 convertTo + /= 255.0)
        const float inv255 = (float)(1.0 / 255.0);
        const long long g0 = a.unit_base + u0;
        long long g1;
        int row1;       // second operand: plane g1, row row1
        float k1;
        if constexpr (IN_ROWS2) {
            const bool has2 = row + a.pair_dist < a.rows_in;
            g1 = g0;
            row1 = has2 ? row + a.pair_dist : row;
            k1 = has2 ? inv255 : 0.f;
        } else {
            g1 = has1 ? a.unit_base + u1 : g0;
            row1 = row;
            k1 = has1 ? inv255 : 0.f;
        }
        if (a.channels == 3) {
            // BGR fast path: compile-time pixel stride, so every load is base + immediate
            const long long i0 = g0 / 3, i1 = g1 / 3;
            const int c0 = (int)(g0 - i0 * 3), c1 = (int)(g1 - i1 * 3);
            const uint8_t* p0 = a.in_u8 + ((i0 * a.img_rows + row) * (long long)a.img_cols) * 3 + c0 + 3 * t;
            const uint8_t* p1 = a.in_u8 + ((i1 * a.img_rows + row1) * (long long)a.img_cols) * 3 + c1 + 3 * t;
            if (active && a.img_cols == N) {
                const float nb0 = -8388608.0f * inv255, nb1 = -8388608.0f * k1;
#pragma unroll
                for (int m = 0; m < E; ++m)
                    v[m] = make_float2(u8_scaled(__ldg(p0 + 3 * T * m), inv255, nb0), u8_scaled(__ldg(p1 + 3 * T * m), k1, nb1));
            } else {
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    float2 z = make_float2(0.f, 0.f);
                    if (active && t + T * m < a.img_cols) {
                        z.x = (float)__ldg(p0 + 3 * T * m) * inv255;
                        z.y = (float)__ldg(p1 + 3 * T * m) * k1;
                    }
                    v[m] = z;
                }
            }
        } else {
            const int C = a.channels;
            const long long i0 = g0 / C, i1 = g1 / C;
            const int c0 = (int)(g0 - i0 * C), c1 = (int)(g1 - i1 * C);
            const uint8_t* p0 = a.in_u8 + ((i0 * a.img_rows + row) * (long long)a.img_cols) * C + c0;
            const uint8_t* p1 = a.in_u8 + ((i1 * a.img_rows + row1) * (long long)a.img_cols) * C + c1;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int x = t + T * m;
                float2 z = make_float2(0.f, 0.f);
                if (active && x < a.img_cols) {
                    z.x = (float)__ldg(p0 + x * C) * inv255;
                    z.y = (float)__ldg(p1 + x * C) * k1;
                }
                v[m] = z;
            }
        }
    }

    row_fft<LOGN>(v, ex, a.tw, t);

    if constexpr (OUT_MODE == ROW_OUT_HALF) {
        // Untangle Z = FFT(row_y + i row_{y+D}): thread t needs Z[N - k] for its k = t + T*m (m < 8), held by thread T - t
        // in register E-1-m.  Every thread parks its upper half in shared memory as word (j, t) = j*T + t (j = m - 8), thread
        // 0 adds Z[0] as word 8*T, and the mirror of (t, m) is read from word (8 - m)*T - t -- also right for t = 0.
        __syncthreads();  // the last exchange of the transform may still be read
#pragma unroll
        for (int j = 0; j < H8; ++j) ex[j * T + t] = v[H8 + j];
        if (t == 0) ex[H8 * T] = v[0];
        __syncthreads();
        const int gy = a.row0 + row, gy2 = gy + a.pair_dist;
        const bool st1 = active && gy < a.hp_rows_store, st2 = active && gy2 < a.hp_rows_store;
        const int mask = (1 << a.hp_shift) - 1;
        const long long b1 = (long long)pair * a.hp_plane + ((long long)gy << a.hp_shift);
        const long long b2 = (long long)pair * a.hp_plane + ((long long)gy2 << a.hp_shift);
#pragma unroll
        for (int m = 0; m < H8; ++m) {
            const float2 z = v[m], zm = ex[(H8 - m) * T - t];
            const int k = t + T * m;
            float2* dst = a.hp_peers[k >> a.hp_shift];
            if (dst) {
                if (st1) dst[b1 + (k & mask)] = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
                if (st2) dst[b2 + (k & mask)] = make_float2(0.5f * (z.y + zm.y), 0.5f * (zm.x - z.x));
            }
        }
        if (t == 0) {  // Z[N/2] = X_y[N/2] + i X_{y+D}[N/2], both real
            float2* nq = a.nyq_peers[pair % a.nyq_world];
            if (nq) {
                nq += (long long)pair * a.nyq_plane;
                if (st1) nq[gy] = make_float2(v[H8].x, 0.f);
                if (st2) nq[gy2] = make_float2(v[H8].y, 0.f);
            }
        }
    } else if constexpr (OUT_MODE == ROW_OUT_SCATTER) {
        if (active) {
            const int cl_mask = (1 << a.peer_shift) - 1;
            const long long off = (long long)pair * a.peer_plane + ((long long)(a.row0 + row) << a.peer_shift);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int x = t + T * m;
                float2* dst = a.peers[x >> a.peer_shift];
                if (dst) dst[off + (x & cl_mask)] = v[m];
            }
        }
    } else if constexpr (OUT_MODE == ROW_OUT_COMPLEX) {
        if (active) {
            float2* dst = a.cout + (long long)pair * a.cplane + (long long)row * N + t;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                float2 z = v[m];
                if constexpr (CONJ) z.y = -z.y;
                dst[T * m] = z;
            }
        }
    } else {
        // inverse row pass of the restoration: input was conj(column result), so the restored
        // pair is conj(v): plane a = v.x, plane b = -v.y.  (ROW_OUT_REAL_ROWS2: row y = v.x, row y + D = -v.y of ONE plane.)
        constexpr bool ROWS2 = (OUT_MODE == ROW_OUT_REAL_ROWS2);
        float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
        if (active) {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                v[m].y = -v[m].y;
                mn0 = fminf(mn0, v[m].x);
                mx0 = fmaxf(mx0, v[m].x);
                mn1 = fminf(mn1, v[m].y);
                mx1 = fmaxf(mx1, v[m].y);
            }
            const int r0 = row, r1 = ROWS2 ? row + a.pair_dist : row;
            const bool w0 = r0 < a.raw_rows, w1 = has1 && r1 < a.raw_rows;
            float* d0 = a.raw + u0 * a.raw_unit_stride + (long long)r0 * a.raw_cols + t;
            float* d1 = a.raw + u1 * a.raw_unit_stride + (long long)r1 * a.raw_cols + t;
            if (a.raw_cols == N && w0 && w1) {  // un-cropped width, both rows stored: no per-element predicates
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    d0[T * m] = v[m].x;
                    d1[T * m] = v[m].y;
                }
            } else if (w0 || w1) {
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    if (t + T * m < a.raw_cols) {
                        if (w0) d0[T * m] = v[m].x;
                        if (w1) d1[T * m] = v[m].y;
                    }
                }
            }
        }
        if constexpr (ROWS2) {  // both halves belong to the same plane
            mn0 = fminf(mn0, mn1);
            mx0 = fmaxf(mx0, mx1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
            if constexpr (!ROWS2) {
                mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
            }
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
                const int slot = row_block & (FDR_MINMAX_SLOTS - 1);
                unsigned int* m0 = a.minmax + (u0 * FDR_MINMAX_SLOTS + slot) * 2;
                atomicMin(m0, f32_ordered(mn0));
                atomicMax(m0 + 1, f32_ordered(mx0));
                if (has1 && !ROWS2) {
                    unsigned int* m1 = a.minmax + (u1 * FDR_MINMAX_SLOTS + slot) * 2;
                    atomicMin(m1, f32_ordered(mn1));
                    atomicMax(m1 + 1, f32_ordered(mx1));
                }
            }
        }
    }
}

// Long rows (N >= 8192) leave room for one or two CTAs per SM, so nothing else hides the latency of a CTA's input loads.
// Thread 0 therefore asks the L2 for the rows of the CTA that will run on this SM next (`a.prefetch_dist` row blocks ahead in
// launch order: SMs x resident CTAs) with one cp.async.bulk.prefetch.L2 per contiguous run; those loads then hit the L2
// while this CTA is still in its butterflies.  Local memory only (peer slabs are not cached here).
__device__ __forceinline__ void l2_prefetch_run(const void* p, long long bytes) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(p);
    const unsigned long long lo = (a + 15) & ~15ull, hi = (a + (unsigned long long)bytes) & ~15ull;
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
}
template <int LOGN, int IN_MODE> __device__ __forceinline__ void row_prefetch_next(const RowPassArgs& a) {
    using Gm = RowGeom<LOGN>;
    constexpr int N = Gm::N;
    if (threadIdx.x != 0 || a.prefetch_dist <= 0) return;
    const int nblk = (a.nrows + Gm::RPC - 1) / Gm::RPC;
    int rb = (int)blockIdx.x + a.prefetch_dist, y = blockIdx.y;
    if (rb >= nblk) {   // wraps into the next pair / plane of the launch
        rb -= nblk;
        y += 1;
        if (rb >= nblk || y >= (int)gridDim.y) return;
    }
    const int row = rb * Gm::RPC, pair = y + a.pair_base;
    const int nr = (a.nrows - row < Gm::RPC) ? a.nrows - row : Gm::RPC;
    if constexpr (IN_MODE == ROW_IN_COMPLEX) {
        l2_prefetch_run(a.cin + (long long)pair * a.cplane + (long long)row * N, 8LL * N * nr);
    } else if constexpr (IN_MODE == ROW_IN_HALF) {
        if (a.hp_shift == LOGN - 1 || a.hp_local) {   // one owner holds whole half rows, or all owner blocks are local
            const int owners = 1 << (LOGN - 1 - a.hp_shift);
            for (int g = 0; g < owners; ++g) {
                const float2* base = a.hp_peers[g] + (long long)pair * a.hp_plane + ((long long)(a.row0 + row) << a.hp_shift);
                l2_prefetch_run(base, (8LL << a.hp_shift) * nr);
                l2_prefetch_run(base + ((long long)a.pair_dist << a.hp_shift), (8LL << a.hp_shift) * nr);
            }
        }
    } else if constexpr (IN_MODE == ROW_IN_PAIR_U8 || IN_MODE == ROW_IN_ROWS2_U8) {
        const bool rows2 = (IN_MODE == ROW_IN_ROWS2_U8);
        const long long g0 = a.unit_base + (rows2 ? (long long)pair : 2LL * pair);
        const long long img = g0 / a.channels;
        const long long rowbytes = (long long)a.img_cols * a.channels;
        const uint8_t* r0 = a.in_u8 + (img * a.img_rows + row) * rowbytes;
        l2_prefetch_run(r0, rowbytes * nr);
        if (rows2 && row + a.pair_dist < a.rows_in) l2_prefetch_run(r0 + (long long)a.pair_dist * rowbytes, rowbytes * nr);
    } else if constexpr (IN_MODE == ROW_IN_PAIR_F32 || IN_MODE == ROW_IN_ROWS2_F32) {
        const bool rows2 = (IN_MODE == ROW_IN_ROWS2_F32);
        const long long u0 = a.unit_base + (rows2 ? (long long)pair : 2LL * pair);
        const float* r0 = a.in_f32 + u0 * a.in_unit_stride + (long long)row * a.in_row_stride;
        l2_prefetch_run(r0, 4LL * a.img_cols);
        if (rows2) {
            if (row + a.pair_dist < a.rows_in) l2_prefetch_run(r0 + (long long)a.pair_dist * a.in_row_stride, 4LL * a.img_cols);
        } else if (u0 + 1 < a.units_total) {
            l2_prefetch_run(r0 + a.in_unit_stride, 4LL * a.img_cols);
        }
    }
}

template <int LOGN, int IN_MODE, int OUT_MODE, bool CONJ>
__global__ void __launch_bounds__(RowGeom<LOGN>::THREADS, RowGeom<LOGN>::MIN_BLOCKS) row_pass_kernel(const RowPassArgs a) {
    if constexpr (LOGN >= 13) row_prefetch_next<LOGN, IN_MODE>(a);
    if (a.mm_reset && blockIdx.x == 0) {   // re-arm the extrema slots of this pair's planes (pass 3 of the same call fills them)
        constexpr bool HALF = (IN_MODE == ROW_IN_ROWS2_F32 || IN_MODE == ROW_IN_ROWS2_U8);
        const int pair = blockIdx.y + a.pair_base;
        const int u0 = HALF ? pair : 2 * pair;
        int nu = HALF ? 1 : 2;
        if (u0 + nu > a.local_units) nu = a.local_units - u0;
        uint2* e = reinterpret_cast<uint2*>(a.mm_reset) + (size_t)u0 * FDR_MINMAX_SLOTS;
        for (int i = threadIdx.x; i < nu * FDR_MINMAX_SLOTS; i += blockDim.x) e[i] = make_uint2(0xFFFFFFFFu, 0u);
    }
    row_pass_body<LOGN, IN_MODE, OUT_MODE, CONJ>(a, blockIdx.x);
}

// Grid-limited persistent form (a.max_ctas CTAs per pair, each looping over row blocks): the NVLink-bound
// exchange passes of the row-sharded path then leave SMs free for the other plane pair's column phase
// (fdr_dist pair pipeline).  Only instantiated for the scatter / gather modes.
template <int LOGN, int IN_MODE, int OUT_MODE, bool CONJ>
__global__ void __launch_bounds__(RowGeom<LOGN>::THREADS, RowGeom<LOGN>::MIN_BLOCKS) row_pass_persist_kernel(const RowPassArgs a) {
    const int nblk = (a.nrows + RowGeom<LOGN>::RPC - 1) / RowGeom<LOGN>::RPC;
    for (int rb = blockIdx.x; rb < nblk; rb += gridDim.x) {
        row_pass_body<LOGN, IN_MODE, OUT_MODE, CONJ>(a, rb);
        __syncthreads();  // shared memory (exchange buffer, min/max scratch) is reused by the next block
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
    static constexpr size_t SMEM = fft_smem_bytes<N, CW>();
};

template <int LOGN, int CW, int MODE, bool CONJ>
__global__ void __launch_bounds__(ColGeom<LOGN, CW>::THREADS, (ColGeom<LOGN, CW>::THREADS <= 512 && MODE == COL_WIENER) ? 1024 / ColGeom<LOGN, CW>::THREADS : 1) col_pass_kernel(const ColPassArgs a) {
    using Gm = ColGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T;
    extern __shared__ float2 smem2[];
    float2* ex = smem2;
    const int tid = threadIdx.x;
    const int c = tid % CW, t = tid / CW;
    const int col = blockIdx.x * CW + c;
    const bool active = col < a.pitch;
    const long long stride = (long long)T * a.pitch;
    const long long first = (long long)t * a.pitch + col;
    const long long wstride = stride, wfirst = first;
    float2* base = a.data + (long long)(blockIdx.y + a.pair_base) * a.cplane + first;

    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        float2 z = make_float2(0.f, 0.f);
        if (active && t + T * m < a.rows_valid) z = base[m * stride];
        if constexpr (CONJ) z.y = -z.y;
        v[m] = z;
    }

    if constexpr (MODE != COL_COPY) fft_forward<N, CW>(v, ex, a.tw, t, c);

    if constexpr (MODE == COL_WIENER) {
        const float2* wf = a.wiener + wfirst;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 w = make_float2(0.f, 0.f);
            if (active) w = __ldg(wf + m * wstride);
            const float2 y = cmul(v[m], w);
            v[m] = make_float2(y.x, -y.y);
        }
        fft_forward<N, CW>(v, ex, a.tw, t, c);
    }
    if constexpr (MODE == COL_FILTER) {
        const float2* wf = a.wiener + wfirst;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 w = make_float2(0.f, 0.f);
            if (active) w = __ldg(wf + m * wstride);
            v[m] = cmul(v[m], w);
        }
    }

    if (!active) return;
    if constexpr (MODE == COL_MAKE_WIENER) {
        float2* wo = a.wiener_out + wfirst;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float hr = v[m].x, hi = v[m].y;
            const float denom = fmaf(hr, hr, hi * hi) + a.K;  // fft_serial.cpp:195-197
            wo[m * wstride] = make_float2(hr / denom, -hi / denom);
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) {
            float2 z = v[m];
            if constexpr (CONJ) z.y = -z.y;
            base[m * stride] = z;
        }
    }
}

template <int LOGN, int IN_MODE, int OUT_MODE, bool CONJ> cudaError_t launch_row_variant(const RowPassArgs& a0, cudaStream_t s) {
    using Gm = RowGeom<LOGN>;
    RowPassArgs a = a0;
    if constexpr (LOGN >= 13) {
        static const int pf = getenv("FDR_ROW_PREFETCH") ? atoi(getenv("FDR_ROW_PREFETCH")) : -1;   // row blocks ahead; 0 = off
        a.prefetch_dist = pf >= 0 ? pf : device_sm_count() * Gm::MIN_BLOCKS;
    }
    if constexpr (Gm::MID) {   // the long-row core has its own twiddle table (the caller's is the 16-point core's)
        cudaError_t e = get_twiddles_mid(1 << LOGN, &a.tw);
        if (e != cudaSuccess) return e;
    }
    if (Gm::SMEM > 48 * 1024) {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(row_pass_kernel<LOGN, IN_MODE, OUT_MODE, CONJ>), Gm::SMEM);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((a.nrows + Gm::RPC - 1) / Gm::RPC, a.npairs);
    if constexpr (IN_MODE == ROW_IN_GATHER || OUT_MODE == ROW_OUT_SCATTER || ((IN_MODE == ROW_IN_HALF || OUT_MODE == ROW_OUT_HALF) && LOGN >= 10)) {
        if (a.max_ctas > 0 && (unsigned)a.max_ctas < grid.x) {
            if (Gm::SMEM > 48 * 1024) {
                cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(row_pass_persist_kernel<LOGN, IN_MODE, OUT_MODE, CONJ>), Gm::SMEM);
                if (e != cudaSuccess) return e;
            }
            grid.x = a.max_ctas;
            a.prefetch_dist = 0;
            row_pass_persist_kernel<LOGN, IN_MODE, OUT_MODE, CONJ><<<grid, Gm::THREADS, Gm::SMEM, s>>>(a);
            return cudaGetLastError();
        }
    }
    row_pass_kernel<LOGN, IN_MODE, OUT_MODE, CONJ><<<grid, Gm::THREADS, Gm::SMEM, s>>>(a);
    return cudaGetLastError();
}

template <int LOGN> cudaError_t launch_row_pass_t(const RowPassArgs& a, cudaStream_t s) {
    if (a.in_mode == ROW_IN_PAIR_U8 && a.out_mode == ROW_OUT_COMPLEX) return launch_row_variant<LOGN, ROW_IN_PAIR_U8, ROW_OUT_COMPLEX, false>(a, s);
    if (a.in_mode == ROW_IN_PAIR_F32 && a.out_mode == ROW_OUT_COMPLEX) return launch_row_variant<LOGN, ROW_IN_PAIR_F32, ROW_OUT_COMPLEX, false>(a, s);
    if (a.in_mode == ROW_IN_COMPLEX && a.out_mode == ROW_OUT_REAL_PAIR) return launch_row_variant<LOGN, ROW_IN_COMPLEX, ROW_OUT_REAL_PAIR, false>(a, s);
    if (a.in_mode == ROW_IN_PAIR_U8 && a.out_mode == ROW_OUT_SCATTER) return launch_row_variant<LOGN, ROW_IN_PAIR_U8, ROW_OUT_SCATTER, false>(a, s);
    if (a.in_mode == ROW_IN_PAIR_F32 && a.out_mode == ROW_OUT_SCATTER) return launch_row_variant<LOGN, ROW_IN_PAIR_F32, ROW_OUT_SCATTER, false>(a, s);
    if (a.in_mode == ROW_IN_GATHER && a.out_mode == ROW_OUT_REAL_PAIR) return launch_row_variant<LOGN, ROW_IN_GATHER, ROW_OUT_REAL_PAIR, false>(a, s);
    if constexpr ((1 << LOGN) >= FDR_HALF_MIN_N) {  // half-plane forms (one real plane per transform)
        if (a.in_mode == ROW_IN_ROWS2_U8 && a.out_mode == ROW_OUT_HALF) return launch_row_variant<LOGN, ROW_IN_ROWS2_U8, ROW_OUT_HALF, false>(a, s);
        if (a.in_mode == ROW_IN_ROWS2_F32 && a.out_mode == ROW_OUT_HALF) return launch_row_variant<LOGN, ROW_IN_ROWS2_F32, ROW_OUT_HALF, false>(a, s);
        if (a.in_mode == ROW_IN_HALF && a.out_mode == ROW_OUT_REAL_ROWS2) return launch_row_variant<LOGN, ROW_IN_HALF, ROW_OUT_REAL_ROWS2, false>(a, s);
    }
    if (a.in_mode == ROW_IN_COMPLEX && a.out_mode == ROW_OUT_COMPLEX && !a.conj)
        return launch_row_variant<LOGN, ROW_IN_COMPLEX, ROW_OUT_COMPLEX, false>(a, s);
    if (a.in_mode == ROW_IN_COMPLEX && a.out_mode == ROW_OUT_COMPLEX && a.conj)
        return launch_row_variant<LOGN, ROW_IN_COMPLEX, ROW_OUT_COMPLEX, true>(a, s);
    return cudaErrorInvalidValue;
}

template <int LOGN, int CW, int MODE, bool CONJ> cudaError_t launch_col_variant(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColGeom<LOGN, CW>;
    if (Gm::SMEM > 48 * 1024) {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(col_pass_kernel<LOGN, CW, MODE, CONJ>), Gm::SMEM);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((a.pitch + CW - 1) / CW, a.npairs);
    col_pass_kernel<LOGN, CW, MODE, CONJ><<<grid, Gm::THREADS, Gm::SMEM, s>>>(a);
    return cudaGetLastError();
}

template <int LOGN, int CW> cudaError_t launch_col_pass_t(const ColPassArgs& a, cudaStream_t s) {
    switch (a.mode) {
        case COL_FFT:
            return a.conj ? launch_col_variant<LOGN, CW, COL_FFT, true>(a, s) : launch_col_variant<LOGN, CW, COL_FFT, false>(a, s);
        case COL_WIENER: return launch_col_variant<LOGN, CW, COL_WIENER, false>(a, s);
        case COL_MAKE_WIENER: return launch_col_variant<LOGN, CW, COL_MAKE_WIENER, false>(a, s);
        case COL_FILTER: return launch_col_variant<LOGN, CW, COL_FILTER, false>(a, s);
        case COL_COPY: return launch_col_variant<LOGN, CW, COL_COPY, false>(a, s);
    }
    return cudaErrorInvalidValue;
}

template <int LOGN> cudaError_t tw_fill_t(float2* tw, cudaStream_t s) {
    constexpr int N = 1 << LOGN;
    constexpr int total = TwTotal<N>::value;
    if (total == 0) return cudaSuccess;
    constexpr int per_stage_max = 15 * FftGeom<N>::T;
    tw_fill_kernel<N><<<(per_stage_max + 255) / 256, 256, 0, s>>>(tw);
    return cudaGetLastError();
}
template <int LOGN> constexpr int tw_total_t() { return TwTotal<(1 << LOGN)>::value; }

// Default tile width per length: keep T*CW <= 512 threads, CW*8 B >= 32 B where possible.
constexpr int default_col_cw(int logn) {
    return logn <= 6 ? 32 : logn <= 8 ? 16 : logn <= 10 ? 8 : logn == 11 ? 4 : logn == 12 ? 2 : 1;
}

}  // namespace fdr
