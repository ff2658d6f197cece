// col_wide.cu -- column pass (COL_WIENER) on the 64-points-per-thread core (fft_wide.cuh).
//
// Same data flow as col_tma.cu -- TMA tile in, FFT along y, TMA Wiener tile, multiply, conj, FFT
// again (= inverse), TMA tile out, in place -- but a column of N points is held by N/64 threads,
// so each transform has ONE shared-memory exchange instead of two and a CTA is 4 warps
// (N = 2048, CW = 4) instead of 16: the pass is bound by the shared-memory / L1 data pipe
// (ncu: 65-70 % busy in col_tma.cu), and this form issues 38 % fewer wavefronts per tile.
// Three CTAs per SM (register-limited, ~168 registers per thread), each with its own 68 KB buffer.
#include <cstdlib>
#include <map>
#include <mutex>

#include "fft_wide.cuh"
#include "passes.h"
#include "tma_util.cuh"

namespace fdr {

template <int LOGN, int CW> struct ColWideGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = WideGeom<N>::T;
    static constexpr int THREADS = T * CW;
    static constexpr size_t SMEM = (size_t)wide_ex_words<N, CW>() * sizeof(float2);
    static constexpr int MIN_BLOCKS = (65536 / (168 * THREADS) < 1) ? 1 : 65536 / (168 * THREADS);
};

// tile order: the column tiles sharing a 128-byte line are neighbours, then the pair, then the next line
template <int CW> __device__ __forceinline__ void wide_tile_coords(int id, int tiles_x, int npairs, int& xt, int& pr) {
    constexpr int XG = (CW * 8 >= 128) ? 1 : 128 / (CW * 8);
    if (tiles_x % XG == 0) {
        const int xlo = id % XG, rest = id / XG;
        pr = rest % npairs;
        xt = (rest / npairs) * XG + xlo;
    } else {
        xt = id % tiles_x;
        pr = id / tiles_x;
    }
}

// PROBE = 0: the pass.  Timing probes (fdr_plan_time_pass): 1 = tile in, tile out, nothing else (the TMA
// floor of two of the three transfers); 2 = all three transfers, registers untouched by any FFT.
template <int LOGN, int CW, int PROBE = 0>
__global__ void __launch_bounds__(ColWideGeom<LOGN, CW>::THREADS, ColWideGeom<LOGN, CW>::MIN_BLOCKS)
    col_wiener_wide_kernel(const __grid_constant__ CUtensorMap tm_data, const __grid_constant__ CUtensorMap tm_w, const ColPassArgs a,
                           const float2* __restrict__ tw) {
    using Gm = ColWideGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = 64, T = Gm::T;
    constexpr int BOX_ROWS = 256;
    constexpr int NBOX = N / BOX_ROWS;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    extern __shared__ __align__(128) float2 smem2[];
    __shared__ __align__(8) unsigned long long bar;
    float2* ex = smem2;
    const int tid = threadIdx.x;
    const int c = tid % CW, t = tid / CW;
    int xt, pr;
    wide_tile_coords<CW>(blockIdx.x, a.pitch / CW, a.npairs, xt, pr);
    const int x0 = xt * CW * 2;               // tensor maps count 32-bit floats along x
    const int y0 = (pr + a.pair_base) * N;    // pair p occupies tensor rows [p*N, (p+1)*N)
    const int nbox_valid = (a.rows_valid + BOX_ROWS - 1) / BOX_ROWS;
    const int kb = (a.wiener_blocks > 1) ? (pr + a.pair_base) % a.wiener_blocks : 0;  // row block of the Wiener factor

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, nbox_valid * BOX_BYTES);
        for (int b = 0; b < nbox_valid; ++b) tma_load_2d(ex + (size_t)b * BOX_ROWS * CW, &tm_data, x0, y0 + b * BOX_ROWS, &bar);
    }
    mbar_wait(&bar, 0);

    // Measured and rejected (B200, 12 pairs of 2048^2): one copy of the FFT code looped over the two
    // phases (20.1 vs 19.2 us per pair although the code shrinks from 4700 to 3200 instructions), and
    // cp.async.bulk.prefetch.tensor of the Wiener tile and of the next CTA's data tile into L2 (23.9 us).
    float2 v[E];
    if (a.rows_valid >= N) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = ex[(size_t)(t + T * m) * CW + c];
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int r = t + T * m;
            v[m] = (r < a.rows_valid) ? ex[(size_t)r * CW + c] : make_float2(0.f, 0.f);
        }
    }
    // The Wiener tile is requested as soon as the exchange of the first transform has been read back: the
    // transfer then overlaps the second-stage butterflies instead of being waited for after them.
    auto load_wiener = [&]() {
        if (tid == 0) {
            mbar_expect_tx(&bar, NBOX * BOX_BYTES);
            if (a.wiener_tiled) {  // one contiguous block per tile
                const float2* src = a.wiener_tiled + ((size_t)kb * (a.pitch / CW) + xt) * N * CW;
                for (int b = 0; b < NBOX; ++b) bulk_load_1d(ex + (size_t)b * BOX_ROWS * CW, src + (size_t)b * BOX_ROWS * CW, BOX_BYTES, &bar);
            } else {
                for (int b = 0; b < NBOX; ++b) tma_load_2d(ex + (size_t)b * BOX_ROWS * CW, &tm_w, x0, kb * N + b * BOX_ROWS, &bar);
            }
        }
    };
    if constexpr (PROBE == 0) {
        fft_wide_forward<N, CW>(v, ex, tw, t, c, CtaBarrier(), load_wiener);  // (its exchange starts with a barrier: the tile is consumed)
    } else {
        __syncthreads();
        if constexpr (PROBE != 1) load_wiener();
    }
    if constexpr (PROBE != 1) {
    mbar_wait(&bar, 1);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const float2 y = cmul(v[m], ex[(size_t)(t + T * m) * CW + c]);
        v[m] = make_float2(y.x, -y.y);
    }
    }
    if constexpr (PROBE == 0) fft_wide_forward<N, CW>(v, ex, tw, t, c);

    __syncthreads();  // exchange reads done
#pragma unroll
    for (int m = 0; m < E; ++m) ex[(size_t)(t + T * m) * CW + c] = v[m];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < NBOX; ++b) tma_store_2d(&tm_data, x0, y0 + b * BOX_ROWS, ex + (size_t)b * BOX_ROWS * CW);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Pipelined form (same scheme as col_wiener_pipe_kernel in col_tma.cu): one persistent CTA per SM
// with NG = 2 groups of T*CW threads and 3 tile buffers.  The transfer-only probes show the per-tile
// kernel above is bound by its own serialisation -- load, wait, Wiener load, wait, store, exit, with
// only three tiles per SM and nothing in flight to HBM for two thirds of a CTA's life -- not by the
// FFT (the probe without any FFT takes as long as the full kernel).  Here the third buffer always has
// the next tile's load in flight while both groups compute.  MEASURED SLOWER (20.6 vs 17.2 us per pair):
// the pass is bound by the throughput of 32-byte-row TMA traffic (5.4 TB/s from HBM, 6.4 TB/s from L2,
// profiles/r1b/tma_copy_probe.txt), and 8 warps per SM hide less than 12.  Kept as an opt-in (FDR_WIDE_PIPE=1).
// ---------------------------------------------------------------------------------------------
template <int LOGN, int CW> struct ColWidePipeGeom {
    using Gm = ColWideGeom<LOGN, CW>;
    static constexpr int GT = Gm::THREADS;
    static constexpr int NG = 2;
    static constexpr int NBUF = 3;
    static constexpr int THREADS = NG * GT;
    static constexpr size_t SMEM = Gm::SMEM * NBUF;
};

template <int LOGN, int CW>
__device__ __noinline__ void wide_issue_tile(int k, float2* smem2, unsigned long long* full, const CUtensorMap* tm_data, int rows_valid, int npairs,
                                             int pair_base, int tiles_x) {
    constexpr int N = 1 << LOGN, BOX_ROWS = 256, NBUF = ColWidePipeGeom<LOGN, CW>::NBUF;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    int xt, pr;
    wide_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, npairs, xt, pr);
    const int b = k % NBUF;
    const int nbox_valid = (rows_valid + BOX_ROWS - 1) / BOX_ROWS;
    float2* dst = smem2 + (size_t)b * wide_ex_words<N, CW>();
    mbar_expect_tx(&full[b], nbox_valid * BOX_BYTES);
    for (int q = 0; q < nbox_valid; ++q)
        tma_load_2d(dst + (size_t)q * BOX_ROWS * CW, tm_data, xt * CW * 2, (pr + pair_base) * N + q * BOX_ROWS, &full[b]);
}

template <int LOGN, int CW>
__global__ void __launch_bounds__(ColWidePipeGeom<LOGN, CW>::THREADS, 1)
    col_wiener_wide_pipe_kernel(const __grid_constant__ CUtensorMap tm_data, const __grid_constant__ CUtensorMap tm_w, const ColPassArgs a,
                                const float2* __restrict__ tw, const int tiles_x, const int ntiles) {
    using Gm = ColWideGeom<LOGN, CW>;
    using Pg = ColWidePipeGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = 64, T = Gm::T, GT = Pg::GT, NG = Pg::NG, NBUF = Pg::NBUF;
    constexpr int BOX_ROWS = 256;
    constexpr int NBOX = N / BOX_ROWS;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    constexpr size_t TILE = wide_ex_words<N, CW>();
    extern __shared__ __align__(128) float2 smem2[];
    __shared__ __align__(8) unsigned long long full[NBUF];
    __shared__ __align__(8) unsigned long long wbar[NG];
    const int tid = threadIdx.x;
    const int g = tid / GT, gt = tid - g * GT;
    const int c = gt % CW, t = gt / CW;
    const GroupBarrier gbar{1 + g, GT};
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        for (int q = 0; q < NG; ++q) mbar_init(&wbar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const int n0 = my_tiles < NBUF ? my_tiles : NBUF;
        for (int k = 0; k < n0; ++k) wide_issue_tile<LOGN, CW>(k, smem2, full, &tm_data, a.rows_valid, a.npairs, a.pair_base, tiles_x);
    }

    int wphase = 0;
    for (int k = g; k < my_tiles; k += NG) {
        const int b = k % NBUF;
        float2* ex = smem2 + (size_t)b * TILE;
        mbar_wait(&full[b], (unsigned)((k / NBUF) & 1));

        float2 v[E];
        if (a.rows_valid >= N) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = ex[(size_t)(t + T * m) * CW + c];
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int r = t + T * m;
                v[m] = (r < a.rows_valid) ? ex[(size_t)r * CW + c] : make_float2(0.f, 0.f);
            }
        }
        fft_wide_forward<N, CW, GroupBarrier>(v, ex, tw, t, c, gbar);

        gbar.sync();  // exchange buffer idle: bring in the Wiener tile
        if (gt == 0) {
            int xt, pr;
            wide_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, a.npairs, xt, pr);
            mbar_expect_tx(&wbar[g], NBOX * BOX_BYTES);
            if (a.wiener_tiled) {
                const float2* src = a.wiener_tiled + (size_t)xt * N * CW;
                for (int q = 0; q < NBOX; ++q) bulk_load_1d(ex + (size_t)q * BOX_ROWS * CW, src + (size_t)q * BOX_ROWS * CW, BOX_BYTES, &wbar[g]);
            } else {
                for (int q = 0; q < NBOX; ++q) tma_load_2d(ex + (size_t)q * BOX_ROWS * CW, &tm_w, xt * CW * 2, q * BOX_ROWS, &wbar[g]);
            }
        }
        mbar_wait(&wbar[g], (unsigned)wphase);
        wphase ^= 1;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float2 y = cmul(v[m], ex[(size_t)(t + T * m) * CW + c]);
            v[m] = make_float2(y.x, -y.y);
        }
        fft_wide_forward<N, CW, GroupBarrier>(v, ex, tw, t, c, gbar);

        gbar.sync();  // exchange reads done
#pragma unroll
        for (int m = 0; m < E; ++m) ex[(size_t)(t + T * m) * CW + c] = v[m];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        gbar.sync();
        if (gt == 0) {
            int xt, pr;
            wide_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, a.npairs, xt, pr);
            for (int q = 0; q < NBOX; ++q) tma_store_2d(&tm_data, xt * CW * 2, (pr + a.pair_base) * N + q * BOX_ROWS, ex + (size_t)q * BOX_ROWS * CW);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (k + NBUF < my_tiles) wide_issue_tile<LOGN, CW>(k + NBUF, smem2, full, &tm_data, a.rows_valid, a.npairs, a.pair_base, tiles_x);
        }
    }
}

// twiddle table of the wide core for length n on the current device (built once, double precision)
template <int N> static cudaError_t wide_twiddles(const float2** out) {
    static std::mutex mu;
    static std::map<int, float2*> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    float2* p = nullptr;
    e = cudaMalloc(&p, sizeof(float2) * WideGeom<N>::TW_ENTRIES);
    if (e != cudaSuccess) return e;
    wide_tw_fill_kernel<N><<<(WideGeom<N>::TW_ENTRIES + 255) / 256, 256>>>(p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaFree(p);
        return e;
    }
    cache[dev] = p;
    *out = p;
    return cudaSuccess;
}

bool col_wide_applicable(const ColPassArgs& a) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* env = getenv("FDR_COL_WIDE");
        enabled = (env && atoi(env) == 0) ? 0 : 1;
    }
    const bool geom = wide_tile_cols(a.n) > 0 && a.pitch % (wide_tile_cols(a.n) < 4 ? 4 : wide_tile_cols(a.n)) == 0;
    if (a.col_variant == 4) return geom;
    if (a.col_variant == 8) return a.n == 4096 && geom;
    if (a.col_variant >= 5 && a.col_variant <= 7) return a.n == 2048 && geom;
    if (!enabled || a.col_variant != 0) return false;
    // 1024: 4.96 vs 5.15 us per pair in large batches but a worse wave fit for single images (cat: 0.096 vs 0.093 ms): opt-in only
    return geom && a.n != 1024;
}

template <int LOGN, int CW, int PROBE = 0> static cudaError_t launch_wide_t(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColWideGeom<LOGN, CW>;
    CUtensorMap tm_data, tm_w;
    if (!tma_make_map(&tm_data, a.data, (long long)(a.pair_base + a.npairs) * a.n, a.pitch, CW, 256)) return cudaErrorInvalidValue;
    if (!tma_make_map(&tm_w, a.wiener, (long long)a.n * (a.wiener_blocks > 1 ? a.wiener_blocks : 1), a.pitch, CW, 256)) return cudaErrorInvalidValue;
    const float2* tw = nullptr;
    cudaError_t e = wide_twiddles<Gm::N>(&tw);
    if (e != cudaSuccess) return e;
    e = ensure_dyn_smem(reinterpret_cast<const void*>(col_wiener_wide_kernel<LOGN, CW, PROBE>), Gm::SMEM);
    if (e != cudaSuccess) return e;
    const int grid = (a.pitch / CW) * a.npairs;
    col_wiener_wide_kernel<LOGN, CW, PROBE><<<grid, Gm::THREADS, Gm::SMEM, s>>>(tm_data, tm_w, a, tw);
    return cudaGetLastError();
}

// Timing probe: TMA copy of 64 KB tiles, in and out, as BW columns x (8192 / BW) rows: what the box width costs.
template <int BW> __global__ void __launch_bounds__(128, 3) tma_copy_probe_kernel(const __grid_constant__ CUtensorMap tm, int tiles_x) {
    constexpr int ROWS = 8192 / BW, BOX_ROWS = 256, NBOX = ROWS / BOX_ROWS;
    extern __shared__ __align__(128) float2 smem2[];
    __shared__ __align__(8) unsigned long long bar;
    const int xt = blockIdx.x % tiles_x, yt = blockIdx.x / tiles_x;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 65536);
        for (int b = 0; b < NBOX; ++b) tma_load_2d(smem2 + (size_t)b * BOX_ROWS * BW, &tm, xt * BW * 2, yt * ROWS + b * BOX_ROWS, &bar);
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBOX; ++b) tma_store_2d(&tm, xt * BW * 2, yt * ROWS + b * BOX_ROWS, smem2 + (size_t)b * BOX_ROWS * BW);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
template <int BW> static cudaError_t launch_copy_probe(const ColPassArgs& a, cudaStream_t s) {
    CUtensorMap tm;
    const long long rows = (long long)(a.pair_base + a.npairs) * a.n;
    if (!tma_make_map(&tm, a.data, rows, a.pitch, BW, 256)) return cudaErrorInvalidValue;
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(tma_copy_probe_kernel<BW>), 69632);
        if (e != cudaSuccess) return e;
    }
    const int tiles_x = a.pitch / BW;
    const long long grid = (long long)tiles_x * (rows / (8192 / BW));
    tma_copy_probe_kernel<BW><<<(unsigned)grid, 128, 69632, s>>>(tm, tiles_x);
    return cudaGetLastError();
}
cudaError_t launch_tma_copy_probe(const ColPassArgs& a, int box_cols, cudaStream_t s) {
    switch (box_cols) {
        case 2: return launch_copy_probe<2>(a, s);
        case 4: return launch_copy_probe<4>(a, s);
        case 8: return launch_copy_probe<8>(a, s);
        case 16: return launch_copy_probe<16>(a, s);
        case 32: return launch_copy_probe<32>(a, s);
    }
    return cudaErrorInvalidValue;
}

__global__ void wiener_retile_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int n, int pitch, int cw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into dst
    if (i >= (long long)n * pitch) return;
    const int c = (int)(i % cw);
    const long long q = i / cw;
    const int row = (int)(q % n), xt = (int)(q / n);
    dst[i] = src[(long long)row * pitch + xt * cw + c];
}
cudaError_t launch_wiener_retile(const float2* src, float2* dst, int n, int pitch, cudaStream_t s) {
    const int cw = wide_tile_cols(n);
    if (cw == 0 || pitch % cw != 0) return cudaErrorInvalidValue;
    const long long total = (long long)n * pitch;
    wiener_retile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dst, n, pitch, cw);
    return cudaGetLastError();
}

template <int LOGN, int CW> static cudaError_t launch_wide_pipe_t(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColWideGeom<LOGN, CW>;
    using Pg = ColWidePipeGeom<LOGN, CW>;
    CUtensorMap tm_data, tm_w;
    if (!tma_make_map(&tm_data, a.data, (long long)(a.pair_base + a.npairs) * a.n, a.pitch, CW, 256)) return cudaErrorInvalidValue;
    if (!tma_make_map(&tm_w, a.wiener, a.n, a.pitch, CW, 256)) return cudaErrorInvalidValue;
    const float2* tw = nullptr;
    cudaError_t e = wide_twiddles<Gm::N>(&tw);
    if (e != cudaSuccess) return e;
    e = ensure_dyn_smem(reinterpret_cast<const void*>(col_wiener_wide_pipe_kernel<LOGN, CW>), Pg::SMEM);
    if (e != cudaSuccess) return e;
    const int nsm = device_sm_count();
    const int tiles_x = a.pitch / CW, ntiles = tiles_x * a.npairs;
    const int grid = ntiles < nsm ? ntiles : nsm;
    col_wiener_wide_pipe_kernel<LOGN, CW><<<grid, Pg::THREADS, Pg::SMEM, s>>>(tm_data, tm_w, a, tw, tiles_x, ntiles);
    return cudaGetLastError();
}

cudaError_t launch_col_wiener_wide(const ColPassArgs& a, cudaStream_t s) {
    static int pipe_enabled = -1;
    // measured slower than the per-tile form (8 warps per SM instead of 12: 20.6 vs 17.2 us per pair of 2048^2): opt-in only
    if (pipe_enabled < 0) pipe_enabled = (getenv("FDR_WIDE_PIPE") && atoi(getenv("FDR_WIDE_PIPE")) == 1) ? 1 : 0;
    const long long ntiles = (long long)(a.pitch / 4) * a.npairs;
    if (a.n == 2048 && a.wiener_blocks <= 1 && (a.col_variant == 7 || (a.col_variant == 0 && pipe_enabled && ntiles >= 4 * 148)))
        return launch_wide_pipe_t<11, 4>(a, s);
    switch (a.n) {
        case 1024: return launch_wide_t<10, 8>(a, s);
        case 4096: return a.col_variant == 8 ? launch_wide_t<12, 4>(a, s) : launch_wide_t<12, 2>(a, s);
        case 2048: return a.col_variant == 5 ? launch_wide_t<11, 4, 1>(a, s) : a.col_variant == 6 ? launch_wide_t<11, 4, 2>(a, s) : launch_wide_t<11, 4>(a, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdr
