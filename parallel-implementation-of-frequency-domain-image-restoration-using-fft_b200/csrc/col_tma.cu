// col_tma.cu -- column pass (COL_WIENER) with TMA-staged tiles.
//
// Measured (profiles/time_passes.py): the plain column kernel is bound by the
// LSU/L1 pipeline, not by HBM -- a warp that touches 32/CW rows x (CW*8) bytes costs one L1 tag
// cycle per row segment, and those cycles are shared with the shared-memory traffic of the FFT
// exchanges.  Here the strided tile (CW columns x N rows of a ROW-MAJOR plane) is moved by the
// Tensor Memory Accelerator instead: cp.async.bulk.tensor 2-D boxes of CW x 256 elements land in
// shared memory as a dense [row][CW] tile and leave the same way, so no LSU address divergence is
// paid for global memory at all, and the row passes keep their fully coalesced row-major layout.
//   load tile (TMA) -> registers -> FFT -> Wiener tile (TMA into the now idle exchange buffer) ->
//   multiply, conj -> FFT -> registers -> shared -> store tile (TMA)
// One 64 KB buffer per CTA serves as TMA landing zone, exchange buffer and TMA source; 2 CTAs/SM.
#include <cstdlib>

#include "passes_impl.cuh"
#include "tma_util.cuh"

namespace fdr {

template <int LOGN, int CW>
__global__ void __launch_bounds__(ColGeom<LOGN, CW>::THREADS, (ColGeom<LOGN, CW>::THREADS <= 512) ? 1024 / ColGeom<LOGN, CW>::THREADS : 1)
    col_wiener_tma_kernel(const __grid_constant__ CUtensorMap tm_data, const __grid_constant__ CUtensorMap tm_w, const ColPassArgs a) {
    using Gm = ColGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T;
    constexpr int BOX_ROWS = (N < 256) ? N : 256;
    constexpr int NBOX = N / BOX_ROWS;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    extern __shared__ __align__(128) float2 smem2[];
    __shared__ __align__(8) unsigned long long bar;
    float2* ex = smem2;
    const int tid = threadIdx.x;
    const int c = tid % CW, t = tid / CW;
    const int x0 = blockIdx.x * CW * 2;   // tensor maps count 32-bit floats along x
    const int y0 = (blockIdx.y + a.pair_base) * N;  // pair p occupies tensor rows [p*N, (p+1)*N)
    const int nbox_valid = (a.rows_valid + BOX_ROWS - 1) / BOX_ROWS;

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

    float2 v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int r = t + T * m;
        v[m] = (r < a.rows_valid) ? ex[(size_t)r * CW + c] : make_float2(0.f, 0.f);
    }
    fft_forward<N, CW>(v, ex, a.tw, t, c);  // (its first exchange starts with a barrier: the tile is consumed)

    __syncthreads();  // exchange buffer idle: bring in the Wiener tile
    if (tid == 0) {
        mbar_expect_tx(&bar, NBOX * BOX_BYTES);
        for (int b = 0; b < NBOX; ++b) tma_load_2d(ex + (size_t)b * BOX_ROWS * CW, &tm_w, x0, b * BOX_ROWS, &bar);
    }
    mbar_wait(&bar, 1);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const float2 y = cmul(v[m], ex[(size_t)(t + T * m) * CW + c]);
        v[m] = make_float2(y.x, -y.y);
    }
    fft_forward<N, CW>(v, ex, a.tw, t, c);

    __syncthreads();  // last exchange reads done
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
// Pipelined form: ONE persistent CTA per SM holds NG independent groups of T*CW threads (the same
// thread count, registers and 64 KB tile as two resident CTAs of the kernel above) and NG + 1 tile
// buffers.  Tile i of the CTA's strided list is processed by group i % NG in buffer i % NBUF; when a
// group has stored tile i, its leader refills that buffer with tile i + NBUF, which the OTHER
// group(s) will consume -- so a group never waits for HBM: its next tile landed while it was
// computing.  ncu (profiles/r1): 34 % of the per-tile kernel's warp samples sit in the mbarrier
// wait of the tile load; this removes that wait from the critical path.  Groups synchronise on
// named barriers (GroupBarrier), never on the whole CTA.
// ---------------------------------------------------------------------------------------------
// Tile order of the pipelined kernel: the XG column tiles that share one 128-byte line are
// neighbours, then the pair index, then the next line -- so CTAs running together touch whole DRAM
// lines, and all pairs visit a column strip back to back: its Wiener tile is read from HBM once and
// then hit in L2.  Only the group leaders evaluate this (kept out of the other threads' registers).
template <int CW> __device__ __forceinline__ void pipe_tile_coords(int id, int tiles_x, int npairs, int& xt, int& pr) {
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
// one thread: bring tile k of this CTA into buffer k % NBUF
template <int LOGN, int CW, int NBUF>
__device__ __noinline__ void pipe_issue_tile(int k, float2* smem2, unsigned long long* full, const CUtensorMap* tm_data, int rows_valid, int npairs, int pair_base, int tiles_x) {
    constexpr int N = 1 << LOGN;
    constexpr int BOX_ROWS = (N < 256) ? N : 256;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    int xt, pr;
    pipe_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, npairs, xt, pr);
    const int b = k % NBUF;
    const int nbox_valid = (rows_valid + BOX_ROWS - 1) / BOX_ROWS;
    float2* dst = smem2 + (size_t)b * ex_words<N, CW>();
    mbar_expect_tx(&full[b], nbox_valid * BOX_BYTES);
    for (int q = 0; q < nbox_valid; ++q)
        tma_load_2d(dst + (size_t)q * BOX_ROWS * CW, tm_data, xt * CW * 2, (pr + pair_base) * N + q * BOX_ROWS, &full[b]);
}

template <int LOGN, int CW> struct ColPipeGeom {
    using Gm = ColGeom<LOGN, CW>;
    static constexpr int GT = Gm::THREADS;                 // threads per group
    static constexpr int NG = (GT >= 1024) ? 1 : 1024 / GT;
    static constexpr int THREADS = NG * GT;
    static constexpr size_t TILE_BYTES = Gm::SMEM;
    static constexpr int NBUF_MAX = (int)((220 * 1024) / TILE_BYTES);
    static constexpr int NBUF = (2 * NG < NBUF_MAX) ? 2 * NG : NBUF_MAX;  // 3 x 64 KB (2048, 4096); 6 x 32 KB (1024)
    static_assert(NBUF > NG, "every group needs its own buffer plus at least one landing buffer");
    static constexpr size_t SMEM = TILE_BYTES * NBUF;
};

template <int LOGN, int CW>
__global__ void __launch_bounds__(ColPipeGeom<LOGN, CW>::THREADS, 1)
    col_wiener_pipe_kernel(const __grid_constant__ CUtensorMap tm_data, const __grid_constant__ CUtensorMap tm_w, const ColPassArgs a,
                           const int tiles_x, const int ntiles) {
    using Gm = ColGeom<LOGN, CW>;
    using Pg = ColPipeGeom<LOGN, CW>;
    constexpr int N = Gm::N, E = Gm::E, T = Gm::T, GT = Pg::GT, NG = Pg::NG, NBUF = Pg::NBUF;
    constexpr int BOX_ROWS = (N < 256) ? N : 256;
    constexpr int NBOX = N / BOX_ROWS;
    constexpr unsigned BOX_BYTES = BOX_ROWS * CW * sizeof(float2);
    constexpr size_t TILE = ex_words<N, CW>();  // float2 elements per buffer (dense tile + exchange skew)
    extern __shared__ __align__(128) float2 smem2[];
    __shared__ __align__(8) unsigned long long full[NBUF];
    __shared__ __align__(8) unsigned long long wbar[NG];
    const int tid = threadIdx.x;
    const int g = tid / GT, gt = tid - g * GT;
    const int c = gt % CW, t = gt / CW;
    const GroupBarrier gbar{1 + g, GT};
    // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0, 1, ...
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        for (int q = 0; q < NG; ++q) mbar_init(&wbar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const int n0 = my_tiles < NBUF ? my_tiles : NBUF;
        for (int k = 0; k < n0; ++k) pipe_issue_tile<LOGN, CW, NBUF>(k, smem2, full, &tm_data, a.rows_valid, a.npairs, a.pair_base, tiles_x);
    }

    int wphase = 0;
    for (int k = g; k < my_tiles; k += NG) {
        const int b = k % NBUF;
        float2* ex = smem2 + (size_t)b * TILE;
        mbar_wait(&full[b], (unsigned)((k / NBUF) & 1));

        float2 v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int r = t + T * m;
            v[m] = (r < a.rows_valid) ? ex[(size_t)r * CW + c] : make_float2(0.f, 0.f);
        }
        fft_forward<N, CW, false, GroupBarrier>(v, ex, a.tw, t, c, gbar);

        gbar.sync();  // exchange buffer idle: bring in the Wiener tile
        if (gt == 0) {
            int xt, pr;
            pipe_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, a.npairs, xt, pr);
            mbar_expect_tx(&wbar[g], NBOX * BOX_BYTES);
            for (int q = 0; q < NBOX; ++q) tma_load_2d(ex + (size_t)q * BOX_ROWS * CW, &tm_w, xt * CW * 2, q * BOX_ROWS, &wbar[g]);
        }
        mbar_wait(&wbar[g], (unsigned)wphase);
        wphase ^= 1;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float2 y = cmul(v[m], ex[(size_t)(t + T * m) * CW + c]);
            v[m] = make_float2(y.x, -y.y);
        }
        fft_forward<N, CW, false, GroupBarrier>(v, ex, a.tw, t, c, gbar);

        gbar.sync();  // last exchange reads done
#pragma unroll
        for (int m = 0; m < E; ++m) ex[(size_t)(t + T * m) * CW + c] = v[m];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        gbar.sync();
        if (gt == 0) {
            int xt, pr;
            pipe_tile_coords<CW>(blockIdx.x + k * gridDim.x, tiles_x, a.npairs, xt, pr);
            for (int q = 0; q < NBOX; ++q) tma_store_2d(&tm_data, xt * CW * 2, (pr + a.pair_base) * N + q * BOX_ROWS, ex + (size_t)q * BOX_ROWS * CW);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (k + NBUF < my_tiles) pipe_issue_tile<LOGN, CW, NBUF>(k + NBUF, smem2, full, &tm_data, a.rows_valid, a.npairs, a.pair_base, tiles_x);
        }
    }
}

// geometry only (no pointers): can a COL_WIENER pass of length n over planes of this pitch use the TMA kernels?
bool col_tma_geometry_ok(int n, int pitch, long long cplane) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* env = getenv("FDR_COL_TMA");
        enabled = (env && atoi(env) == 0) ? 0 : 1;
    }
    if (!enabled) return false;
    if (n < 256 || n > 4096 || (n & (n - 1))) return false;
    const int cw = col_pass_tile_width(n);  // (the 1024 case falls back to this width when 4 does not divide the pitch)
    if (pitch % cw != 0 || cplane != (long long)n * pitch) return false;
    return tma_get_encode() != nullptr;
}

bool col_tma_applicable(const ColPassArgs& a) {
    if (a.mode != COL_WIENER || a.conj) return false;
    if (!col_tma_geometry_ok(a.n, a.pitch, a.cplane)) return false;
    return !((reinterpret_cast<uintptr_t>(a.data) & 15) || (reinterpret_cast<uintptr_t>(a.wiener) & 15));
}

template <int LOGN, int CW = default_col_cw(LOGN)> static cudaError_t launch_t(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColGeom<LOGN, CW>;
    constexpr int BOX_ROWS = (Gm::N < 256) ? Gm::N : 256;
    CUtensorMap tm_data, tm_w;
    if (!tma_make_map(&tm_data, a.data, (long long)(a.pair_base + a.npairs) * a.n, a.pitch, CW, BOX_ROWS)) return cudaErrorInvalidValue;
    if (!tma_make_map(&tm_w, a.wiener, a.n, a.pitch, CW, BOX_ROWS)) return cudaErrorInvalidValue;
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(col_wiener_tma_kernel<LOGN, CW>), Gm::SMEM);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(a.pitch / CW, a.npairs);
    col_wiener_tma_kernel<LOGN, CW><<<grid, Gm::THREADS, Gm::SMEM, s>>>(tm_data, tm_w, a);
    return cudaGetLastError();
}

template <int LOGN, int CW = default_col_cw(LOGN)> static cudaError_t launch_pipe_t(const ColPassArgs& a, cudaStream_t s) {
    using Gm = ColGeom<LOGN, CW>;
    using Pg = ColPipeGeom<LOGN, CW>;
    constexpr int BOX_ROWS = (Gm::N < 256) ? Gm::N : 256;
    CUtensorMap tm_data, tm_w;
    if (!tma_make_map(&tm_data, a.data, (long long)(a.pair_base + a.npairs) * a.n, a.pitch, CW, BOX_ROWS)) return cudaErrorInvalidValue;
    if (!tma_make_map(&tm_w, a.wiener, a.n, a.pitch, CW, BOX_ROWS)) return cudaErrorInvalidValue;
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(col_wiener_pipe_kernel<LOGN, CW>), Pg::SMEM);
        if (e != cudaSuccess) return e;
    }
    const int nsm = device_sm_count();
    const int tiles_x = a.pitch / CW, ntiles = tiles_x * a.npairs;
    const int grid = ntiles < nsm ? ntiles : nsm;
    col_wiener_pipe_kernel<LOGN, CW><<<grid, Pg::THREADS, Pg::SMEM, s>>>(tm_data, tm_w, a, tiles_x, ntiles);
    return cudaGetLastError();
}

// Persistent pipelined variant: chosen when the launch has several tiles per SM (otherwise the
// per-tile kernel's finer granularity wins).  FDR_COL_PIPE=0 disables it.
static bool col_pipe_wanted(const ColPassArgs& a) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* env = getenv("FDR_COL_PIPE");
        enabled = (env && atoi(env) == 0) ? 0 : 1;
    }
    if (a.col_variant == 3) return true;
    if (a.col_variant == 2 || !enabled) return false;
    if (a.n != 2048 && a.n != 4096) return false;
    const long long ntiles = (long long)(a.pitch / col_pass_tile_width(a.n)) * a.npairs;
    return ntiles >= 4 * 148;
}

cudaError_t launch_col_wiener_tma(const ColPassArgs& a, cudaStream_t s) {
    if (col_pipe_wanted(a)) {
        switch (a.n) {
            case 1024: if (a.pitch % 4 == 0) return launch_pipe_t<10, 4>(a, s); break;
            case 2048: return launch_pipe_t<11>(a, s);
            case 4096: return launch_pipe_t<12>(a, s);
        }
    }
    switch (a.n) {
        case 256: return launch_t<8>(a, s);
        case 512: return launch_t<9>(a, s);
        case 1024: return (a.pitch % 4 == 0) ? launch_t<10, 4>(a, s) : launch_t<10>(a, s);  // 32 KB tiles, 4 CTAs/SM: measured best
        case 2048: return launch_t<11>(a, s);
        case 4096: return launch_t<12>(a, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdr
