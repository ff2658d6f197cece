// shard.cu -- one rank's share of a ROW-SHARDED restoration of a single large image over
// `world` GPUs (BASELINE configs[4]: 16384 x 16384 over 2/4/8 B200).
//
// Decomposition (the reference's MPI mode is the behavioural model,
// /root/reference/fft/fft_mpi.cpp:89-100 row slabs, :170-279 all-to-all transpose, :311-470):
//   rank g owns padded rows [g*Rp/G, (g+1)*Rp/G) in the row passes and padded columns
//   [g*Cp/G, (g+1)*Cp/G) in the column pass.  Every rank allocates one column slab
//   [pairs][Rp][Cp/G] (complex).
//     phase 1  local rows: load u8, FFT along x, and STORE each element straight into the column
//              slab of the GPU that owns its column (peer-mapped pointer, NVLink) -- the
//              all-to-all transpose is the epilogue of the row pass, not a separate collective;
//     barrier  (caller: any stream-ordered cross-rank barrier, e.g. an NCCL all-reduce)
//     phase 2  local column slab: FFT along y, Wiener factor, inverse FFT along y, in place;
//     barrier
//     phase 3  local rows again: LOAD every element from the owning GPU's slab (peer reads),
//              inverse FFT along x, split the two packed planes, local min/max;
//     all-reduce(min), all-reduce(max) of 2 floats per plane (caller; doubles as the barrier
//              that protects the slabs before the next image)
//     phase 4  normalise + pack the local rows.
//   2 exchanges per plane pair instead of the reference's 6 MPI_Alltoallv per channel
//   (fft_mpi.cpp:386,393,423).  No scatter/gather to a root: every rank reads and writes only
//   its own rows of the image.
// HALF-PLANE MODE (default when the geometry allows it; FDR_SHARD_HALF=0 selects the plane-pair form above): every colour
// plane is its own pipeline unit.  Phase 1 packs local rows y and y + D of ONE plane into a complex row transform, untangles
// the two Hermitian row spectra and scatters only columns 0 .. Cp/2-1 (slab [C][Rp][Cl/2] per rank) plus the real Nyquist
// column (a vector [Rp] on rank `plane % world`); phase 2 runs on half the columns; phase 3 rebuilds the packed rows from
// the half planes.  For BGR that is 3 x 1/2 = 1.5 complex planes through NVLink and through the column phase instead of 2
// (-25 %), and three equal units that the driver (fdr_dist.ShardedRestorer) runs as a pipeline: unit u's NVLink-bound row
// phases overlap unit u-1's HBM-bound column phase.  The API keeps its names: a "pair" index is the unit index.
// Several shards may live in one process on one device (peers = plain device pointers): that is
// how the single-GPU test suite exercises this code.
#include <cstring>
#include <new>
#include <vector>

#include "capi_internal.h"
#include "tma_util.cuh"

using namespace fdr;

struct fdr_shard {
    int device = 0, rank = 0, world = 1;
    int H = 0, W = 0, C = 0, Rp = 0, Cp = 0;
    int Rl = 0, Cl = 0;          // padded rows / columns per rank
    int row0 = 0, rows_local = 0;  // image rows [row0, row0 + rows_local) live here
    int npairs = 0;
    float K = 0.f;
    bool have_wiener = false, have_peers = false;
    cudaStream_t stream = nullptr;
    DevBuf<float2> slab;      // [npairs][Rp][Cl]
    DevBuf<float2> wiener;    // [Rp][Cl]
    DevBuf<float> raw;        // [C][rows_local][W]
    DevBuf<unsigned int> mm;  // [C][FDR_MINMAX_SLOTS][2] ordered
    DevBuf<float> mmf;        // [C][2] floats: min, max (all-reduced by the caller)
    DevBuf<float2> ss;        // [C] scale, shift
    DevBuf<float> psf;
    DevBuf<float2*> peers;    // [world] device copy of the peer slab table
    DevBuf<float2*> self_only;  // [world] table with only this rank's slab (Wiener build)
    int psf_rows = 0, psf_cols = 0;
    const float2* tw_rows = nullptr;
    const float2* tw_cols = nullptr;
    long long launches = 0;
    bool col_split = false;
    // half-plane mode
    bool half = false;
    int units = 0;              // pipeline units: plane pairs, or planes in half-plane mode
    int Ch = 0;                 // columns of a half plane per rank = Cl / 2
    size_t nyq_off = 0;         // element offset of the Nyquist vectors [C][Rp] inside the slab allocation
    DevBuf<float2> wiener_nyq;  // [Rp]
    std::vector<float2*> peer_host;  // host copy of the peer slab table
    int row_ctas = 0;           // > 0: exchange passes run as at most this many persistent CTAs per unit
    int minmax_neg = 0;         // mmf holds (min, -max): one all-reduce(MIN) folds both
    // peer synchronisation area at the tail of the slab allocation (so the slab's IPC handle covers it): barrier flags
    // [SYNC_SETS][FDR_MAX_PEERS] u32, then the extrema mailbox [2][FDR_MAX_PEERS][C][2] f32, then one status word
    // STAGED mode (half-plane mode on more than one rank; FDR_SHARD_STAGED=0 keeps the fused form): the row passes work on
    // LOCAL staging planes [C][owner][Rl][Ch] (+ Nyquist [C][Rl]) and a separate link kernel on a few SMs pushes the
    // column blocks to / from the peers' slabs, so the NVLink-bound transfers run beside the HBM-bound passes of other units
    // instead of holding every SM (exchange1 / exchange3 below).  The staging planes live in the same allocation as the slab
    // because the peers store into them in exchange 3.
    bool staged = false;
    size_t stage_off = 0, stage_nyq_off = 0;   // element offsets inside the slab allocation
    int link_ctas = 24;
    int link_mode = 0;          // 0: link kernel (bulk copies from link_ctas CTAs); 1: the device's copy engines (cudaMemcpyAsync per block)
    int ce_streams = 4;         // copy-engine mode: the blocks of one exchange are spread over this many streams (engines)
    cudaStream_t st_ce[8] = {};
    cudaEvent_t ev_ce[9] = {};
    // native pipelined driver (fdr_shard_restore_rows): compute / link / barrier streams and the events between them
    cudaStream_t st_cmp = nullptr, st_link = nullptr, st_bar = nullptr;
    std::vector<cudaEvent_t> ev;               // [7 * units + 2]
    size_t sync_off = 0;        // element (float2) offset of the area inside the slab allocation
    unsigned int epoch[16] = {};  // barriers issued so far per flag set (every rank issues the same sequence)
    unsigned int mm_epoch = 0;
};

namespace {
cudaStream_t pick(fdr_shard* s, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : s->stream; }

// ---- cross-GPU synchronisation through peer memory (replaces MPI_Barrier / the implicit synchronisation of
// MPI_Alltoallv, fft_mpi.cpp:170-279, and the MPI_Allreduce-style min/max of a distributed normalize) ----
constexpr int SYNC_SETS = 16;
constexpr unsigned long long SYNC_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU
struct SyncPeers {
    unsigned int* area[FDR_MAX_PEERS];  // every rank's synchronisation area
};
size_t sync_area_elems(int channels) {  // in float2 units
    const size_t bytes = sizeof(unsigned int) * SYNC_SETS * FDR_MAX_PEERS + sizeof(float) * 2 * FDR_MAX_PEERS * channels * 2 + 64;
    return (bytes + 7) / 8;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// thread i: tell rank i that `rank` reached `epoch` of flag set `set`, then wait until rank i has told us the same
__device__ __forceinline__ void peer_signal_wait(const SyncPeers& sp, int rank, int world, int set, unsigned int epoch, unsigned int* status) {
    const int i = threadIdx.x;
    if (i < world) {
        __threadfence_system();  // everything this GPU wrote before (also by earlier kernels of the stream) precedes the flag
        st_release_sys(sp.area[i] + set * FDR_MAX_PEERS + rank, epoch);
        const unsigned int* mine = sp.area[rank] + set * FDR_MAX_PEERS + i;
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_ns() - t0 > SYNC_TIMEOUT_NS) {
                atomicExch(status, 1u);
                break;
            }
        }
    }
    __syncthreads();
}
__global__ void peer_barrier_kernel(SyncPeers sp, int rank, int world, int set, unsigned int epoch, unsigned int* status) {
    peer_signal_wait(sp, rank, world, set, epoch, status);
}
// All-reduce of the [C][2] extrema over peer memory in one launch: every rank drops its vector into every rank's mailbox
// (slot `epoch & 1`), barrier, fold.  negated: both columns fold with MIN (the vector holds (min, -max)).
__global__ void peer_minmax_kernel(SyncPeers sp, int rank, int world, int set, unsigned int epoch, unsigned int* status, float* mmf,
                                   int n, int negated) {
    const size_t box_off = (size_t)SYNC_SETS * FDR_MAX_PEERS;  // in 4-byte words
    const size_t slot = (size_t)(epoch & 1u) * FDR_MAX_PEERS * n;
    for (int j = threadIdx.x; j < world * n; j += blockDim.x) {
        const int dst = j / n, e = j - dst * n;
        float* box = reinterpret_cast<float*>(sp.area[dst] + box_off) + slot + (size_t)rank * n;
        box[e] = mmf[e];
    }
    __syncthreads();
    peer_signal_wait(sp, rank, world, set, epoch, status);
    const float* boxes = reinterpret_cast<const float*>(sp.area[rank] + box_off) + slot;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        float v = __ldcg(boxes + e);
        for (int r = 1; r < world; ++r) {
            const float w = __ldcg(boxes + (size_t)r * n + e);
            v = ((e & 1) && !negated) ? fmaxf(v, w) : fminf(v, w);
        }
        mmf[e] = v;
    }
}
SyncPeers sync_peers(const fdr_shard* s) {
    SyncPeers sp{};
    for (int i = 0; i < s->world; ++i) sp.area[i] = reinterpret_cast<unsigned int*>(s->peer_host[(size_t)i] + s->sync_off);
    return sp;
}
unsigned int* sync_status(const fdr_shard* s) {
    return reinterpret_cast<unsigned int*>(s->slab.p + s->sync_off) + SYNC_SETS * FDR_MAX_PEERS + 2 * FDR_MAX_PEERS * s->C * 2;
}

// half-plane forms: column owners' slabs and the Nyquist vectors (inside the same allocations) by value in the launch arguments
void fill_half_peers(const fdr_shard* s, RowPassArgs& r) {
    for (int i = 0; i < s->world; ++i) {
        r.hp_peers[i] = s->peer_host[(size_t)i];
        r.nyq_peers[i] = s->peer_host[(size_t)i] + s->nyq_off;
    }
    r.hp_shift = ilog2(s->Ch);
    r.hp_plane = (long long)s->Rp * s->Ch;
    r.nyq_world = s->world;
    r.nyq_plane = s->Rp;
}

// staged mode: the row passes see every column owner's block in LOCAL memory -- staging plane [unit][owner][Rl][Ch], row
// index = local row -- except this rank's own block, which is its slab itself (rows row0.. of it): nothing is copied for it.
// world * Rl = Rp, so both layouts share the unit stride Rp * Ch.
void fill_half_staged(const fdr_shard* s, RowPassArgs& r) {
    for (int g = 0; g < s->world; ++g)
        r.hp_peers[g] = (g == s->rank) ? s->slab.p + (size_t)s->row0 * s->Ch : s->slab.p + s->stage_off + (size_t)g * s->Rl * s->Ch;
    r.nyq_peers[0] = s->slab.p + s->stage_nyq_off;
    r.hp_shift = ilog2(s->Ch);
    r.hp_plane = (long long)s->Rp * s->Ch;
    r.hp_local = 1;
    r.nyq_world = 1;
    r.nyq_plane = s->Rl;
    r.row0 = 0;
}

// ---- link kernel: the all-to-all of the reference's MPI_Alltoallv (fft_mpi.cpp:170-279) as plain stores into peer memory,
// issued by a FEW persistent CTAs.  A job copies `rows` rows of `row_elems` complex values between two pitched planes. ----
struct PushJob {
    const float2* src;
    float2* dst;
    long long src_pitch, dst_pitch;   // elements
    int rows, row_shift;              // row_elems = 1 << row_shift
};
struct PushArgs {
    PushJob big[FDR_MAX_PEERS];       // equal shapes, one per destination rank
    int nbig;
    const float2* small_src[FDR_MAX_PEERS];   // contiguous runs (Nyquist vectors)
    float2* small_dst[FDR_MAX_PEERS];
    int small_n[FDR_MAX_PEERS];
    int nsmall;
};
// V = elements per access (2: 16-byte vectors, needs row_elems even and 16-byte aligned planes; 1 otherwise).  Work items are
// interleaved over the destinations (item = vector * nbig + job) so that every NVLink carries traffic all the time.
template <int V> __global__ void __launch_bounds__(512) peer_push_kernel(PushArgs a) {
    using Vec = typename std::conditional<V == 2, float4, float2>::type;
    if (blockIdx.x == 0) {
        for (int j = 0; j < a.nsmall; ++j)
            for (int i = threadIdx.x; i < a.small_n[j]; i += blockDim.x) a.small_dst[j][i] = a.small_src[j][i];
    }
    if (a.nbig == 0) return;
    const int vshift = a.big[0].row_shift - (V == 2 ? 1 : 0);      // vectors per row = 1 << vshift
    const long long per_job = (long long)a.big[0].rows << vshift;
    const long long total = per_job * a.nbig;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 8;
    long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i0 < total; i0 += stride * U) {
        Vec v[U];
        long long off[U];
        int jb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            jb[u] = -1;
            if (i < total) {
                const int j = (int)(i % a.nbig);
                const long long w = i / a.nbig;
                const long long row = w >> vshift, col = w & ((1LL << vshift) - 1);
                jb[u] = j;
                off[u] = row * a.big[j].dst_pitch + col * V;
                v[u] = *reinterpret_cast<const Vec*>(a.big[j].src + row * a.big[j].src_pitch + col * V);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (jb[u] >= 0) *reinterpret_cast<Vec*>(a.big[jb[u]].dst + off[u]) = v[u];
    }
}
// The same jobs moved by the bulk-copy engine (cp.async.bulk global -> shared -> peer global; SASS: UBLKCP): ONE thread per CTA
// keeps a ring of PUSH_STAGES buffers going, so a CTA has ~190 KB in flight instead of what its load/store queues hold.  Over
// NVLink an SM's plain stores stop at ~11 GB/s (measured, 8 x B200: 64 CTAs were needed for 700 GB/s); bulk copies take the
// SM's issue slots and registers out of the picture, which is what lets the exchange run on a few SMs beside the FFT passes.
constexpr int PUSH_STAGE_BYTES = 32768;
constexpr int PUSH_STAGES = 6;
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(128) peer_push_bulk_kernel(PushArgs a, int rows_per_item, int segs_per_row, unsigned seg_bytes,
                                                             long long items_per_job, int contig) {
    extern __shared__ __align__(128) unsigned char push_smem[];
    __shared__ unsigned long long bar[PUSH_STAGES];
    if (blockIdx.x == 0 && threadIdx.x >= 32) {
        for (int j = 0; j < a.nsmall; ++j)
            for (int i = threadIdx.x - 32; i < a.small_n[j]; i += blockDim.x - 32) a.small_dst[j][i] = a.small_src[j][i];
    }
    if (threadIdx.x != 0 || a.nbig == 0) return;
    for (int i = 0; i < PUSH_STAGES; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long total = items_per_job * a.nbig;
    const long long mine = (total > blockIdx.x) ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto decode = [&](long long k, int& j, long long& row0, int& nrows, long long& c0) {
        const long long idx = blockIdx.x + k * gridDim.x;
        j = (int)(idx % a.nbig);
        const long long w = idx / a.nbig;
        const long long rg = w / segs_per_row;
        c0 = (w - rg * segs_per_row) * (long long)(seg_bytes / 8);   // elements
        row0 = rg * rows_per_item;
        const long long left = a.big[j].rows - row0;
        nrows = left < rows_per_item ? (int)left : rows_per_item;
    };
    auto issue_load = [&](long long k) {
        int j, nrows;
        long long row0, c0;
        decode(k, j, row0, nrows, c0);
        const int st = (int)(k % PUSH_STAGES);
        mbar_expect_tx(&bar[st], (unsigned)nrows * seg_bytes);
        if (contig) {   // rows adjacent on both sides: the item is one run
            bulk_load_1d(push_smem + (size_t)st * PUSH_STAGE_BYTES, a.big[j].src + row0 * a.big[j].src_pitch, (unsigned)nrows * seg_bytes, &bar[st]);
            return;
        }
        for (int r = 0; r < nrows; ++r)
            bulk_load_1d(push_smem + (size_t)st * PUSH_STAGE_BYTES + (size_t)r * seg_bytes, a.big[j].src + (row0 + r) * a.big[j].src_pitch + c0,
                         seg_bytes, &bar[st]);
    };
    const long long pro = mine < PUSH_STAGES - 1 ? mine : PUSH_STAGES - 1;
    for (long long k = 0; k < pro; ++k) issue_load(k);
    for (long long k = 0; k < mine; ++k) {
        const int st = (int)(k % PUSH_STAGES);
        mbar_wait(&bar[st], (unsigned)((k / PUSH_STAGES) & 1));
        int j, nrows;
        long long row0, c0;
        decode(k, j, row0, nrows, c0);
        if (contig)
            bulk_store_1d(a.big[j].dst + row0 * a.big[j].dst_pitch, push_smem + (size_t)st * PUSH_STAGE_BYTES, (unsigned)nrows * seg_bytes);
        else
            for (int r = 0; r < nrows; ++r)
                bulk_store_1d(a.big[j].dst + (row0 + r) * a.big[j].dst_pitch + c0, push_smem + (size_t)st * PUSH_STAGE_BYTES + (size_t)r * seg_bytes,
                              seg_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (k + PUSH_STAGES - 1 < mine) {
            // the buffer of item k-1 is loaded next: its stores must have finished reading shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue_load(k + PUSH_STAGES - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the kernel (and the barrier after it) ends
}

// The same jobs on the copy engines: no SM, register or shared-memory footprint at all, so the passes of the other units keep
// the whole GPU (the link kernel's CTAs each take an SM away from kernels that need all of its registers).
cudaError_t push_copy_engine(fdr_shard* s, const PushArgs& a, cudaStream_t st) {
    // one stream keeps one engine busy (~360 GB/s measured over NVLink); the blocks go round-robin over `ce_streams` side
    // streams forked from and joined back into `st`, so the exchange still looks like one stream-ordered operation
    const int NL = s->ce_streams < 1 ? 1 : (s->ce_streams > 8 ? 8 : s->ce_streams);
    cudaError_t e = cudaSuccess;
    if (NL > 1) {
        if (!s->st_ce[0]) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaStreamCreateWithPriority(&s->st_ce[i], cudaStreamNonBlocking, hi);
            for (int i = 0; i < 9 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev_ce[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        e = cudaEventRecord(s->ev_ce[8], st);
        for (int i = 0; i < NL && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(s->st_ce[i], s->ev_ce[8], 0);
        if (e != cudaSuccess) return e;
    }
    int k = 0;
    for (int j = 0; j < a.nbig; ++j) {
        const PushJob& b = a.big[j];
        if (b.rows <= 0) continue;
        cudaStream_t q = NL > 1 ? s->st_ce[k++ % NL] : st;
        const size_t row_bytes = sizeof(float2) << b.row_shift;
        if (b.src_pitch == (1LL << b.row_shift) && b.dst_pitch == b.src_pitch)
            e = cudaMemcpyAsync(b.dst, b.src, row_bytes * b.rows, cudaMemcpyDeviceToDevice, q);
        else
            e = cudaMemcpy2DAsync(b.dst, (size_t)b.dst_pitch * sizeof(float2), b.src, (size_t)b.src_pitch * sizeof(float2), row_bytes, b.rows,
                                  cudaMemcpyDeviceToDevice, q);
        if (e != cudaSuccess) return e;
    }
    for (int j = 0; j < a.nsmall; ++j) {
        if (a.small_n[j] <= 0) continue;
        cudaStream_t q = NL > 1 ? s->st_ce[k++ % NL] : st;
        e = cudaMemcpyAsync(a.small_dst[j], a.small_src[j], sizeof(float2) * a.small_n[j], cudaMemcpyDeviceToDevice, q);
        if (e != cudaSuccess) return e;
    }
    if (NL > 1) {
        for (int i = 0; i < NL && e == cudaSuccess; ++i) {
            e = cudaEventRecord(s->ev_ce[i], s->st_ce[i]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, s->ev_ce[i], 0);
        }
    }
    return e;
}

cudaError_t launch_push(PushArgs& a, int ctas, cudaStream_t st) {
    static const bool use_bulk = !(getenv("FDR_SHARD_BULK") && atoi(getenv("FDR_SHARD_BULK")) == 0);
    if (use_bulk && a.nbig > 0 && a.big[0].row_shift >= 1) {
        const long long row_bytes = 8LL << a.big[0].row_shift;
        bool ok = true;
        for (int j = 0; j < a.nbig && ok; ++j)
            ok = ((reinterpret_cast<uintptr_t>(a.big[j].src) | reinterpret_cast<uintptr_t>(a.big[j].dst)) % 16 == 0) &&
                 a.big[j].src_pitch % 2 == 0 && a.big[j].dst_pitch % 2 == 0 && a.big[j].rows == a.big[0].rows;
        if (ok) {
            const unsigned seg_bytes = (unsigned)(row_bytes < PUSH_STAGE_BYTES ? row_bytes : PUSH_STAGE_BYTES);
            const int segs_per_row = (int)(row_bytes / seg_bytes);
            int rows_per_item = PUSH_STAGE_BYTES / (int)seg_bytes;
            bool contig = segs_per_row == 1;
            for (int j = 0; j < a.nbig && contig; ++j) contig = a.big[j].src_pitch == (1LL << a.big[0].row_shift) && a.big[j].dst_pitch == a.big[j].src_pitch;
            if (!contig && rows_per_item > 64) rows_per_item = 64;   // tiny rows (tests): bound the copies one thread issues per item
            const long long items_per_job = (long long)((a.big[0].rows + rows_per_item - 1) / rows_per_item) * segs_per_row;
            const size_t smem = (size_t)PUSH_STAGES * PUSH_STAGE_BYTES;
            cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(peer_push_bulk_kernel), smem);
            if (e != cudaSuccess) return e;
            long long total = items_per_job * a.nbig;
            int grid = ctas;
            if (total < grid) grid = total > 0 ? (int)total : 1;
            peer_push_bulk_kernel<<<grid, 128, smem, st>>>(a, rows_per_item, segs_per_row, seg_bytes, items_per_job, contig ? 1 : 0);
            return cudaGetLastError();
        }
    }
    bool v2 = a.nbig > 0 && a.big[0].row_shift >= 1;
    for (int j = 0; j < a.nbig && v2; ++j)
        v2 = ((reinterpret_cast<uintptr_t>(a.big[j].src) | reinterpret_cast<uintptr_t>(a.big[j].dst)) % 16 == 0) &&
             a.big[j].src_pitch % 2 == 0 && a.big[j].dst_pitch % 2 == 0;
    if (v2)
        peer_push_kernel<2><<<ctas, 512, 0, st>>>(a);
    else
        peer_push_kernel<1><<<ctas, 512, 0, st>>>(a);
    return cudaGetLastError();
}

int build_wiener(fdr_shard* s) {
    if (s->psf_rows > s->Rp || s->psf_cols > s->Cp)
        return set_error(FDR_E_INVALID, "PSF %dx%d larger than the padded image %dx%d", s->psf_rows, s->psf_cols, s->Rp, s->Cp);
    cudaStream_t st = s->stream;
    if (s->half) {
        // Row spectra of the PSF rows, keeping this rank's half-plane columns [rank*Ch, (rank+1)*Ch): the scatter table has one
        // entry per Ch columns (2*world of them), all NULL but ours.  The Nyquist column comes from its own tiny kernel.
        RowPassArgs r{};
        r.n = s->Cp;
        r.nrows = s->psf_rows;
        r.npairs = 1;
        r.in_mode = ROW_IN_PAIR_F32;
        r.out_mode = ROW_OUT_SCATTER;
        r.in_f32 = s->psf.p;
        r.in_unit_stride = (long long)s->psf_rows * s->psf_cols;
        r.in_row_stride = s->psf_cols;
        r.channels = 1;
        r.img_rows = s->psf_rows;
        r.img_cols = s->psf_cols;
        r.units_total = 1;
        r.tw = s->tw_rows;
        r.peers = s->self_only.p;
        r.peer_shift = ilog2(s->Ch);
        r.peer_plane = (long long)s->Rp * s->Ch;
        r.row0 = 0;
        FDR_CUDA(launch_row_pass(r, st));
        float2* nyq0 = s->slab.p + s->nyq_off;  // plane 0's Nyquist vector as scratch
        FDR_CUDA(launch_psf_nyquist(s->psf.p, s->psf_rows, s->psf_cols, nyq0, st));
        ColPassArgs c{};
        c.n = s->Rp;
        c.pitch = s->Ch;
        c.npairs = 1;
        c.mode = COL_MAKE_WIENER;
        c.rows_valid = s->psf_rows;
        c.data = s->slab.p;
        c.cplane = (long long)s->Rp * s->Ch;
        c.wiener_out = s->wiener.p;
        c.K = s->K;
        c.tw = s->tw_cols;
        if (s->col_split)
            FDR_CUDA(launch_col_split(c, st, nullptr));
        else
            FDR_CUDA(launch_col_pass(c, st));
        ColPassArgs n = c;
        n.pitch = 1;
        n.data = nyq0;
        n.cplane = s->Rp;
        n.wiener_out = s->wiener_nyq.p;
        FDR_CUDA(launch_col_pass(n, st));
        FDR_CUDA(cudaStreamSynchronize(st));
        s->have_wiener = true;
        return FDR_OK;
    }
    // PSF rows are few (S <= Rp): every rank transforms all of them and keeps its own columns.
    RowPassArgs r{};
    r.n = s->Cp;
    r.nrows = s->psf_rows;
    r.npairs = 1;
    r.in_mode = ROW_IN_PAIR_F32;
    r.out_mode = ROW_OUT_SCATTER;
    r.in_f32 = s->psf.p;
    r.in_unit_stride = (long long)s->psf_rows * s->psf_cols;
    r.in_row_stride = s->psf_cols;
    r.channels = 1;
    r.img_rows = s->psf_rows;
    r.img_cols = s->psf_cols;
    r.units_total = 1;
    r.tw = s->tw_rows;
    r.peers = s->self_only.p;
    r.peer_shift = ilog2(s->Cl);
    r.peer_plane = (long long)s->Rp * s->Cl;
    r.row0 = 0;
    FDR_CUDA(launch_row_pass(r, st));
    ColPassArgs c{};
    c.n = s->Rp;
    c.pitch = s->Cl;
    c.npairs = 1;
    c.mode = COL_MAKE_WIENER;
    c.rows_valid = s->psf_rows;
    c.data = s->slab.p;
    c.cplane = (long long)s->Rp * s->Cl;
    c.wiener_out = s->wiener.p;
    c.K = s->K;
    c.tw = s->tw_cols;
    if (s->col_split)
        FDR_CUDA(launch_col_split(c, st, nullptr));
    else
        FDR_CUDA(launch_col_pass(c, st));
    FDR_CUDA(cudaStreamSynchronize(st));
    s->have_wiener = true;
    return FDR_OK;
}
}  // namespace

FDR_API int fdr_shard_create(fdr_shard** out, int rows, int cols, int channels, int rank, int world, int device) {
    if (!out) return set_error(FDR_E_INVALID, "shard is NULL");
    *out = nullptr;
    if (rows < 1 || cols < 1 || channels < 1) return set_error(FDR_E_INVALID, "bad image geometry %dx%dx%d", rows, cols, channels);
    if (world < 1 || !is_pow2(world) || rank < 0 || rank >= world)
        return set_error(FDR_E_INVALID, "world=%d must be a power of two and 0 <= rank=%d < world", world, rank);
    const int Rp = next_pow2(rows), Cp = next_pow2(cols);
    if (Rp > 16384 || Cp > 16384) return set_error(FDR_E_INVALID, "padded size above 16384 is not supported");
    if (Rp / world < 1 || Cp / world < 1) return set_error(FDR_E_INVALID, "image %dx%d too small to split over %d ranks", Rp, Cp, world);
    FDR_CUDA(cudaSetDevice(device));
    fdr_shard* s = new (std::nothrow) fdr_shard();
    if (!s) return set_error(FDR_E_NOMEM, "out of host memory");
    s->device = device;
    s->rank = rank;
    s->world = world;
    s->H = rows;
    s->W = cols;
    s->C = channels;
    s->Rp = Rp;
    s->Cp = Cp;
    s->Rl = Rp / world;
    s->Cl = Cp / world;
    s->row0 = rank * s->Rl;
    int r1 = s->row0 + s->Rl;
    if (r1 > rows) r1 = rows;
    s->rows_local = r1 > s->row0 ? r1 - s->row0 : 0;
    s->npairs = (channels + 1) / 2;
    {
        const char* hv = getenv("FDR_SHARD_HALF");
        s->half = !(hv && atoi(hv) == 0) && Cp >= FDR_HALF_MIN_N && s->Cl >= 2 && s->Rl >= 2 && world <= FDR_MAX_PEERS;
        s->Ch = s->Cl / 2;
        s->units = s->half ? channels : s->npairs;
        const char* rc = getenv("FDR_SHARD_ROW_CTAS");
        s->row_ctas = (rc && atoi(rc) > 0) ? atoi(rc) : 0;
    }
    {
        ColPassArgs probe{};
        probe.n = Rp;
        probe.pitch = s->half ? s->Ch : Cp / world;
        probe.mode = COL_WIENER;
        const char* cs = getenv("FDR_COL_SPLIT");
        s->col_split = col_split_applicable(probe) && !(cs && atoi(cs) == 0);
    }
    int rc = FDR_OK;
    cudaError_t e = get_twiddles(Cp, &s->tw_rows);
    if (e == cudaSuccess) e = get_twiddles(Rp, &s->tw_cols);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "shard setup: %s", cudaGetErrorString(e));
    if (s->half) {
        s->nyq_off = (size_t)channels * Rp * s->Ch;
        size_t end = s->nyq_off + (size_t)channels * Rp;
        const char* sg = getenv("FDR_SHARD_STAGED");
        s->staged = world > 1 && !(sg && atoi(sg) == 0);
        if (s->staged) {
            s->stage_off = (end + 31) & ~(size_t)31;
            s->stage_nyq_off = s->stage_off + (size_t)channels * Rp * s->Ch;   // [C][owner][Rl][Ch]
            end = s->stage_nyq_off + (size_t)channels * s->Rl;
        }
        const char* lm = getenv("FDR_SHARD_LINK");   // "ce": copy engines, "bulk": link kernel
        if (lm) s->link_mode = (lm[0] == 'c') ? 1 : 0;
        const char* lc = getenv("FDR_SHARD_LINK_CTAS");
        if (lc && atoi(lc) > 0) s->link_ctas = atoi(lc);
        s->sync_off = (end + 31) & ~(size_t)31;
        if (rc == FDR_OK) rc = s->slab.ensure(s->sync_off + sync_area_elems(channels));
        if (rc == FDR_OK) rc = s->wiener.ensure((size_t)Rp * s->Ch);
        if (rc == FDR_OK) rc = s->wiener_nyq.ensure((size_t)Rp);
    } else {
        s->sync_off = ((size_t)s->npairs * Rp * s->Cl + 31) & ~(size_t)31;
        if (rc == FDR_OK) rc = s->slab.ensure(s->sync_off + sync_area_elems(channels));
        if (rc == FDR_OK) rc = s->wiener.ensure((size_t)Rp * s->Cl);
    }
    if (rc == FDR_OK && world > FDR_MAX_PEERS) rc = set_error(FDR_E_INVALID, "at most %d ranks", FDR_MAX_PEERS);
    if (rc == FDR_OK) {
        // Load the waiting kernels now: with lazy module loading the first launch of a kernel may block the host until the
        // device is idle, which never happens while another shard of this process spins in a barrier.
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(peer_barrier_kernel));
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(peer_minmax_kernel));
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(peer_push_kernel<1>));
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(peer_push_kernel<2>));
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(peer_push_bulk_kernel));
        if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "sync kernels: %s", cudaGetErrorString(e));
    }
    if (rc == FDR_OK) {  // flags and mailboxes start at zero (before any peer can have the handle)
        e = cudaMemset(s->slab.p + s->sync_off, 0, sync_area_elems(channels) * sizeof(float2));
        if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "sync area: %s", cudaGetErrorString(e));
    }
    if (rc == FDR_OK) rc = s->raw.ensure((size_t)channels * (s->rows_local > 0 ? s->rows_local : 1) * cols);
    if (rc == FDR_OK) rc = s->mm.ensure((size_t)channels * 2 * FDR_MINMAX_SLOTS);
    if (rc == FDR_OK) rc = s->mmf.ensure((size_t)channels * 2);
    if (rc == FDR_OK) rc = s->ss.ensure((size_t)channels);
    if (rc == FDR_OK) rc = s->peers.ensure((size_t)world);
    if (rc == FDR_OK) rc = s->self_only.ensure((size_t)2 * world);
    if (rc == FDR_OK) {
        std::vector<float2*> tbl((size_t)2 * world, nullptr);  // (half-plane Wiener build: one entry per Cl/2 columns)
        tbl[(size_t)rank] = s->slab.p;
        e = cudaMemcpy(s->self_only.p, tbl.data(), sizeof(float2*) * 2 * world, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = set_error(FDR_E_CUDA, "peer table: %s", cudaGetErrorString(e));
    }
    if (rc != FDR_OK) {
        fdr_shard_destroy(s);
        return rc;
    }
    *out = s;
    return FDR_OK;
}

FDR_API int fdr_shard_destroy(fdr_shard* s) {
    if (!s) return FDR_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    s->slab.release();
    s->wiener.release();
    s->wiener_nyq.release();
    s->raw.release();
    s->mm.release();
    s->mmf.release();
    s->ss.release();
    s->psf.release();
    s->peers.release();
    s->self_only.release();
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : s->ev_ce)
        if (e) cudaEventDestroy(e);
    for (auto& q : s->st_ce)
        if (q) cudaStreamDestroy(q);
    if (s->st_cmp) cudaStreamDestroy(s->st_cmp);
    if (s->st_link) cudaStreamDestroy(s->st_link);
    if (s->st_bar) cudaStreamDestroy(s->st_bar);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return FDR_OK;
}

FDR_API int fdr_shard_geometry(const fdr_shard* s, int* first_row, int* n_rows, int* padded_rows, int* padded_cols,
                               int* cols_per_rank) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (first_row) *first_row = s->row0;
    if (n_rows) *n_rows = s->rows_local;
    if (padded_rows) *padded_rows = s->Rp;
    if (padded_cols) *padded_cols = s->Cp;
    if (cols_per_rank) *cols_per_rank = s->Cl;
    return FDR_OK;
}

FDR_API int fdr_shard_local_slab(const fdr_shard* s, void** d_slab, size_t* bytes) {
    if (!s || !d_slab) return set_error(FDR_E_INVALID, "bad arguments");
    *d_slab = s->slab.p;
    if (bytes) *bytes = s->slab.n * sizeof(float2);
    return FDR_OK;
}

FDR_API int fdr_ipc_export(const void* dptr, unsigned char handle[64]) {
    if (!dptr || !handle) return set_error(FDR_E_INVALID, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    FDR_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
    memcpy(handle, &h, 64);
    return FDR_OK;
}

FDR_API int fdr_ipc_open(const unsigned char handle[64], void** dptr) {
    if (!dptr || !handle) return set_error(FDR_E_INVALID, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    FDR_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FDR_OK;
}

FDR_API int fdr_ipc_close(void* dptr) {
    if (dptr) FDR_CUDA(cudaIpcCloseMemHandle(dptr));
    return FDR_OK;
}

FDR_API int fdr_shard_set_peers(fdr_shard* s, void* const* slabs) {
    if (!s || !slabs) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_CUDA(cudaSetDevice(s->device));
    std::vector<float2*> tbl((size_t)s->world);
    for (int i = 0; i < s->world; ++i) {
        tbl[(size_t)i] = (i == s->rank) ? s->slab.p : static_cast<float2*>(slabs[i]);
        if (!tbl[(size_t)i]) return set_error(FDR_E_INVALID, "peer %d slab is NULL", i);
    }
    FDR_CUDA(cudaMemcpy(s->peers.p, tbl.data(), sizeof(float2*) * s->world, cudaMemcpyHostToDevice));
    s->peer_host = tbl;
    s->have_peers = true;
    return FDR_OK;
}

FDR_API int fdr_shard_set_psf_motion(fdr_shard* s, int length, double angle_deg, float K) {
    if (!s || length < 1) return set_error(FDR_E_INVALID, "bad PSF arguments");
    FDR_CUDA(cudaSetDevice(s->device));
    FDR_TRY(s->psf.ensure((size_t)length * length));
    FDR_CUDA(launch_motion_psf(s->psf.p, length, motion_affine(length, angle_deg), s->stream));
    s->psf_rows = s->psf_cols = length;
    s->K = K;
    return build_wiener(s);
}

FDR_API int fdr_shard_set_psf_host(fdr_shard* s, const float* psf, int psf_rows, int psf_cols, float K) {
    if (!s || !psf || psf_rows < 1 || psf_cols < 1) return set_error(FDR_E_INVALID, "bad PSF arguments");
    FDR_CUDA(cudaSetDevice(s->device));
    FDR_TRY(s->psf.ensure((size_t)psf_rows * psf_cols));
    FDR_CUDA(cudaMemcpy(s->psf.p, psf, sizeof(float) * (size_t)psf_rows * psf_cols, cudaMemcpyHostToDevice));
    s->psf_rows = psf_rows;
    s->psf_cols = psf_cols;
    s->K = K;
    return build_wiener(s);
}

// phase 1: rows forward + scatter to the owners of the columns.  d_in_rows: this rank's rows of
// the image, interleaved u8 [rows_local][W][C].
static int check_pairs(const fdr_shard* s, int pair_first, int pair_count) {
    if (pair_first < 0 || pair_count < 1 || pair_first + pair_count > s->units)
        return set_error(FDR_E_INVALID, "unit range [%d, +%d) outside 0..%d", pair_first, pair_count, s->units);
    return FDR_OK;
}

FDR_API int fdr_shard_phase1_rows(fdr_shard* s, const void* d_in_rows_u8, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    return fdr_shard_phase1_pairs(s, d_in_rows_u8, 0, s->units, stream);
}

FDR_API int fdr_shard_phase1_pairs(fdr_shard* s, const void* d_in_rows_u8, int pair_first, int pair_count, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (!s->have_peers || !s->have_wiener) return set_error(FDR_E_STATE, "set peers and PSF before phase 1");
    FDR_TRY(check_pairs(s, pair_first, pair_count));
    FDR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = pick(s, stream);
    if (pair_first == 0) s->launches = 0;
    {   // extrema of the planes these units carry
        const int per = s->half ? 1 : 2;
        const int u0 = per * pair_first;
        int nu = per * pair_count;
        if (u0 + nu > s->C) nu = s->C - u0;
        FDR_CUDA(launch_minmax_reset(s->mm.p + (size_t)2 * FDR_MINMAX_SLOTS * u0, nu, st));
        s->launches += 1;
    }
    if (s->rows_local == 0) return FDR_OK;  // slab entirely inside the zero padding
    if (!d_in_rows_u8) return set_error(FDR_E_INVALID, "input rows are NULL");
    if (s->half) {
        RowPassArgs r{};
        r.n = s->Cp;
        r.pair_dist = (s->rows_local + 1) / 2;
        r.nrows = r.pair_dist;
        r.npairs = pair_count;
        r.pair_base = pair_first;
        r.in_mode = ROW_IN_ROWS2_U8;
        r.out_mode = ROW_OUT_HALF;
        r.in_u8 = static_cast<const uint8_t*>(d_in_rows_u8);
        r.channels = s->C;
        r.img_rows = s->rows_local;
        r.img_cols = s->W;
        r.rows_in = s->rows_local;
        r.hp_rows_store = s->H;
        r.unit_base = 0;
        r.units_total = s->C;
        r.tw = s->tw_rows;
        if (s->staged) {   // local full-width half planes; fdr_shard_exchange1 moves them
            fill_half_staged(s, r);
            r.hp_rows_store = s->rows_local;
        } else {
            fill_half_peers(s, r);
            r.row0 = s->row0;
            r.max_ctas = s->row_ctas;
        }
        FDR_CUDA(launch_row_pass(r, st));
        s->launches += 1;
        return FDR_OK;
    }
    RowPassArgs r{};
    r.n = s->Cp;
    r.nrows = s->rows_local;
    r.npairs = pair_count;
    r.pair_base = pair_first;
    r.in_mode = ROW_IN_PAIR_U8;
    r.out_mode = ROW_OUT_SCATTER;
    r.in_u8 = static_cast<const uint8_t*>(d_in_rows_u8);
    r.channels = s->C;
    r.img_rows = s->rows_local;
    r.img_cols = s->W;
    r.unit_base = 0;
    r.units_total = s->C;
    r.tw = s->tw_rows;
    r.peers = s->peers.p;
    r.peer_shift = ilog2(s->Cl);
    r.peer_plane = (long long)s->Rp * s->Cl;
    r.row0 = s->row0;
    r.max_ctas = s->row_ctas;
    FDR_CUDA(launch_row_pass(r, st));
    s->launches += 1;
    return FDR_OK;
}

// phase 2: columns of the local slab, in place (FFT, Wiener factor, inverse FFT).
FDR_API int fdr_shard_phase2_cols(fdr_shard* s, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    return fdr_shard_phase2_pairs(s, 0, s->units, stream);
}

FDR_API int fdr_shard_phase2_pairs(fdr_shard* s, int pair_first, int pair_count, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    FDR_TRY(check_pairs(s, pair_first, pair_count));
    FDR_CUDA(cudaSetDevice(s->device));
    ColPassArgs c{};
    c.n = s->Rp;
    c.pitch = s->half ? s->Ch : s->Cl;
    c.npairs = pair_count;
    c.pair_base = pair_first;
    c.mode = COL_WIENER;
    c.rows_valid = s->H;
    c.data = s->slab.p;
    c.cplane = (long long)s->Rp * c.pitch;
    c.wiener = s->wiener.p;
    c.K = s->K;
    c.tw = s->tw_cols;
    if (s->col_split) {
        int nl = 0;
        FDR_CUDA(launch_col_split(c, pick(s, stream), &nl));
        s->launches += nl;
    } else {
        FDR_CUDA(launch_col_pass(c, pick(s, stream)));
        s->launches += 1;
    }
    if (s->half) {
        // the Nyquist columns this rank owns: one more column of the same problem each (plain column kernel, natural order)
        for (int u = pair_first; u < pair_first + pair_count; ++u) {
            if (u % s->world != s->rank) continue;
            ColPassArgs n = c;
            n.pitch = 1;
            n.npairs = 1;
            n.pair_base = 0;
            n.data = s->slab.p + s->nyq_off + (size_t)u * s->Rp;
            n.cplane = s->Rp;
            n.wiener = s->wiener_nyq.p;
            n.wiener_tiled = nullptr;
            FDR_CUDA(launch_col_pass(n, pick(s, stream)));
            s->launches += 1;
        }
    }
    return FDR_OK;
}

// phase 3: gather the local padded rows from every slab, inverse rows, min/max of the local part.
FDR_API int fdr_shard_phase3_rows(fdr_shard* s, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    return fdr_shard_phase3_pairs(s, 0, s->units, stream);
}

FDR_API int fdr_shard_phase3_pairs(fdr_shard* s, int pair_first, int pair_count, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    FDR_TRY(check_pairs(s, pair_first, pair_count));
    FDR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = pick(s, stream);
    RowPassArgs r{};
    r.n = s->Cp;
    r.nrows = s->half ? s->Rl / 2 : s->Rl;
    r.npairs = pair_count;
    r.pair_base = pair_first;
    r.in_mode = s->half ? ROW_IN_HALF : ROW_IN_GATHER;
    r.out_mode = s->half ? ROW_OUT_REAL_ROWS2 : ROW_OUT_REAL_PAIR;
    r.pair_dist = s->Rl / 2;
    if (s->half) fill_half_peers(s, r);
    r.unit_base = 0;
    r.units_total = s->C;
    r.raw = s->raw.p;
    r.raw_unit_stride = (long long)(s->rows_local > 0 ? s->rows_local : 1) * s->W;
    r.raw_rows = s->rows_local;
    r.raw_cols = s->W;
    r.minmax = s->mm.p;
    r.local_units = s->C;
    r.tw = s->tw_rows;
    r.peers = s->peers.p;
    r.peer_shift = ilog2(s->Cl);
    r.peer_plane = (long long)s->Rp * s->Cl;
    r.row0 = s->row0;
    r.max_ctas = s->row_ctas;
    if (s->staged) {   // the peers' exchange 3 has filled the local staging planes
        fill_half_staged(s, r);
        r.max_ctas = 0;
    }
    FDR_CUDA(launch_row_pass(r, st));
    {
        const int per = s->half ? 1 : 2;
        const int u0 = per * pair_first;
        int nu = per * pair_count;
        if (u0 + nu > s->C) nu = s->C - u0;
        FDR_CUDA(launch_minmax_decode(s->mm.p + (size_t)2 * FDR_MINMAX_SLOTS * u0, s->mmf.p + 2 * u0, nu, s->minmax_neg, st));
    }
    s->launches += 2;
    return FDR_OK;
}

// [channels][2] floats on the device (min, max of the local rows of every padded plane).  The
// caller all-reduces column 0 with MIN and column 1 with MAX across ranks, in place.
FDR_API int fdr_shard_minmax_device(fdr_shard* s, void** d_minmax_f32) {
    if (!s || !d_minmax_f32) return set_error(FDR_E_INVALID, "bad arguments");
    *d_minmax_f32 = s->mmf.p;
    return FDR_OK;
}

// phase 4: normalise with the global extrema and pack this rank's rows: u8 [rows_local][W][C].
FDR_API int fdr_shard_phase4_pack(fdr_shard* s, void* d_out_rows_u8, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    FDR_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = pick(s, stream);
    FDR_CUDA(launch_scale_shift_from_f32(s->mmf.p, s->ss.p, s->C, s->minmax_neg, st));
    s->launches += 1;
    if (s->rows_local == 0) return FDR_OK;
    if (!d_out_rows_u8) return set_error(FDR_E_INVALID, "output rows are NULL");
    FDR_CUDA(launch_pack_u8(s->raw.p, (long long)s->rows_local * s->W, s->ss.p, static_cast<uint8_t*>(d_out_rows_u8), 1, s->C,
                            s->rows_local, s->W, st));
    s->launches += 1;
    return FDR_OK;
}

// Staged mode only (no-ops otherwise).  exchange 1: this rank's row spectra (local staging planes, written by phase 1) ->
// the column owners' slabs; exchange 3: this rank's filtered columns (its slab, after phase 2) -> the row owners' staging
// planes.  Plain stores over NVLink from `link_ctas` CTAs; a barrier must follow before the data is consumed.
FDR_API int fdr_shard_exchange1(fdr_shard* s, int unit_first, int unit_count, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (!s->staged) return FDR_OK;
    if (!s->have_peers) return set_error(FDR_E_STATE, "exchange before fdr_shard_set_peers");
    FDR_TRY(check_pairs(s, unit_first, unit_count));
    FDR_CUDA(cudaSetDevice(s->device));
    if (s->rows_local == 0) return FDR_OK;
    for (int u = unit_first; u < unit_first + unit_count; ++u) {
        PushArgs a{};
        for (int g = 1; g < s->world; ++g) {
            const int d = (g + s->rank) % s->world;   // every rank starts on a different link; our own block is already in place
            PushJob& j = a.big[a.nbig++];
            j.src = s->slab.p + s->stage_off + (size_t)u * s->Rp * s->Ch + (size_t)d * s->Rl * s->Ch;
            j.dst = s->peer_host[(size_t)d] + (size_t)u * s->Rp * s->Ch + (size_t)s->row0 * s->Ch;
            j.src_pitch = s->Ch;
            j.dst_pitch = s->Ch;
            j.rows = s->rows_local;
            j.row_shift = ilog2(s->Ch);
        }
        const int owner = u % s->world;
        a.small_src[0] = s->slab.p + s->stage_nyq_off + (size_t)u * s->Rl;
        a.small_dst[0] = s->peer_host[(size_t)owner] + s->nyq_off + (size_t)u * s->Rp + s->row0;
        a.small_n[0] = s->rows_local;
        a.nsmall = 1;
        if (s->link_mode == 1)
            FDR_CUDA(push_copy_engine(s, a, pick(s, stream)));
        else
            FDR_CUDA(launch_push(a, s->link_ctas, pick(s, stream)));
        s->launches += 1;
    }
    return FDR_OK;
}

FDR_API int fdr_shard_exchange3(fdr_shard* s, int unit_first, int unit_count, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (!s->staged) return FDR_OK;
    if (!s->have_peers) return set_error(FDR_E_STATE, "exchange before fdr_shard_set_peers");
    FDR_TRY(check_pairs(s, unit_first, unit_count));
    FDR_CUDA(cudaSetDevice(s->device));
    for (int u = unit_first; u < unit_first + unit_count; ++u) {
        PushArgs a{};
        for (int g = 0; g < s->world; ++g) {
            const int d = (g + s->rank) % s->world;
            if (g > 0) {   // (g == 0: our own rows of our own columns stay in the slab, where phase 3 reads them)
                PushJob& j = a.big[a.nbig++];
                j.src = s->slab.p + (size_t)u * s->Rp * s->Ch + (size_t)d * s->Rl * s->Ch;
                j.dst = s->peer_host[(size_t)d] + s->stage_off + (size_t)u * s->Rp * s->Ch + (size_t)s->rank * s->Rl * s->Ch;
                j.src_pitch = s->Ch;
                j.dst_pitch = s->Ch;
                j.rows = s->Rl;
                j.row_shift = ilog2(s->Ch);
            }
            if (u % s->world == s->rank) {   // the Nyquist column of this unit lives here
                a.small_src[a.nsmall] = s->slab.p + s->nyq_off + (size_t)u * s->Rp + (size_t)d * s->Rl;
                a.small_dst[a.nsmall] = s->peer_host[(size_t)d] + s->stage_nyq_off + (size_t)u * s->Rl;
                a.small_n[a.nsmall] = s->Rl;
                a.nsmall++;
            }
        }
        if (s->link_mode == 1)
            FDR_CUDA(push_copy_engine(s, a, pick(s, stream)));
        else
            FDR_CUDA(launch_push(a, s->link_ctas, pick(s, stream)));
        s->launches += 1;
    }
    return FDR_OK;
}

FDR_API int fdr_shard_set_ce_streams(fdr_shard* s, int streams) {
    if (!s || streams < 1 || streams > 8) return set_error(FDR_E_INVALID, "1..8 copy streams");
    s->ce_streams = streams;
    return FDR_OK;
}

FDR_API int fdr_shard_staged(const fdr_shard* s, int* enabled) {
    if (!s || !enabled) return set_error(FDR_E_INVALID, "bad arguments");
    *enabled = s->staged ? 1 : 0;
    return FDR_OK;
}

FDR_API int fdr_shard_set_link_ctas(fdr_shard* s, int ctas) {
    if (!s || ctas < 0) return set_error(FDR_E_INVALID, "bad arguments");
    if (ctas == 0) {   // 0 selects the copy engines
        s->link_mode = 1;
    } else {
        s->link_mode = 0;
        s->link_ctas = ctas;
    }
    return FDR_OK;
}

// The whole restoration of this rank's rows as a pipeline over the units (planes), issued in one call: compute passes on one
// stream, link kernels and barriers on two high-priority streams, events in between -- unit u's transfers run beside the
// passes of the other units.  Every rank must call it (it contains the cross-rank barriers of fdr_shard_barrier).
// Replaces the per-channel loop of mpi.cpp:95-111 with its blocking MPI_Alltoallv calls (fft_mpi.cpp:284-307).
FDR_API int fdr_shard_restore_rows(fdr_shard* s, const void* d_in_rows_u8, void* d_out_rows_u8, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (!s->have_peers || !s->have_wiener) return set_error(FDR_E_STATE, "set peers and PSF before restoring");
    if (2 * s->units > SYNC_SETS - 3) return set_error(FDR_E_INVALID, "too many units (%d) for the pipelined driver", s->units);
    FDR_CUDA(cudaSetDevice(s->device));
    const int U = s->units;
    if (!s->st_cmp) {
        int lo = 0, hi = 0;
        FDR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        FDR_CUDA(cudaStreamCreateWithPriority(&s->st_cmp, cudaStreamNonBlocking, lo));
        FDR_CUDA(cudaStreamCreateWithPriority(&s->st_link, cudaStreamNonBlocking, hi));
        FDR_CUDA(cudaStreamCreateWithPriority(&s->st_bar, cudaStreamNonBlocking, hi));
        s->ev.resize((size_t)7 * U + 2);
        const bool timeline = getenv("FDR_SHARD_TIMELINE") && atoi(getenv("FDR_SHARD_TIMELINE")) != 0;   // timing events: fdr_shard_timeline
        for (auto& e : s->ev) FDR_CUDA(cudaEventCreateWithFlags(&e, timeline ? cudaEventDefault : cudaEventDisableTiming));
    }
    cudaStream_t caller = pick(s, stream);
    auto E = [&](int kind, int u) { return s->ev[(size_t)kind * U + u]; };   // kinds 0..6
    cudaEvent_t ev_fork = s->ev[(size_t)7 * U], ev_join = s->ev[(size_t)7 * U + 1];
    FDR_CUDA(cudaEventRecord(ev_fork, caller));
    FDR_CUDA(cudaStreamWaitEvent(s->st_cmp, ev_fork, 0));
    FDR_CUDA(cudaStreamWaitEvent(s->st_link, ev_fork, 0));
    FDR_CUDA(cudaStreamWaitEvent(s->st_bar, ev_fork, 0));
    const bool one = s->world == 1;
    for (int u = 0; u < U; ++u) {   // rows forward
        FDR_TRY(fdr_shard_phase1_pairs(s, d_in_rows_u8, u, 1, s->st_cmp));
        FDR_CUDA(cudaEventRecord(E(0, u), s->st_cmp));
    }
    if (!one) {
        for (int u = 0; u < U; ++u) {   // exchange 1 + barrier
            if (s->staged) {
                FDR_CUDA(cudaStreamWaitEvent(s->st_link, E(0, u), 0));
                FDR_TRY(fdr_shard_exchange1(s, u, 1, s->st_link));
                FDR_CUDA(cudaEventRecord(E(1, u), s->st_link));
                FDR_CUDA(cudaStreamWaitEvent(s->st_bar, E(1, u), 0));
            } else {
                FDR_CUDA(cudaStreamWaitEvent(s->st_bar, E(0, u), 0));
            }
            FDR_TRY(fdr_shard_barrier(s, 2 * u, s->st_bar));
            FDR_CUDA(cudaEventRecord(E(2, u), s->st_bar));
        }
    }
    for (int u = 0; u < U; ++u) {   // columns
        if (!one) FDR_CUDA(cudaStreamWaitEvent(s->st_cmp, E(2, u), 0));
        FDR_TRY(fdr_shard_phase2_pairs(s, u, 1, s->st_cmp));
        FDR_CUDA(cudaEventRecord(E(3, u), s->st_cmp));
    }
    if (!one) {
        for (int u = 0; u < U; ++u) {   // exchange 3 + barrier
            if (s->staged) {
                FDR_CUDA(cudaStreamWaitEvent(s->st_link, E(3, u), 0));
                FDR_TRY(fdr_shard_exchange3(s, u, 1, s->st_link));
                FDR_CUDA(cudaEventRecord(E(4, u), s->st_link));
                FDR_CUDA(cudaStreamWaitEvent(s->st_bar, E(4, u), 0));
            } else {
                FDR_CUDA(cudaStreamWaitEvent(s->st_bar, E(3, u), 0));
            }
            FDR_TRY(fdr_shard_barrier(s, 2 * u + 1, s->st_bar));
            FDR_CUDA(cudaEventRecord(E(5, u), s->st_bar));
        }
    }
    for (int u = 0; u < U; ++u) {   // rows inverse
        if (!one) FDR_CUDA(cudaStreamWaitEvent(s->st_cmp, E(5, u), 0));
        FDR_TRY(fdr_shard_phase3_pairs(s, u, 1, s->st_cmp));
        FDR_CUDA(cudaEventRecord(E(6, u), s->st_cmp));
    }
    if (!one) FDR_TRY(fdr_shard_minmax_allreduce(s, s->st_cmp));
    FDR_TRY(fdr_shard_phase4_pack(s, d_out_rows_u8, s->st_cmp));
    FDR_CUDA(cudaEventRecord(ev_join, s->st_cmp));
    FDR_CUDA(cudaStreamWaitEvent(caller, ev_join, 0));
    return FDR_OK;
}

// Diagnostics (FDR_SHARD_TIMELINE=1 at the first fdr_shard_restore_rows): milliseconds from the start of the last restore to
// the END of every step, ms[kind * units + u] with kind 0 phase 1, 1 exchange 1, 2 barrier, 3 phase 2, 4 exchange 3, 5 barrier,
// 6 phase 3, and ms[7 * units] = the end of phase 4.  Synchronises the device.  Entries of skipped steps are negative.
FDR_API int fdr_shard_timeline(fdr_shard* s, float* ms, int capacity, int* count) {
    if (!s || !ms || !count) return set_error(FDR_E_INVALID, "bad arguments");
    if (s->ev.empty()) return set_error(FDR_E_STATE, "no restore has run");
    FDR_CUDA(cudaSetDevice(s->device));
    FDR_CUDA(cudaDeviceSynchronize());
    const int U = s->units, n = 7 * U + 1;
    if (capacity < n) return set_error(FDR_E_INVALID, "need room for %d values", n);
    cudaEvent_t t0 = s->ev[(size_t)7 * U];
    for (int i = 0; i < n; ++i) {
        cudaEvent_t e = (i < 7 * U) ? s->ev[(size_t)i] : s->ev[(size_t)7 * U + 1];
        float v = -1.f;
        if (cudaEventElapsedTime(&v, t0, e) != cudaSuccess) {
            v = -1.f;
            cudaGetLastError();
        }
        ms[i] = v;
    }
    *count = n;
    return FDR_OK;
}

// Cross-rank barrier on `stream` through flags in peer memory: every rank must call it with the same `set` in the same order.
// Work queued on the stream after it starts only when every rank's work queued before its own call has completed and is
// visible.  Different sets are independent sequences (one per concurrently running pipeline unit and phase).
FDR_API int fdr_shard_barrier(fdr_shard* s, int set, void* stream) {
    if (!s || set < 0 || set >= SYNC_SETS) return set_error(FDR_E_INVALID, "bad barrier arguments (sets 0..%d)", SYNC_SETS - 1);
    if (!s->have_peers) return set_error(FDR_E_STATE, "barrier before fdr_shard_set_peers");
    FDR_CUDA(cudaSetDevice(s->device));
    const unsigned int ep = ++s->epoch[set];
    peer_barrier_kernel<<<1, 32, 0, pick(s, stream)>>>(sync_peers(s), s->rank, s->world, set, ep, sync_status(s));
    FDR_CUDA(cudaGetLastError());
    s->launches += 1;
    return FDR_OK;
}

// Global extrema of every padded plane: all-reduce of the [channels][2] vector of fdr_shard_minmax_device over peer
// memory (one launch, also a barrier: afterwards every rank has finished phase 3, so the slabs are free for the next image).
FDR_API int fdr_shard_minmax_allreduce(fdr_shard* s, void* stream) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    if (!s->have_peers) return set_error(FDR_E_STATE, "all-reduce before fdr_shard_set_peers");
    FDR_CUDA(cudaSetDevice(s->device));
    const unsigned int ep = ++s->mm_epoch;
    peer_minmax_kernel<<<1, 128, 0, pick(s, stream)>>>(sync_peers(s), s->rank, s->world, SYNC_SETS - 1, ep, sync_status(s), s->mmf.p,
                                                        2 * s->C, s->minmax_neg);
    FDR_CUDA(cudaGetLastError());
    s->launches += 1;
    return FDR_OK;
}

// 0 = every barrier so far completed; 1 = one gave up after its time-out (a peer never arrived).  Synchronises `stream`.
FDR_API int fdr_shard_sync_status(fdr_shard* s, void* stream, int* timed_out) {
    if (!s || !timed_out) return set_error(FDR_E_INVALID, "bad arguments");
    FDR_CUDA(cudaSetDevice(s->device));
    FDR_CUDA(cudaStreamSynchronize(pick(s, stream)));
    unsigned int v = 0;
    FDR_CUDA(cudaMemcpy(&v, sync_status(s), sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out = v != 0;
    return FDR_OK;
}

FDR_API int fdr_shard_pair_count(const fdr_shard* s, int* pairs) {
    if (!s || !pairs) return set_error(FDR_E_INVALID, "bad arguments");
    *pairs = s->units;
    return FDR_OK;
}

// Exchange passes (phase 1 scatter, phase 3 gather) as at most `ctas` persistent CTAs per unit (0 = the whole grid), so
// that another unit's column phase finds free SMs while this one waits on NVLink.  Default: FDR_SHARD_ROW_CTAS or 0.
FDR_API int fdr_shard_set_row_ctas(fdr_shard* s, int ctas) {
    if (!s || ctas < 0) return set_error(FDR_E_INVALID, "bad arguments");
    s->row_ctas = ctas;
    return FDR_OK;
}

// enabled: the device min/max vector (fdr_shard_minmax_device) holds (min, -max) per plane, so ONE all-reduce(MIN) over the
// whole [channels][2] vector yields the global extrema; phase 4 undoes the sign.
FDR_API int fdr_shard_set_minmax_negated(fdr_shard* s, int enabled) {
    if (!s) return set_error(FDR_E_INVALID, "shard is NULL");
    s->minmax_neg = enabled != 0;
    return FDR_OK;
}

FDR_API int fdr_shard_half_plane(const fdr_shard* s, int* enabled) {
    if (!s || !enabled) return set_error(FDR_E_INVALID, "bad arguments");
    *enabled = s->half ? 1 : 0;
    return FDR_OK;
}

FDR_API int fdr_shard_last_launch_count(const fdr_shard* s, long long* launches) {
    if (!s || !launches) return set_error(FDR_E_INVALID, "bad arguments");
    *launches = s->launches;
    return FDR_OK;
}

// Rows [first_row, first_row + n_rows) of synthetic image `image` (counter hash of the whole image).
FDR_API int fdr_synth_rows_device_u8(void* d_out, uint32_t seed, long long image, int channels, int rows_total, int cols,
                                     int first_row, int n_rows, void* stream) {
    if (!d_out || channels < 1 || rows_total < 1 || cols < 1 || first_row < 0 || n_rows < 0 || first_row + n_rows > rows_total)
        return set_error(FDR_E_INVALID, "bad arguments");
    if (n_rows == 0) return FDR_OK;
    FDR_CUDA(launch_synth_u8(static_cast<uint8_t*>(d_out), seed, image, 1, channels, (long long)rows_total * cols,
                             (long long)first_row * cols, (long long)n_rows * cols, static_cast<cudaStream_t>(stream)));
    return FDR_OK;
}
