// passes.h -- host-visible argument blocks and launchers of the fused FFT passes.
//
// The restoration of one pair of real planes (a, b) packed as z = a + i*b is
//   pass 1  rows  : load a,b (f32 or u8), zero-pad, FFT along x           -> spectrum S1
//   pass 2  cols  : FFT along y, multiply by the Wiener factor
//                   Wf = conj(H)/(|H|^2+K), inverse FFT along y            (in place)
//   pass 3  rows  : inverse FFT along x, split Re -> plane a, Im -> plane b,
//                   min/max of every padded plane                          -> raw planes
//   pass 4        : min-max normalise + crop + 8-bit pack (or f32 planes)
// which replaces the reference's 12 launches per channel
// (/root/reference/fft/fft_gpu.cu:338-368).  Packing two real planes into one complex
// transform is valid because Wf is the spectrum of a real kernel (Hermitian), so
// IFFT(Wf * FFT(a + i b)) = a' + i b'.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdr {

constexpr int FDR_MINMAX_SLOTS = 128;  // atomic slots per plane for the min/max of pass 3

// ROW_IN_GATHER / ROW_OUT_SCATTER are the row-sharded (multi-GPU) forms: the row lives on this
// GPU, its columns are spread over `world` column slabs [rows_padded][n/world], one per GPU, reached
// through peer-mapped pointers (NVLink).  The transpose of the reference's MPI_Alltoallv
// (/root/reference/fft/fft_mpi.cpp:170-279) is thereby fused into the row passes.
//
// HALF-PLANE forms (one real plane per transform instead of a pair of planes): rows y and y + D of the SAME plane are
// packed as z = row_y + i*row_{y+D}; the row spectra of real rows are Hermitian, so after untangling
//   X_y[k] = (Z[k] + conj Z[N-k]) / 2,   X_{y+D}[k] = (Z[k] - conj Z[N-k]) / (2i)
// only columns 0 .. N/2-1 are stored (a half-width plane [rows][N/2]) plus the real Nyquist column k = N/2 in a side
// vector.  The column pass then runs on half the columns, and the inverse row pass rebuilds Z[k] and Z[N-k] from the
// half plane (ROW_IN_HALF).  An odd colour plane (the R of BGR) therefore costs half of a pair instead of a whole pair:
// -25 % of passes 1-3 for one RGB image and -25 % of the NVLink volume of the row-sharded path.  The reference has no
// counterpart (it transforms every channel as a full complex plane, /root/reference/fft/fft_gpu.cu:325-368).
enum RowInMode { ROW_IN_PAIR_F32 = 0, ROW_IN_PAIR_U8 = 1, ROW_IN_COMPLEX = 2, ROW_IN_GATHER = 3,
                 ROW_IN_ROWS2_F32 = 4, ROW_IN_ROWS2_U8 = 5, ROW_IN_HALF = 6 };
enum RowOutMode { ROW_OUT_COMPLEX = 0, ROW_OUT_REAL_PAIR = 1, ROW_OUT_SCATTER = 2, ROW_OUT_HALF = 3, ROW_OUT_REAL_ROWS2 = 4 };
constexpr int FDR_MAX_PEERS = 16;
constexpr int FDR_HALF_MIN_N = 64;  // shortest row length the half-plane forms are instantiated for
enum ColMode { COL_FFT = 0, COL_WIENER = 1, COL_MAKE_WIENER = 2, COL_FILTER = 3, COL_COPY = 4 /* timing probe: load + store only */ };

struct RowPassArgs {
    int n;               // transform length (padded columns), power of two
    int nrows;           // rows transformed per pair (grid.x covers them)
    int npairs;          // grid.y
    int pair_base;       // first pair of this launch (pairs pair_base .. pair_base+npairs-1 of the workspace)
    int in_mode, out_mode;
    int conj;            // complex->complex only: conjugate on load and on store (inverse = conj FFT conj)
    const float2* tw;    // twiddle table of length n (get_twiddles)
    // ---- real-pair input (pass 1) ----
    const float* in_f32;         // planar f32: unit u at in_f32 + u*in_unit_stride, row stride in_row_stride
    long long in_unit_stride;
    long long in_row_stride;
    const uint8_t* in_u8;        // interleaved u8 [img][H][W][C]; unit u = img*C + c
    int channels;                // C
    int img_rows, img_cols;      // H, W of the unpadded image: x >= W reads 0
    long long unit_base;         // global index of local unit 0
    long long units_total;       // global unit count (a pair's second unit may not exist)
    // ---- complex input / output ----
    const float2* cin;           // pair p at cin + p*cplane, row r at + r*n
    float2* cout;
    long long cplane;
    // ---- real-pair output (pass 3) ----
    float* raw;                  // local unit u at raw + u*raw_unit_stride, cropped rows x cols
    long long raw_unit_stride;
    int raw_rows, raw_cols;      // H, W : only y < H, x < W is stored
    unsigned int* minmax;        // [local unit][FDR_MINMAX_SLOTS][2] ordered-uint encoded min, max over the PADDED plane;
                                 // CTAs spread their atomics over the slots (same-address atomics from every CTA of a
                                 // plane cost pass 3 up to 70 %, profiles/ubench), the finalize/decode kernels fold them
    int local_units;             // units in this chunk
    // ---- row-sharded exchange (ROW_IN_GATHER / ROW_OUT_SCATTER) ----
    float2* const* peers;        // device array [world]: base of every GPU's column slab (NULL = skip that peer)
    int peer_shift;              // log2(n / world): column x belongs to peer x >> peer_shift
    long long peer_plane;        // elements per pair in a slab = rows_padded * (n / world)
    int row0;                    // global (padded) index of local row 0
    int max_ctas;                // > 0: scatter / gather passes run as at most this many persistent CTAs per pair
    unsigned int* mm_reset;      // != NULL: CTA 0 of every pair / plane re-arms the min/max slots of its planes (same layout as `minmax`;
                                 // saves the separate reset launch; needs local_units)
    int prefetch_dist;           // long rows: L2 prefetch of the input of the row block this many blocks ahead (set by the launcher)
    // ---- half-plane forms (ROW_IN_ROWS2_*, ROW_OUT_HALF, ROW_IN_HALF, ROW_OUT_REAL_ROWS2): blockIdx.y counts PLANES ----
    int pair_dist;               // D: row `r` of the launch carries rows r and r + D (local indices, like `row0 + r` globally)
    int rows_in;                 // ROWS2 input: local rows present; the second row reads as zero when r + D >= rows_in
    int hp_rows_store;           // ROW_OUT_HALF: global rows >= this are not stored (the image height; the column pass zero-fills)
    float2* hp_peers[FDR_MAX_PEERS];  // half planes by column owner: column k < n/2 of plane u, global row g at
                                      // hp_peers[k >> hp_shift] + u*hp_plane + (g << hp_shift) + (k & mask); one entry when unsharded
    int hp_shift;                // log2(columns of the half plane per owner)
    int hp_local;                // every hp_peers entry is memory of this GPU (L2 prefetch allowed)
    long long hp_plane;          // elements per plane in a half-plane slab = rows_padded << hp_shift
    float2* nyq_peers[FDR_MAX_PEERS]; // Nyquist columns: plane u, global row g at nyq_peers[u % nyq_world] + u*nyq_plane + g
    int nyq_world;
    long long nyq_plane;         // = rows_padded
};

struct ColPassArgs {
    int n;                // transform length (padded rows), power of two
    int pitch;            // elements per row (padded columns)
    int npairs;           // grid.y
    int pair_base;        // first pair of this launch
    int mode;             // ColMode
    int conj;             // COL_FFT only: inverse transform
    const float2* tw;     // twiddle table of length n
    int rows_valid;       // rows >= rows_valid are read as zero (pass 1 skipped them)
    float2* data;         // in place; pair p at data + p*cplane
    long long cplane;
    const float2* wiener; // COL_WIENER: Wf, row-major n x pitch
    const float2* wiener_tiled;  // optional tile-major copy [pitch/cw][n][cw] (launch_wiener_retile): each column tile of
                                 // the wide kernel is then ONE contiguous 64 KB block instead of n short rows
    int wiener_blocks;    // wide kernel: > 1 = the factor is a stack of that many n-row blocks and pair p uses block
                          // (pair_base + p) % wiener_blocks (the K x 2048 long-column scheme, col_blocks.cu)
    float2* wiener_out;   // COL_MAKE_WIENER
    float K;
    int col_variant;      // COL_WIENER kernel choice (timing probe): 0 = default dispatch, 1 = plain-load kernel,
                          // 2 = TMA kernel, one tile per CTA, 3 = TMA kernel, persistent + pipelined (col_tma.cu),
                          // 4 = TMA kernel on the 64-points-per-thread core (col_wide.cu), 5 / 6 = its transfer-only probes,
                          // 7 = wide core, persistent + pipelined
};

// Opt-in to more than 48 KB of dynamic shared memory for `func` on the CURRENT device; done once per (function, device),
// thread-safe (several plans or shards on different devices may launch from different host threads).
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);
// SM count of the current device (cached per device).
int device_sm_count();
// Launchers (defined in passes_*.cu).  Return cudaGetLastError() of the launch.
cudaError_t launch_row_pass(const RowPassArgs& a, cudaStream_t s);
cudaError_t launch_col_pass(const ColPassArgs& a, cudaStream_t s);
// Twiddle table for power-of-two length n on the current device (cached).
cudaError_t get_twiddles(int n, const float2** out);
// table of the 32-points-per-thread long-row core (fft_mid.cuh), n = 8192 or 16384
cudaError_t get_twiddles_mid(int n, const float2** out);
// Long columns (n = 8192, 16384), COL_WIENER / COL_MAKE_WIENER: split column pass -- K x 2048 blocks (col_blocks.cu) when the
// wide TMA kernel applies, else 128 x 128 four-step (col_split.cuh).  The Wiener factor it reads/writes is in digit-swapped row
// order: row M*k1 + k2 holds frequency k1 + (n/M)*k2 with M = col_split_block_len(args) (2048 or 128).
bool col_split_applicable(const ColPassArgs& a);
cudaError_t launch_col_split(const ColPassArgs& a, cudaStream_t s, int* launches);
int col_split_block_len(const ColPassArgs& a);
bool col_blocks_applicable(const ColPassArgs& a);
cudaError_t launch_col_blocks(const ColPassArgs& a, cudaStream_t s, int* launches);
// COL_WIENER with TMA-staged tiles (col_tma.cu), 256 <= n <= 4096, row-major planes.
bool col_tma_applicable(const ColPassArgs& a);
bool col_tma_geometry_ok(int n, int pitch, long long cplane);
cudaError_t launch_col_wiener_tma(const ColPassArgs& a, cudaStream_t s);
// COL_WIENER on the 64-points-per-thread core (col_wide.cu), n = 2048; same preconditions as the TMA kernel.
bool col_wide_applicable(const ColPassArgs& a);
cudaError_t launch_col_wiener_wide(const ColPassArgs& a, cudaStream_t s);
// timing probe: copy the workspace through 64 KB shared-memory tiles of box_cols columns with TMA
cudaError_t launch_tma_copy_probe(const ColPassArgs& a, int box_cols, cudaStream_t s);
// columns per tile of the wide kernel for length n (64 KB tiles): 8 at 1024, 4 at 2048, 2 at 4096; 0 = length not served
constexpr int wide_tile_cols(int n) { return n == 1024 ? 8 : n == 2048 ? 4 : n == 4096 ? 2 : 0; }
// tile-major copy of the Wiener factor: dst[(xt*n + row)*cw + c] = src[row*pitch + xt*cw + c], cw = wide_tile_cols(n)
cudaError_t launch_wiener_retile(const float2* src, float2* dst, int n, int pitch, cudaStream_t s);
// Tile width (columns per CTA) the column pass uses for length n.
int col_pass_tile_width(int n);

// ---- small kernels (kernels_misc.cu) ----
struct PsfAffine {  // inverse rotation, computed on the host in double (utils.hpp:20 + warpAffine)
    double a00, a01, b0, a10, a11, b1;
};
cudaError_t launch_motion_psf(float* psf, int size, PsfAffine m, cudaStream_t s);
cudaError_t launch_minmax_reset(unsigned int* minmax, int units, cudaStream_t s);
// scale_shift[u] = {scale, shift} as floats from the double-precision min/max rule
cudaError_t launch_minmax_finalize(const unsigned int* minmax, float2* scale_shift, float* minmax_f32, int units,
                                   cudaStream_t s);
// u8 interleaved [img][H][W][C] from raw planes of local units img*C + c
cudaError_t launch_pack_u8(const float* raw, long long raw_unit_stride, const float2* scale_shift, uint8_t* out,
                           int imgs, int channels, int rows, int cols, cudaStream_t s);
// Lab white balance + 8-bit pack for 3-channel images (gpu.cpp:123-134); orig_u8 or orig_f32 is the blurred input
cudaError_t launch_white_balance_pack_u8(const float* raw, long long raw_unit_stride, const float2* scale_shift,
                                         const uint8_t* orig_u8, const float* orig_f32, long long orig_unit_stride, double* sums,
                                         uint8_t* out, int imgs, int rows, int cols, cudaStream_t s);
// normalised f32 planes [unit][H][W]
cudaError_t launch_normalize_f32(const float* raw, long long raw_unit_stride, const float2* scale_shift, float* out,
                                 long long out_unit_stride, int units, int rows, int cols, cudaStream_t s);
cudaError_t launch_synth_u8(uint8_t* out, uint32_t seed, long long img0, int imgs, int channels, long long plane_px,
                            long long px0, long long npx, cudaStream_t s);
// negate_max / negated_max: the f32 vector holds (min, -max) per plane (one all-reduce(MIN) across ranks folds both)
// pack of a few BGR images with the slot fold + scale/shift (launch_minmax_finalize) done by every CTA itself: one launch less
cudaError_t launch_pack_u8_c3_fused(const float* raw, long long raw_unit_stride, const unsigned int* minmax, float2* scale_shift, float* minmax_f32,
                                    uint8_t* out, int imgs, int rows, int cols, cudaStream_t s);
bool pack_u8_c3_fused_applicable(const float* raw, long long raw_unit_stride, const uint8_t* out, int imgs, int channels, int rows, int cols);
cudaError_t launch_minmax_decode(const unsigned int* minmax, float* minmax_f32, int units, int negate_max, cudaStream_t s);
cudaError_t launch_scale_shift_from_f32(const float* minmax_f32, float2* scale_shift, int units, int negated_max, cudaStream_t s);
// out[y] = column n/2 of the row spectrum of PSF row y (real): sum_x psf[y][x] * (-1)^x
cudaError_t launch_psf_nyquist(const float* psf, int rows, int cols, float2* out, cudaStream_t s);
// out-of-place batched O(n^2) DFT: element i of batch b at in[b*batch_stride + i*elem_stride]
cudaError_t launch_dft_naive(const float2* in, float2* out, int n, long long elem_stride, int batch,
                             long long batch_stride, int inverse, cudaStream_t s);
cudaError_t launch_l2_flush(void* buf, size_t bytes, cudaStream_t s);

}  // namespace fdr
