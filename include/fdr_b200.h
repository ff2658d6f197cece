/* fdr_b200.h -- C ABI of the B200-native frequency-domain restoration path.
 *
 * This is the drop-in boundary for the reference's gpu mode.  Each entry point names the
 * reference interface it replaces (file:line under the reference repository); the C++
 * functions of fft/fft.hpp (namespace fft_gpu) and the ./gpu CLI in this repository are thin
 * wrappers over these calls, and INTEGRATION.md shows the binding a maintainer of the
 * reference would add.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 (FDR_OK) or a negative FDR_E_* code; fdr_last_error() gives the
 *     text of the calling thread's most recent failure (the C++ wrappers turn a non-zero status
 *     into the reference's "Error: <file>:<line>, <msg>" + exit(1), fft/fft_gpu.cu:59-66);
 *   - pointers are borrowed for the duration of the call; the plan owns all device memory;
 *   - images are H x W, padded internally to powers of two (utils.hpp:27-31,
 *     fft_gpu.cu:287-288); planes are fp32 in [0,1]; complex data is interleaved (re, im) fp32;
 *   - results follow the reference's SERIAL path (fft_serial.cpp:141-261 + serial.cpp:33-39):
 *     zero-pad, FFT, G*conj(H)/(|H|^2+K), unscaled inverse FFT, real part, min-max normalise
 *     over the PADDED plane, crop, and for the 8-bit entry points rint(255*x) saturated;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     FDR_E_CUDA;
 *   - streams: a plan owns ONE workspace.  Device entry points are asynchronous on the stream the
 *     caller passes; calls on different streams (and fdr_plan_set_psf_*, which rebuilds the Wiener
 *     factor on the plan's own stream) are ordered one after the other through an event, so they are
 *     safe but never concurrent.  Use one plan per stream for concurrency.  A NULL stream means the
 *     plan's own non-blocking stream, not the legacy default stream.
 */
#ifndef FDR_B200_H
#define FDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDR_OK 0
#define FDR_E_INVALID (-1) /* bad argument */
#define FDR_E_CUDA (-2)    /* CUDA runtime error, see fdr_last_error() */
#define FDR_E_STATE (-3)   /* call order (e.g. restore before a PSF was set) */
#define FDR_E_NOMEM (-4)

typedef struct fdr_plan fdr_plan;

/* ---- library ------------------------------------------------------------------------- */
const char* fdr_last_error(void);
int fdr_version(void);
int fdr_device_count(int* count);
/* Pinned host memory for the host entry points (cudaMallocHost, fft_gpu.cu:304-306). */
int fdr_host_alloc(void** ptr, size_t bytes);
int fdr_host_free(void* ptr);

/* ---- plan: sizes, workspace, Wiener factor -------------------------------------------- */
/* rows x cols: unpadded image size; channels: planes per image (3 for BGR); max_images: upper
 * bound of images per restore call (workspace is chunked, so this only bounds staging);
 * device: CUDA ordinal.  Replaces the per-call cudaMalloc block of
 * fft_gpu::wienerDeblur_RGB_optimized (fft_gpu.cu:304-322). */
int fdr_plan_create(fdr_plan** plan, int rows, int cols, int channels, int max_images, int device);
int fdr_plan_destroy(fdr_plan* plan);
/* Padded sizes chosen by the plan (nextPowerOfTwo, utils.hpp:27-31). */
int fdr_plan_padded_size(const fdr_plan* plan, int* padded_rows, int* padded_cols);
/* 8-bit image outputs pass through the reference drivers' post stage (gpu.cpp:123-134 = serial.cpp:43-54):
 * BGR -> Lab, L scaled by mean(L_original)/(mean(L_restored)+1e-6) and clamped (applyWhiteBalance,
 * utils.hpp:55-71), Lab -> BGR, convertTo(CV_8U, 255).  Off by default (direct x255 pack). */
int fdr_plan_set_white_balance(fdr_plan* plan, int enabled);
/* enabled = 1 when a chunk with an odd plane count (one BGR image, a greyscale image) sends its last plane through the
 * half-plane path: two rows of that plane per complex row transform, Hermitian half spectrum, half the column work --
 * instead of a half-empty complex pair (the reference transforms every channel as a full complex plane,
 * fft_gpu.cu:325-368).  Geometry-dependent; FDR_HALF=0 / FDR_HALF_MIN_PIXELS=<padded pixels> override it. */
int fdr_plan_half_plane(const fdr_plan* plan, int* enabled);
/* The post stage alone, for callers that hold normalised planes (what fft_gpu::wienerDeblur_RGB_* returns):
 * restored and original planes in B, G, R order, fp32 in [0,1], contiguous rows x cols -> 8-bit BGR image. */
int fdr_white_balance_pack_host(const float* const* restored_planes, const float* const* original_planes, int rows, int cols,
                                uint8_t* out_bgr);
/* Images per internal chunk (0 = automatic).  Tuning knob; results do not depend on it. */
int fdr_plan_set_chunk_images(fdr_plan* plan, int images);

/* PSF given by the caller, as motionBlurKernel() returns it (utils.hpp:15-24): psf_rows x
 * psf_cols fp32, anchored top-left when padded (fft_serial.cpp:166-170).  Builds
 * Wf = conj(H)/(|H|^2 + K) on the device once (the reference recomputes the PSF spectrum per
 * channel, fft_gpu.cu:329-356). */
int fdr_plan_set_psf_host(fdr_plan* plan, const float* psf, int psf_rows, int psf_cols, size_t psf_stride_bytes,
                          float K);
/* PSF built ON THE DEVICE from (length, angle) -- utils.hpp:15-24 motionBlurKernel with
 * OpenCV's fixed-point bilinear warpAffine, bit-identical to cv2 4.13. */
int fdr_plan_set_psf_motion(fdr_plan* plan, int length, double angle_deg, float K);
/* Copies of plan state for inspection / parity tests. */
int fdr_plan_get_psf_host(const fdr_plan* plan, float* psf_out, int capacity_floats, int* psf_rows, int* psf_cols);
int fdr_plan_get_wiener_host(const fdr_plan* plan, float* wf_interleaved /* padded_rows*padded_cols*2 */);

/* ---- restoration: host buffers (the reference-facing calls) --------------------------- */
/* fft_gpu::wienerDeblur_RGB_optimized / _naive (fft/fft.hpp:32-33, fft_gpu.cu:279-394,
 * 400-512): n_planes fp32 planes of rows x cols (row stride in bytes) in, the same number of
 * min-max normalised fp32 planes out.  in and out may alias. */
int fdr_restore_planes_host_f32(fdr_plan* plan, const float* const* in_planes, size_t in_stride_bytes,
                                float* const* out_planes, size_t out_stride_bytes, int n_planes);
/* Whole images: interleaved 8-bit [image][row][col][channel] in (as cv::imread gives,
 * gpu.cpp:68), restored 8-bit images out (gpu.cpp:134 without the Lab stage).  The /255
 * conversion (gpu.cpp:70-71) happens on the device. */
int fdr_restore_images_host_u8(fdr_plan* plan, const uint8_t* in_images, uint8_t* out_images, int n_images);

/* ---- restoration: device-resident buffers (benchmarks, pipelines) --------------------- */
/* stream: a cudaStream_t (NULL = the plan's own stream).  Asynchronous with respect to the
 * host; inputs and outputs are device pointers on the plan's device. */
int fdr_restore_images_device_u8(fdr_plan* plan, const void* d_in_images, void* d_out_images, int n_images,
                                 void* stream);
/* d_in_planes: [n_planes][rows][cols] fp32 contiguous.  Either output may be NULL:
 * d_out_planes_f32 [n_planes][rows][cols] normalised fp32; d_out_images_u8 interleaved 8-bit
 * (n_planes must then be a multiple of the plan's channel count). */
int fdr_restore_planes_device_f32(fdr_plan* plan, const void* d_in_planes, void* d_out_planes_f32,
                                  void* d_out_images_u8, int n_planes, void* stream);
/* min and max of each UN-normalised padded plane of the most recent restore chunk
 * ([plane][2] floats, at most `capacity_planes` planes). */
int fdr_plan_last_minmax_host(fdr_plan* plan, float* minmax, int capacity_planes);
/* Six buckets in ms of the most recent HOST restore call, in the order of the reference's
 * Profiler (fft_gpu.cu:17-57): alloc, H2D, pre-process, compute, D2H, post-process. */
int fdr_plan_get_profile(const fdr_plan* plan, float ms[6]);
/* Per-kernel device timing.  When enabled every pass launch is bracketed by a CUDA event pair
 * on the launching stream.  get_kernel_timing drains the records accumulated since the last
 * call: for kind 0 = pass 1 (rows forward), 1 = pass 2 (columns + Wiener), 2 = pass 3 (rows
 * inverse + min/max), 3 = pass 4 (normalise/pack) it returns the summed duration in ms, the
 * launch count and the summed algorithmic bytes (this formulation's own count). */
int fdr_plan_set_kernel_timing(fdr_plan* plan, int enabled);
int fdr_plan_get_kernel_timing(fdr_plan* plan, double total_ms[4], long long launches[4], double bytes[4]);
/* Timing probe for kernel work: runs one pass `reps` times back to back on a synthetic workspace
 * of `npairs` plane pairs and returns the mean device time in ms.  pass 1 = rows forward,
 * 2 = columns, 3 = rows inverse + min/max.  Column variants: 0 = Wiener, default dispatch; 1 = Wiener with plain
 * loads; 2 = single forward FFT; 3 = load+store only; 4 / 5 = 16-point TMA kernel, one tile per CTA / persistent
 * pipelined; 6 = 64-point (wide) TMA kernel; 7 / 8 = its transfer-only probes (2 / 3 tile transfers, no FFT);
 * 9 = wide kernel, persistent pipelined; 10 = wide kernel at 4096 rows with 4-column tiles; 100 + w = TMA copy of
 * 64 KB tiles w columns wide, in and out (what the box width costs).  Probes leave garbage in the workspace. */
int fdr_plan_time_pass(fdr_plan* plan, int pass, int variant, int npairs, int reps, float* ms_avg);
/* Number of kernels the most recent restore call launched. */
int fdr_plan_last_launch_count(const fdr_plan* plan, long long* launches);

/* ---- spectra for the parity gates (1e-4 relative L2 against fft_serial) --------------- */
/* G = FFT2(zero-padded plane) (fft_serial.cpp:176), padded_rows*padded_cols*2 floats. */
int fdr_plan_forward_spectrum_host(fdr_plan* plan, const float* plane, size_t stride_bytes, float* G_interleaved);
/* F = G * conj(H)/(|H|^2+K) (fft_serial.cpp:186-224). */
int fdr_plan_filtered_spectrum_host(fdr_plan* plan, const float* plane, size_t stride_bytes, float* F_interleaved);

/* ---- building blocks named after fft/fft.hpp's fft_gpu declarations ------------------- */
/* fft_gpu::my_dft2D(Mat&, bool) (fft.hpp:40; empty stub in fft_gpu.cu:515): in-place 2-D DFT of
 * a rows x cols interleaved complex matrix, unscaled in both directions
 * (fft_serial.cpp:113-139). */
int fdr_dft2d_host(float* interleaved, int rows, int cols, int inverse);
/* fft_gpu::fft_radix2_kernel(float*, int, bool) (fft.hpp:35; stub fft_gpu.cu:514): in-place
 * power-of-two FFT of n interleaved complex points (fft_serial.cpp:40-68). */
int fdr_fft_radix2_host(float* interleaved, int n, int inverse);
/* fft_gpu::dft_naive_kernel (fft.hpp:37; never defined by the reference): O(n^2) DFT of any
 * length (fft_serial.cpp:71-87). */
int fdr_dft_naive_host(float* interleaved, int n, int inverse);
/* fft_gpu::transform_row_kernel (fft.hpp:39; never defined): radix-2 when n is a power of
 * two, naive DFT otherwise (fft_serial.cpp:90-108).  `rows` independent rows. */
int fdr_transform_rows_host(float* interleaved, int rows, int n, int inverse);

/* ---- one large image row-sharded over several GPUs ------------------------------------- */
/* Replaces the reference's MPI mode for this path (fft/fft_mpi.cpp:311-470: Bcast dims,
 * Scatterv rows, [row FFT, MPI_Alltoallv transpose, column FFT, Alltoallv back] x3, Gatherv).
 * One fdr_shard per GPU (one process per GPU, or several shards in one process).  Rank g owns
 * padded rows [g*Rp/world, (g+1)*Rp/world) and, in the column pass, padded columns
 * [g*Cp/world, (g+1)*Cp/world).  The transposes are fused into the row passes as peer (NVLink)
 * stores and loads; the caller supplies the cross-rank barriers and the 2-float-per-plane
 * min/max all-reduce (torch.distributed / NCCL), see <package>/fdr_dist.py:
 *     phase1 | barrier | phase2 | barrier | phase3 | all-reduce(min,max) | phase4            */
typedef struct fdr_shard fdr_shard;
int fdr_shard_create(fdr_shard** shard, int rows, int cols, int channels, int rank, int world, int device);
int fdr_shard_destroy(fdr_shard* shard);
/* first_row/n_rows: the image rows this rank reads and writes (calculate_distribution,
 * fft_mpi.cpp:89-100, applied to the padded rows). */
int fdr_shard_geometry(const fdr_shard* shard, int* first_row, int* n_rows, int* padded_rows, int* padded_cols,
                       int* cols_per_rank);
/* This rank's column slab (device pointer) for export to the peers. */
int fdr_shard_local_slab(const fdr_shard* shard, void** d_slab, size_t* bytes);
/* CUDA IPC plumbing for one-process-per-GPU runs: 64-byte handles travel through any host
 * channel (torch.distributed all_gather_object, MPI_Allgather).  Tear-down order: every rank closes the
 * handles it opened (fdr_ipc_close), a cross-rank barrier, and only then fdr_shard_destroy -- CUDA leaves
 * freeing exported memory that a peer still has mapped undefined. */
int fdr_ipc_export(const void* dptr, unsigned char handle[64]);
int fdr_ipc_open(const unsigned char handle[64], void** dptr);
int fdr_ipc_close(void* dptr);
/* slabs[world]: every rank's slab as a pointer valid on THIS device (entry `rank` is ignored). */
int fdr_shard_set_peers(fdr_shard* shard, void* const* slabs);
/* Both build the rank's slab of the Wiener factor using the rank's COLUMN SLAB as scratch: fence across ranks (all ranks
 * returned from this call) before the first fdr_shard_phase1 of any rank, which stores into the peers' slabs. */
int fdr_shard_set_psf_motion(fdr_shard* shard, int length, double angle_deg, float K);
int fdr_shard_set_psf_host(fdr_shard* shard, const float* psf, int psf_rows, int psf_cols, float K);
/* d_in_rows_u8 / d_out_rows_u8: this rank's rows, interleaved 8-bit [n_rows][cols][channels]. */
int fdr_shard_phase1_rows(fdr_shard* shard, const void* d_in_rows_u8, void* stream);
int fdr_shard_phase2_cols(fdr_shard* shard, void* stream);
int fdr_shard_phase3_rows(fdr_shard* shard, void* stream);
/* The same phases restricted to plane pairs [pair_first, pair_first + pair_count) (a 3-channel image has
 * 2 pairs: B+iG and R+i0).  Lets the caller pipeline the pairs on separate streams so that one pair's
 * NVLink-bound row phase overlaps the other pair's HBM-bound column phase (fdr_dist.ShardedRestorer). */
int fdr_shard_phase1_pairs(fdr_shard* shard, const void* d_in_rows_u8, int pair_first, int pair_count, void* stream);
int fdr_shard_phase2_pairs(fdr_shard* shard, int pair_first, int pair_count, void* stream);
int fdr_shard_phase3_pairs(fdr_shard* shard, int pair_first, int pair_count, void* stream);
int fdr_shard_pair_count(const fdr_shard* shard, int* pairs);
/* Half-plane mode (default when the geometry allows; FDR_SHARD_HALF=0 disables): every colour plane is its own pipeline
 * unit -- local rows y and y+D of ONE plane are packed into a complex row transform, untangled, and only columns
 * 0 .. Cp/2-1 (+ the Nyquist column, kept on rank `plane % world`) travel and go through the column phase: 1.5 instead of
 * 2 complex planes for a BGR image.  In this mode the "pair" arguments above index planes and fdr_shard_pair_count
 * returns the channel count.  Replaces the three per-channel calls of mpi.cpp:95-111. */
int fdr_shard_half_plane(const fdr_shard* shard, int* enabled);
/* Exchange passes (phase 1, phase 3) as at most `ctas` persistent CTAs per unit (0 = whole grid; default
 * FDR_SHARD_ROW_CTAS), leaving SMs to another unit's column phase in the pipelined driver. */
int fdr_shard_set_row_ctas(fdr_shard* shard, int ctas);
/* enabled: the vector of fdr_shard_minmax_device holds (min, -max) per plane so that ONE all-reduce(MIN) over all of it
 * folds both extrema; phase 4 undoes the sign.  Default off (column 0 MIN, column 1 MAX). */
int fdr_shard_set_minmax_negated(fdr_shard* shard, int enabled);
/* STAGED mode (the default in half-plane mode on more than one rank; FDR_SHARD_STAGED=0 keeps the fused stores/loads): phase 1
 * and phase 3 work on LOCAL staging planes and these two calls move the column blocks between the ranks -- exchange 1 after
 * phase 1 (row spectra -> the column owners' slabs), exchange 3 after phase 2 (filtered columns -> the row owners' staging
 * planes) -- as plain stores over NVLink from a few persistent CTAs (fdr_shard_set_link_ctas, default 24), so a transfer
 * runs beside the passes of other units instead of holding every SM.  A barrier must follow each exchange.  No-ops when the
 * shard is not staged.  They are the MPI_Alltoallv calls of fft_mpi.cpp:284-307. */
int fdr_shard_exchange1(fdr_shard* shard, int unit_first, int unit_count, void* stream);
int fdr_shard_exchange3(fdr_shard* shard, int unit_first, int unit_count, void* stream);
int fdr_shard_staged(const fdr_shard* shard, int* enabled);
int fdr_shard_set_ce_streams(fdr_shard* shard, int streams);   /* copy-engine mode: blocks of an exchange spread over 1..8 streams */
int fdr_shard_set_link_ctas(fdr_shard* shard, int ctas);   /* 0: move the blocks with the copy engines instead (FDR_SHARD_LINK=ce) */
/* The whole restoration of this rank's rows in one call, pipelined over the units: passes on a compute stream, exchanges
 * and barriers on two high-priority streams, events in between; ordered after / before `stream`.  Collective: every rank
 * calls it.  Needs the peer-memory barriers below (no NCCL / MPI on the per-image path).  mpi.cpp:95-111. */
int fdr_shard_restore_rows(fdr_shard* shard, const void* d_in_rows_u8, void* d_out_rows_u8, void* stream);
/* Diagnostics (environment FDR_SHARD_TIMELINE=1 before the first fdr_shard_restore_rows): ms from the start of the last restore to
 * the end of every step, ms[kind * units + u], kinds 0 phase 1, 1 exchange 1, 2 barrier, 3 phase 2, 4 exchange 3, 5 barrier,
 * 6 phase 3; ms[7 * units] = end of phase 4.  Synchronises the device. */
int fdr_shard_timeline(fdr_shard* shard, float* ms, int capacity, int* count);
/* Cross-rank synchronisation through flags in peer memory (the slab allocation carries them, so fdr_shard_set_peers is all
 * the set-up they need).  They replace the synchronisation implied by MPI_Alltoallv (fft_mpi.cpp:170-279) and keep NCCL /
 * MPI out of the per-image path.  fdr_shard_barrier: stream-ordered barrier; every rank calls it with the same `set`
 * (0..15; independent sequences, one per concurrently running pipeline unit and phase) in the same order.
 * fdr_shard_minmax_allreduce: all-reduce of the extrema vector below in one launch (it is also a barrier).  A barrier gives
 * up after 20 s if a peer never arrives; fdr_shard_sync_status synchronises `stream` and reports that.  Shards that share
 * one device (tests) must issue these on one stream per shard, or the waits would queue behind each other. */
int fdr_shard_barrier(fdr_shard* shard, int set, void* stream);
int fdr_shard_minmax_allreduce(fdr_shard* shard, void* stream);
int fdr_shard_sync_status(fdr_shard* shard, void* stream, int* timed_out);
/* [channels][2] floats (min, max of this rank's part of every padded plane) to all-reduce in
 * place: column 0 with MIN, column 1 with MAX (fdr_shard_minmax_allreduce does it over peer memory). */
int fdr_shard_minmax_device(fdr_shard* shard, void** d_minmax_f32);
int fdr_shard_phase4_pack(fdr_shard* shard, void* d_out_rows_u8, void* stream);
int fdr_shard_last_launch_count(const fdr_shard* shard, long long* launches);
/* Rows [first_row, first_row+n_rows) of synthetic image `image` (same stream as
 * fdr_synth_images_device_u8 for that image). */
int fdr_synth_rows_device_u8(void* d_out, uint32_t seed, long long image, int channels, int rows_total, int cols,
                             int first_row, int n_rows, void* stream);

/* motionBlurKernel(size, angle) (utils.hpp:15-24): length x length fp32 PSF built on the device
 * (current CUDA device) and copied to psf_out.  Bit-identical to OpenCV 4.x warpAffine. */
int fdr_motion_psf_host(int length, double angle_deg, float* psf_out);

/* ---- synthetic input + measurement helpers (bench, tests) ----------------------------- */
/* Counter-hash u8 images generated on the device, identical to oracle/orc_synth_u8. */
int fdr_synth_images_device_u8(void* d_out_images, uint32_t seed, long long first_image, int n_images, int channels,
                               int rows, int cols, void* stream);
/* Synchronous copy for harnesses holding raw device pointers: kind 0 = H2D, 1 = D2H, 2 = D2D. */
int fdr_memcpy(void* dst, const void* src, size_t bytes, int kind);
/* Overwrites `bytes` of scratch to evict L2 between timed iterations. */
int fdr_l2_flush_device(void* d_scratch, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FDR_B200_H */
