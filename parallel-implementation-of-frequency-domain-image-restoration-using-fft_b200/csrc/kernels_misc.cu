// kernels_misc.cu -- PSF build, min/max bookkeeping, normalise + pack, synthetic input,
// naive DFT for non power-of-two lengths, L2 flush.
#include <float.h>

#include "passes.h"

namespace fdr {

// ---------------------------------------------------------------------------------
// Motion-blur PSF (reference utils.hpp:15-24: S x S zeros, row S/2 = 1/S, rotated about
// (S/2,S/2) with getRotationMatrix2D + warpAffine defaults).  OpenCV evaluates the warp in
// fixed point (AB_BITS = 10, INTER_BITS = 5); this kernel reproduces that arithmetic so the
// result is bit-identical to cv2 4.13 (tests/golden/psf_*.npy).  The inverse matrix comes
// from the host in double; products and sums below are individually rounded (no FMA).
// ---------------------------------------------------------------------------------
__global__ void motion_psf_kernel(float* psf, int S, PsfAffine M) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= S || y >= S) return;
    const double AB = 1024.0;
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M.a01, (double)y), M.b0), AB)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M.a11, (double)y), M.b1), AB)) + 16;
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M.a00, (double)x), AB));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M.a10, (double)x), AB));
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = X >> 5, sy = Y >> 5;
    const float fx = (float)(X & 31) * (1.f / 32.f), fy = (float)(Y & 31) * (1.f / 32.f);
    const float w00 = __fmul_rn(1.f - fy, 1.f - fx), w01 = __fmul_rn(1.f - fy, fx);
    const float w10 = __fmul_rn(fy, 1.f - fx), w11 = __fmul_rn(fy, fx);
    // source kernel: only row S/2 is non-zero, value (float)(1.0/S)
    const float val = (float)(1.0 / (double)S);
    const int cy = S / 2;
    const bool xin0 = (sx >= 0 && sx < S), xin1 = (sx + 1 >= 0 && sx + 1 < S);
    const float t00 = (sy == cy && xin0) ? val : 0.f;
    const float t01 = (sy == cy && xin1) ? val : 0.f;
    const float t10 = (sy + 1 == cy && xin0) ? val : 0.f;
    const float t11 = (sy + 1 == cy && xin1) ? val : 0.f;
    float acc = __fmul_rn(t00, w00);
    acc = __fadd_rn(acc, __fmul_rn(t01, w01));
    acc = __fadd_rn(acc, __fmul_rn(t10, w10));
    acc = __fadd_rn(acc, __fmul_rn(t11, w11));
    psf[(size_t)y * S + x] = acc;
}

cudaError_t launch_motion_psf(float* psf, int size, PsfAffine m, cudaStream_t s) {
    dim3 b(16, 16), g((size + 15) / 16, (size + 15) / 16);
    motion_psf_kernel<<<g, b, 0, s>>>(psf, size, m);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// min/max bookkeeping.  Pass 3 keeps per-plane extrema as order-preserving uints.
// ---------------------------------------------------------------------------------
__global__ void minmax_reset_kernel(unsigned int* mm, int units) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one (unit, slot) entry each
    if (i < units * FDR_MINMAX_SLOTS) {
        mm[2 * i] = 0xFFFFFFFFu;
        mm[2 * i + 1] = 0u;
    }
}
cudaError_t launch_minmax_reset(unsigned int* minmax, int units, cudaStream_t s) {
    minmax_reset_kernel<<<(units * FDR_MINMAX_SLOTS + 127) / 128, 128, 0, s>>>(minmax, units);
    return cudaGetLastError();
}

__device__ __forceinline__ float f32_from_ordered(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// fold the FDR_MINMAX_SLOTS slots of plane i
__device__ __forceinline__ void minmax_fold(const unsigned int* mm, int i, unsigned int& lo, unsigned int& hi) {
    lo = 0xFFFFFFFFu;
    hi = 0u;
    const uint2* e = reinterpret_cast<const uint2*>(mm) + (size_t)i * FDR_MINMAX_SLOTS;
    for (int k = 0; k < FDR_MINMAX_SLOTS; ++k) {
        const uint2 v = e[k];
        lo = min(lo, v.x);
        hi = max(hi, v.y);
    }
}

// cv::normalize(NORM_MINMAX, 0, 1) (fft_serial.cpp:246) as OpenCV 4.x evaluates it for a CV_32F
// destination: scale = (float)(1/(max-min)) (0 when the range <= DBL_EPSILON), shift =
// -(float)(min*scale) with the product in double; applied as one fused multiply-add.
__global__ void minmax_finalize_kernel(const unsigned int* mm, float2* ss, float* mmf, int units) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= units) return;
    unsigned int lo, hi;
    minmax_fold(mm, i, lo, hi);
    const double smin = (double)f32_from_ordered(lo);
    const double smax = (double)f32_from_ordered(hi);
    const double scale = (smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0;
    const float a = (float)scale;
    const float b = 0.0f - (float)__dmul_rn(smin, (double)a);
    ss[i] = make_float2(a, b);
    if (mmf) {
        mmf[2 * i] = (float)smin;
        mmf[2 * i + 1] = (float)smax;
    }
}
// negate_max: the vector holds (min, -max), so that a single all-reduce(MIN) across ranks folds both extrema
__global__ void minmax_decode_kernel(const unsigned int* mm, float* mmf, int units, int negate_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= units) return;
    unsigned int lo, hi;
    minmax_fold(mm, i, lo, hi);
    const float mx = f32_from_ordered(hi);
    mmf[2 * i] = f32_from_ordered(lo);
    mmf[2 * i + 1] = negate_max ? -mx : mx;
}
cudaError_t launch_minmax_decode(const unsigned int* minmax, float* minmax_f32, int units, int negate_max, cudaStream_t s) {
    minmax_decode_kernel<<<(units + 127) / 128, 128, 0, s>>>(minmax, minmax_f32, units, negate_max);
    return cudaGetLastError();
}

__global__ void scale_shift_from_f32_kernel(const float* mmf, float2* ss, int units, int negated_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= units) return;
    const double smin = (double)mmf[2 * i], smax = negated_max ? -(double)mmf[2 * i + 1] : (double)mmf[2 * i + 1];
    const double scale = (smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0;
    const float a = (float)scale;
    ss[i] = make_float2(a, 0.0f - (float)__dmul_rn(smin, (double)a));
}
cudaError_t launch_scale_shift_from_f32(const float* minmax_f32, float2* scale_shift, int units, int negated_max, cudaStream_t s) {
    scale_shift_from_f32_kernel<<<(units + 127) / 128, 128, 0, s>>>(minmax_f32, scale_shift, units, negated_max);
    return cudaGetLastError();
}

// Column N/2 of the PSF's row spectra without a transform: H_y[N/2] = sum_x h[y][x] * (-1)^x (real).  One thread per PSF
// row; rows are a few hundred pixels at most.
__global__ void psf_nyquist_kernel(const float* __restrict__ psf, int rows, int cols, float2* __restrict__ out) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= rows) return;
    const float* r = psf + (size_t)y * cols;
    float even = 0.f, odd = 0.f;
    for (int x = 0; x + 1 < cols; x += 2) {
        even += r[x];
        odd += r[x + 1];
    }
    if (cols & 1) even += r[cols - 1];
    out[y] = make_float2(even - odd, 0.f);
}
cudaError_t launch_psf_nyquist(const float* psf, int rows, int cols, float2* out, cudaStream_t s) {
    psf_nyquist_kernel<<<(rows + 63) / 64, 64, 0, s>>>(psf, rows, cols, out);
    return cudaGetLastError();
}

cudaError_t launch_minmax_finalize(const unsigned int* minmax, float2* scale_shift, float* minmax_f32, int units,
                                   cudaStream_t s) {
    minmax_finalize_kernel<<<(units + 127) / 128, 128, 0, s>>>(minmax, scale_shift, minmax_f32, units);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// Pass 4a: normalise (one fused multiply-add, as cv2's convertTo) + 8-bit pack
// u8 = saturate(rint(255 * n))  (serial.cpp:54 convertTo(CV_8U, 255.0)).
// Output interleaved [img][H][W][C], B,G,R order = unit order.
// ---------------------------------------------------------------------------------
template <int C>
__global__ void pack_u8_kernel(const float* raw, long long ustride, const float2* ss, uint8_t* out, int rows, int cols) {
    const int img = blockIdx.y;
    const long long npx = (long long)rows * cols;
    const float* r = raw + (long long)img * C * ustride;
    uint8_t* o = out + (long long)img * npx * C;
    float2 k[C];
#pragma unroll
    for (int c = 0; c < C; ++c) k[c] = ss[img * C + c];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float n = fmaf(__ldg(r + c * ustride + i), k[c].x, k[c].y);
            int q = __float2int_rn(n * 255.0f);
            q = min(max(q, 0), 255);
            o[i * C + c] = (uint8_t)q;
        }
    }
}

// C = 3, four pixels per thread: three 16-byte plane loads, three 4-byte stores (12 output bytes).
__global__ void pack_u8_c3_vec4_kernel(const float* __restrict__ raw, long long ustride, const float2* __restrict__ ss,
                                       uint8_t* __restrict__ out, long long nquads) {
    const int img = blockIdx.y;
    const float4* r0 = reinterpret_cast<const float4*>(raw + (long long)img * 3 * ustride);
    const float4* r1 = reinterpret_cast<const float4*>(raw + ((long long)img * 3 + 1) * ustride);
    const float4* r2 = reinterpret_cast<const float4*>(raw + ((long long)img * 3 + 2) * ustride);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (long long)img * nquads * 12);
    const float2 k0 = ss[img * 3], k1 = ss[img * 3 + 1], k2 = ss[img * 3 + 2];
    auto q8 = [](float v, float2 k) -> uint32_t {
        int q = __float2int_rn(fmaf(v, k.x, k.y) * 255.0f);
        return (uint32_t)min(max(q, 0), 255);
    };
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nquads; j += (long long)gridDim.x * blockDim.x) {
        const float4 b = __ldg(r0 + j), g = __ldg(r1 + j), r = __ldg(r2 + j);
        const uint32_t b0 = q8(b.x, k0), g0 = q8(g.x, k1), c0 = q8(r.x, k2);
        const uint32_t b1 = q8(b.y, k0), g1 = q8(g.y, k1), c1 = q8(r.y, k2);
        const uint32_t b2 = q8(b.z, k0), g2 = q8(g.z, k1), c2 = q8(r.z, k2);
        const uint32_t b3 = q8(b.w, k0), g3 = q8(g.w, k1), c3 = q8(r.w, k2);
        o[3 * j] = b0 | (g0 << 8) | (c0 << 16) | (b1 << 24);
        o[3 * j + 1] = g1 | (c1 << 8) | (b2 << 16) | (g2 << 24);
        o[3 * j + 2] = c2 | (b3 << 8) | (g3 << 16) | (c3 << 24);
    }
}

// The same pack for a FEW images with launch_minmax_finalize folded in: warps 0-2 of every CTA fold the slots of the image's three
// planes and derive scale / shift exactly as minmax_finalize_kernel does; CTA 0 of an image also publishes them (last_minmax API).
__global__ void pack_u8_c3_vec4_fused_kernel(const float* __restrict__ raw, long long ustride, const unsigned int* __restrict__ mm,
                                             float2* __restrict__ ss_out, float* __restrict__ mmf_out, uint8_t* __restrict__ out,
                                             long long nquads) {
    __shared__ float2 kk[3];
    const int img = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 3) {
        const int u = img * 3 + warp;
        const uint2* e = reinterpret_cast<const uint2*>(mm) + (size_t)u * FDR_MINMAX_SLOTS;
        unsigned int lo = 0xFFFFFFFFu, hi = 0u;
        for (int k = lane; k < FDR_MINMAX_SLOTS; k += 32) {
            const uint2 v = e[k];
            lo = min(lo, v.x);
            hi = max(hi, v.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            const double smin = (double)f32_from_ordered(lo), smax = (double)f32_from_ordered(hi);
            const double scale = (smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0;
            const float a = (float)scale;
            const float b = 0.0f - (float)__dmul_rn(smin, (double)a);
            kk[warp] = make_float2(a, b);
            if (blockIdx.x == 0) {
                ss_out[u] = make_float2(a, b);
                if (mmf_out) {
                    mmf_out[2 * u] = (float)smin;
                    mmf_out[2 * u + 1] = (float)smax;
                }
            }
        }
    }
    __syncthreads();
    const float4* r0 = reinterpret_cast<const float4*>(raw + (long long)img * 3 * ustride);
    const float4* r1 = reinterpret_cast<const float4*>(raw + ((long long)img * 3 + 1) * ustride);
    const float4* r2 = reinterpret_cast<const float4*>(raw + ((long long)img * 3 + 2) * ustride);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (long long)img * nquads * 12);
    const float2 k0 = kk[0], k1 = kk[1], k2 = kk[2];
    auto q8 = [](float v, float2 k) -> uint32_t {
        int q = __float2int_rn(fmaf(v, k.x, k.y) * 255.0f);
        return (uint32_t)min(max(q, 0), 255);
    };
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nquads; j += (long long)gridDim.x * blockDim.x) {
        const float4 b = __ldg(r0 + j), g = __ldg(r1 + j), r = __ldg(r2 + j);
        const uint32_t b0 = q8(b.x, k0), g0 = q8(g.x, k1), c0 = q8(r.x, k2);
        const uint32_t b1 = q8(b.y, k0), g1 = q8(g.y, k1), c1 = q8(r.y, k2);
        const uint32_t b2 = q8(b.z, k0), g2 = q8(g.z, k1), c2 = q8(r.z, k2);
        const uint32_t b3 = q8(b.w, k0), g3 = q8(g.w, k1), c3 = q8(r.w, k2);
        o[3 * j] = b0 | (g0 << 8) | (c0 << 16) | (b1 << 24);
        o[3 * j + 1] = g1 | (c1 << 8) | (b2 << 16) | (g2 << 24);
        o[3 * j + 2] = c2 | (b3 << 8) | (g3 << 16) | (c3 << 24);
    }
}
bool pack_u8_c3_fused_applicable(const float* raw, long long raw_unit_stride, const uint8_t* out, int imgs, int channels, int rows, int cols) {
    const long long npx = (long long)rows * cols;
    return channels == 3 && imgs <= 2 && npx % 4 == 0 && raw_unit_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 3) == 0;
}
cudaError_t launch_pack_u8_c3_fused(const float* raw, long long raw_unit_stride, const unsigned int* minmax, float2* scale_shift, float* minmax_f32,
                                    uint8_t* out, int imgs, int rows, int cols, cudaStream_t s) {
    const long long npx = (long long)rows * cols;
    int bq = (int)((npx / 4 + 255) / 256);
    if (bq > 148 * 8) bq = 148 * 8;
    pack_u8_c3_vec4_fused_kernel<<<dim3(bq, imgs), 256, 0, s>>>(raw, raw_unit_stride, minmax, scale_shift, minmax_f32, out, npx / 4);
    return cudaGetLastError();
}

__global__ void pack_u8_generic_kernel(const float* raw, long long ustride, const float2* ss, uint8_t* out, int C,
                                       int rows, int cols) {
    const int img = blockIdx.y;
    const long long npx = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < C; ++c) {
            const float2 k = ss[img * C + c];
            const float n = fmaf(raw[((long long)img * C + c) * ustride + i], k.x, k.y);
            int q = __float2int_rn(n * 255.0f);
            out[((long long)img * npx + i) * C + c] = (uint8_t)min(max(q, 0), 255);
        }
}

cudaError_t launch_pack_u8(const float* raw, long long raw_unit_stride, const float2* scale_shift, uint8_t* out,
                           int imgs, int channels, int rows, int cols, cudaStream_t s) {
    const long long npx = (long long)rows * cols;
    int bx = (int)((npx + 255) / 256);
    if (bx > 148 * 8) bx = 148 * 8;
    dim3 g(bx, imgs);
    if (channels == 3 && npx % 4 == 0 && raw_unit_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        int bq = (int)((npx / 4 + 255) / 256);
        if (bq > 148 * 8) bq = 148 * 8;
        pack_u8_c3_vec4_kernel<<<dim3(bq, imgs), 256, 0, s>>>(raw, raw_unit_stride, scale_shift, out, npx / 4);
    } else if (channels == 3)
        pack_u8_kernel<3><<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, out, rows, cols);
    else if (channels == 1)
        pack_u8_kernel<1><<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, out, rows, cols);
    else
        pack_u8_generic_kernel<<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, out, channels, rows, cols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// Pass 4c: the post stage of every reference driver (gpu.cpp:123-134, serial.cpp:43-54):
//   merged BGR -> Lab, L *= mean(L_original)/(mean(L_restored)+1e-6), clamp L to [0,100]
//   (applyWhiteBalance, utils.hpp:55-71), Lab -> BGR, convertTo(CV_8U, 255).
// Lab follows OpenCV's definition for float images (sRGB gamma, D65 white point
// {0.950456, 1, 1.088754}, L in [0,100]) evaluated with the closed-form expressions; OpenCV itself
// interpolates a 33^3 fixed-point table for float input, so the two agree to within +-1 LSB of the
// 8-bit result (tests/test_gpu_whitebalance.py), not bit for bit.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float srgb_to_linear(float c) {
    return c <= 0.04045f ? c * (1.0f / 12.92f) : powf((c + 0.055f) * (1.0f / 1.055f), 2.4f);
}
__device__ __forceinline__ float linear_to_srgb(float c) {
    c = fminf(fmaxf(c, 0.f), 1.f);
    return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float lab_f(float t) { return t > 0.008856f ? cbrtf(t) : 7.787f * t + 16.0f / 116.0f; }
__device__ __forceinline__ float3 bgr_to_lab(float b, float g, float r) {
    const float R = srgb_to_linear(r), G = srgb_to_linear(g), B = srgb_to_linear(b);
    const float X = (0.412453f * R + 0.357580f * G + 0.180423f * B) * (1.0f / 0.950456f);
    const float Y = 0.212671f * R + 0.715160f * G + 0.072169f * B;
    const float Z = (0.019334f * R + 0.119193f * G + 0.950227f * B) * (1.0f / 1.088754f);
    const float fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
    const float L = Y > 0.008856f ? 116.0f * fy - 16.0f : 903.3f * Y;
    return make_float3(L, 500.0f * (fx - fy), 200.0f * (fy - fz));
}
__device__ __forceinline__ float lab_finv(float t) {
    return t > (7.787f * 0.008856f + 16.0f / 116.0f) ? t * t * t : (t - 16.0f / 116.0f) * (1.0f / 7.787f);
}
__device__ __forceinline__ float3 lab_to_bgr(float L, float a, float b) {
    float fy, Y;
    if (L <= 0.008856f * 903.3f) {
        Y = L * (1.0f / 903.3f);
        fy = 7.787f * Y + 16.0f / 116.0f;
    } else {
        fy = (L + 16.0f) * (1.0f / 116.0f);
        Y = fy * fy * fy;
    }
    const float X = lab_finv(fy + a * (1.0f / 500.0f)) * 0.950456f;
    const float Z = lab_finv(fy - b * (1.0f / 200.0f)) * 1.088754f;
    const float R = 3.240479f * X - 1.537150f * Y - 0.498535f * Z;
    const float G = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
    const float B = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
    return make_float3(linear_to_srgb(B), linear_to_srgb(G), linear_to_srgb(R));
}

// sums[img][0] += sum of L over the original image, sums[img][1] += sum of L over the restored image
__global__ void wb_stats_kernel(const float* __restrict__ raw, long long ustride, const float2* __restrict__ ss,
                                const uint8_t* __restrict__ orig_u8, const float* __restrict__ orig_f32, long long orig_ustride,
                                double* sums, long long npx) {
    const int img = blockIdx.y;
    const float* r = raw + (long long)img * 3 * ustride;
    const float2 k0 = ss[img * 3], k1 = ss[img * 3 + 1], k2 = ss[img * 3 + 2];
    double so = 0.0, sd = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        float ob, og, orr;
        if (orig_u8) {
            const uint8_t* o = orig_u8 + ((long long)img * npx + i) * 3;
            const float inv255 = (float)(1.0 / 255.0);
            ob = (float)o[0] * inv255;
            og = (float)o[1] * inv255;
            orr = (float)o[2] * inv255;
        } else {
            const float* o = orig_f32 + (long long)img * 3 * orig_ustride + i;
            ob = o[0];
            og = o[orig_ustride];
            orr = o[2 * orig_ustride];
        }
        so += (double)bgr_to_lab(ob, og, orr).x;
        const float nb = fmaf(r[i], k0.x, k0.y), ng = fmaf(r[ustride + i], k1.x, k1.y), nr = fmaf(r[2 * ustride + i], k2.x, k2.y);
        sd += (double)bgr_to_lab(nb, ng, nr).x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        so += __shfl_xor_sync(0xffffffffu, so, o);
        sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    __shared__ double red[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[warp][0] = so;
        red[warp][1] = sd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += red[w][0];
            b += red[w][1];
        }
        atomicAdd(&sums[2 * img], a);
        atomicAdd(&sums[2 * img + 1], b);
    }
}

__global__ void wb_apply_kernel(const float* __restrict__ raw, long long ustride, const float2* __restrict__ ss,
                                const double* __restrict__ sums, uint8_t* __restrict__ out, long long npx) {
    const int img = blockIdx.y;
    const float* r = raw + (long long)img * 3 * ustride;
    uint8_t* o = out + (long long)img * npx * 3;
    const float2 k0 = ss[img * 3], k1 = ss[img * 3 + 1], k2 = ss[img * 3 + 2];
    // gain = avgL_orig / (avgL_deblur + 1e-6) in double (utils.hpp:60-62); applied as a float factor
    const float gain = (float)((sums[2 * img] / (double)npx) / (sums[2 * img + 1] / (double)npx + 1e-6));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        const float nb = fmaf(r[i], k0.x, k0.y), ng = fmaf(r[ustride + i], k1.x, k1.y), nr = fmaf(r[2 * ustride + i], k2.x, k2.y);
        float3 lab = bgr_to_lab(nb, ng, nr);
        lab.x = fminf(fmaxf(lab.x * gain, 0.0f), 100.0f);
        const float3 bgr = lab_to_bgr(lab.x, lab.y, lab.z);
        o[3 * i] = (uint8_t)min(max(__float2int_rn(bgr.x * 255.0f), 0), 255);
        o[3 * i + 1] = (uint8_t)min(max(__float2int_rn(bgr.y * 255.0f), 0), 255);
        o[3 * i + 2] = (uint8_t)min(max(__float2int_rn(bgr.z * 255.0f), 0), 255);
    }
}

cudaError_t launch_white_balance_pack_u8(const float* raw, long long raw_unit_stride, const float2* scale_shift,
                                         const uint8_t* orig_u8, const float* orig_f32, long long orig_unit_stride, double* sums,
                                         uint8_t* out, int imgs, int rows, int cols, cudaStream_t s) {
    const long long npx = (long long)rows * cols;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * imgs, s);
    if (e != cudaSuccess) return e;
    int bx = (int)((npx + 255) / 256);
    if (bx > 148 * 4) bx = 148 * 4;
    dim3 g(bx, imgs);
    wb_stats_kernel<<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, orig_u8, orig_f32, orig_unit_stride, sums, npx);
    wb_apply_kernel<<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, sums, out, npx);
    return cudaGetLastError();
}

// Pass 4b: normalised f32 planes (what fft_gpu::wienerDeblur_RGB_* hands back, fft_gpu.cu:379-384).
__global__ void normalize_f32_kernel(const float* raw, long long ustride, const float2* ss, float* out,
                                     long long ostride, long long npx) {
    const int u = blockIdx.y;
    const float2 k = ss[u];
    const float* r = raw + (long long)u * ustride;
    float* o = out + (long long)u * ostride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x)
        o[i] = fmaf(__ldg(r + i), k.x, k.y);
}
cudaError_t launch_normalize_f32(const float* raw, long long raw_unit_stride, const float2* scale_shift, float* out,
                                 long long out_unit_stride, int units, int rows, int cols, cudaStream_t s) {
    const long long npx = (long long)rows * cols;
    int bx = (int)((npx + 255) / 256);
    if (bx > 148 * 8) bx = 148 * 8;
    dim3 g(bx, units);
    normalize_f32_kernel<<<g, 256, 0, s>>>(raw, raw_unit_stride, scale_shift, out, out_unit_stride, npx);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// Synthetic input: counter hash, iid uniform u8, reproducible on the host
// (oracle/wiener_oracle.c orc_synth_u8).  idx = ((img*C + c)*H + y)*W + x.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lowbias32(uint32_t v) {
    v ^= v >> 16;
    v *= 0x7feb352dU;
    v ^= v >> 15;
    v *= 0x846ca68bU;
    v ^= v >> 16;
    return v;
}
__global__ void synth_u8_kernel(uint8_t* out, uint32_t seed, long long img0, int C, long long plane_px, long long px0,
                                long long npx) {
    const int img = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < C; ++c) {
            const unsigned long long idx = ((unsigned long long)(img0 + img) * C + c) * (unsigned long long)plane_px + px0 + i;
            const uint32_t hi = lowbias32(seed + (uint32_t)(idx >> 32));
            out[((long long)img * npx + i) * C + c] = (uint8_t)(lowbias32((uint32_t)idx ^ hi) >> 24);
        }
}
// Pixels [px0, px0 + npx) of every plane (plane_px pixels each) of images img0 .. img0+imgs-1.
cudaError_t launch_synth_u8(uint8_t* out, uint32_t seed, long long img0, int imgs, int channels, long long plane_px,
                            long long px0, long long npx, cudaStream_t s) {
    int bx = (int)((npx + 255) / 256);
    if (bx > 148 * 8) bx = 148 * 8;
    if (bx < 1) bx = 1;
    dim3 g(bx, imgs);
    synth_u8_kernel<<<g, 256, 0, s>>>(out, seed, img0, channels, plane_px, px0, npx);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// O(n^2) DFT for lengths that are not powers of two (fft_serial.cpp:71-87 / the declared
// but undefined fft_gpu::dft_naive_kernel, fft.hpp:36).  Not on the hot path.
// ---------------------------------------------------------------------------------
__global__ void dft_naive_kernel(const float2* in, float2* out, int n, long long elem_stride, long long batch_stride,
                                 int inverse) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float2* src = in + (long long)blockIdx.y * batch_stride;
    float2* dst = out + (long long)blockIdx.y * batch_stride;
    float sr = 0.f, si = 0.f;
    for (int t = 0; t < n; ++t) {
        const long long kt = ((long long)k * t) % n;  // exact angle reduction
        float s, c;
        sincospif(2.0f * (float)kt / (float)n, &s, &c);
        if (!inverse) s = -s;
        const float2 a = src[(long long)t * elem_stride];
        sr += a.x * c - a.y * s;
        si += a.x * s + a.y * c;
    }
    dst[(long long)k * elem_stride] = make_float2(sr, si);
}
cudaError_t launch_dft_naive(const float2* in, float2* out, int n, long long elem_stride, int batch,
                             long long batch_stride, int inverse, cudaStream_t s) {
    dim3 g((n + 127) / 128, batch);
    dft_naive_kernel<<<g, 128, 0, s>>>(in, out, n, elem_stride, batch_stride, inverse);
    return cudaGetLastError();
}

__global__ void l2_flush_kernel(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4((unsigned)i, 0u, 0u, 0u);
}
cudaError_t launch_l2_flush(void* buf, size_t bytes, cudaStream_t s) {
    l2_flush_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<uint4*>(buf), bytes / 16);
    return cudaGetLastError();
}

}  // namespace fdr
