/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Plain-C restatement of the reference's SERIAL restoration path, operation for
 * operation, so that it reproduces the unmodified reference (oracle/_ref, built
 * from /root/reference by oracle/Makefile) BIT FOR BIT on the float planes.  The
 * pinning tests are tests/test_oracle.py (port == reference, PSF == cv2 goldens).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Compile with -ffp-contract=off: every fp32 operation below is one rounding, as
 * in the reference's build (-O2, no -mfma; OpenCV's element-wise ops likewise).
 *
 * Citations are file:line under /root/reference.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.1415926535897932384626433832795 /* OpenCV's CV_PI */

/* ---- fft/fft_serial.cpp:40-68  fft_radix2_inplace ------------------------------
 * in-place bit reversal (:45-51), iterative DIT stages (:53-66) with the fp32
 * twiddle RECURRENCE w *= wlen (:63); forward sign -, inverse sign +, no scaling.
 * std::complex<float> products are (ac-bd, ad+bc) with separate roundings. */
void orc_fft_radix2(float* a, int n, int inverse) {
    if (n <= 1) return;
    int j = 0;
    for (int i = 1; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            float tr = a[2 * i], ti = a[2 * i + 1];
            a[2 * i] = a[2 * j];
            a[2 * i + 1] = a[2 * j + 1];
            a[2 * j] = tr;
            a[2 * j + 1] = ti;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        /* :54  float ang = 2.0f * CV_PI / len * (inverse ? 1.0f : -1.0f);  (double expr -> float) */
        float ang = (float)(2.0f * ORC_PI / len * (inverse ? 1.0f : -1.0f));
        float wlr = cosf(ang), wli = sinf(ang);
        int half = len / 2;
        for (int i = 0; i < n; i += len) {
            float wr = 1.0f, wi = 0.0f;
            for (int k = 0; k < half; ++k) {
                float ur = a[2 * (i + k)], ui = a[2 * (i + k) + 1];
                float xr = a[2 * (i + k + half)], xi = a[2 * (i + k + half) + 1];
                float vr = xr * wr - xi * wi;
                float vi = xr * wi + xi * wr;
                a[2 * (i + k)] = ur + vr;
                a[2 * (i + k) + 1] = ui + vi;
                a[2 * (i + k + half)] = ur - vr;
                a[2 * (i + k + half) + 1] = ui - vi;
                float nr = wr * wlr - wi * wli;
                float ni = wr * wli + wi * wlr;
                wr = nr;
                wi = ni;
            }
        }
    }
}

/* ---- fft/fft_serial.cpp:71-87  dft_naive_inplace (non power-of-two lengths) ---- */
void orc_dft_naive(float* a, int n, int inverse) {
    if (n <= 1) return;
    float* out = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    const float sign = inverse ? 1.0f : -1.0f;
    for (int k = 0; k < n; ++k) {
        float sr = 0.f, si = 0.f;
        for (int t = 0; t < n; ++t) {
            /* :79  float ang = 2.0f * CV_PI * k * t / n * sign;  (evaluated in double) */
            float ang = (float)(2.0f * ORC_PI * k * t / n * sign);
            float wr = cosf(ang), wi = sinf(ang);
            float ar = a[2 * t], ai = a[2 * t + 1];
            float pr = ar * wr - ai * wi;
            float pi = ar * wi + ai * wr;
            sr = sr + pr;
            si = si + pi;
        }
        out[2 * k] = sr;
        out[2 * k + 1] = si;
    }
    memcpy(a, out, sizeof(float) * 2 * (size_t)n);
    free(out);
}

static int orc_is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; } /* utils.hpp:50-52 */

/* ---- fft/fft_serial.cpp:90-108  transform_row_inplace ------------------------- */
void orc_transform_row(float* row, int n, int inverse) {
    if (orc_is_pow2(n))
        orc_fft_radix2(row, n, inverse);
    else
        orc_dft_naive(row, n, inverse);
}

/* ---- fft/fft_serial.cpp:113-139  my_dft2D --------------------------------------
 * rows, transpose, rows, transpose back.  The transposes are pure data movement,
 * so the same numbers come from transforming each column through a gather buffer. */
void orc_dft2d(float* m, int rows, int cols, int inverse) {
    for (int r = 0; r < rows; ++r) orc_transform_row(m + 2 * (size_t)r * cols, cols, inverse);
    float* col = (float*)malloc(sizeof(float) * 2 * (size_t)rows);
    for (int c = 0; c < cols; ++c) {
        for (int r = 0; r < rows; ++r) {
            col[2 * r] = m[2 * ((size_t)r * cols + c)];
            col[2 * r + 1] = m[2 * ((size_t)r * cols + c) + 1];
        }
        orc_transform_row(col, rows, inverse);
        for (int r = 0; r < rows; ++r) {
            m[2 * ((size_t)r * cols + c)] = col[2 * r];
            m[2 * ((size_t)r * cols + c) + 1] = col[2 * r + 1];
        }
    }
    free(col);
}

/* ---- utils.hpp:27-31  nextPowerOfTwo ------------------------------------------ */
int orc_next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

/* ---- utils.hpp:40-47 + fft_serial.cpp:157-171: zero-pad bottom/right, imag = 0 -- */
void orc_pad_complex(const float* src, int rows, int cols, size_t src_stride_elems, float* dst, int prow, int pcol) {
    memset(dst, 0, sizeof(float) * 2 * (size_t)prow * pcol);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) dst[2 * ((size_t)r * pcol + c)] = src[(size_t)r * src_stride_elems + c];
}

/* ---- fft_serial.cpp:186-224  Wiener filter with whole-matrix OpenCV ops --------
 * mag = sqrt(re^2+im^2) (:195), mag2 = mag*mag (:196), denom = mag2 + K (:197),
 * conj(H) (:200), numerator = G * conj(H) as 4 mul + sub/add (:209-210),
 * result = numerator / denom, two true divisions (:219-220).  In place on G. */
void orc_wiener(float* G, const float* H, size_t n, float K) {
    for (size_t i = 0; i < n; ++i) {
        float hr = H[2 * i], hi = H[2 * i + 1];
        float mag = sqrtf(hr * hr + hi * hi);
        float mag2 = mag * mag;
        float denom = mag2 + K;
        float b0 = hr, b1 = -hi;
        float a0 = G[2 * i], a1 = G[2 * i + 1];
        float c0 = a0 * b0 - a1 * b1;
        float c1 = a0 * b1 + a1 * b0;
        G[2 * i] = c0 / denom;
        G[2 * i + 1] = c1 / denom;
    }
}

/* ---- fft_serial.cpp:246  normalize(NORM_MINMAX, 0, 1) ---------------------------
 * min/max as double, scale = (float)(1/(max-min)) (0 if the range <= DBL_EPSILON),
 * shift = -(float)(min*scale), applied as fmaf(x, scale, shift) -- cv2 4.13 fuses the
 * multiply-add (pinned by tests/golden/normalize_*.npy).  In place.
 * mm[0], mm[1] receive min and max when mm != NULL. */
void orc_normalize_minmax(float* x, size_t n, double* mm) {
    double smin = DBL_MAX, smax = -DBL_MAX;
    for (size_t i = 0; i < n; ++i) {
        double v = (double)x[i];
        if (v < smin) smin = v;
        if (v > smax) smax = v;
    }
    /* OpenCV 4.x with a CV_32F destination rounds scale to float before deriving the shift:
     * scale = (float)scale; shift = (float)dmin - (float)(smin*scale)  (pinned against cv2 4.13) */
    double scale = (smax - smin) > DBL_EPSILON ? 1. / (smax - smin) : 0.;
    float a = (float)scale;
    float b = 0.0f - (float)(smin * (double)a);
    for (size_t i = 0; i < n; ++i) x[i] = fmaf(x[i], a, b);
    if (mm) {
        mm[0] = smin;
        mm[1] = smax;
    }
}

/* ---- fft_serial.cpp:141-261  wienerDeblur_myfft ---------------------------------
 * img: rows x cols (powers of two: the drivers pre-pad, serial.cpp:36), psf: prows x
 * pcols anchored top-left (:166-170).  Outputs (any may be NULL):
 *   out_norm  rows x cols  min-max normalised restored plane (the return value)
 *   out_G     rows x cols complex  forward spectrum of the image      (:176)
 *   out_H     rows x cols complex  forward spectrum of the padded PSF (:182)
 *   out_F     rows x cols complex  filtered spectrum                  (:186-224)
 *   out_raw   rows x cols  Re(unscaled inverse transform)             (:229,238-240)
 *   mm        {min, max} of out_raw */
int orc_wiener_deblur(const float* img, int rows, int cols, const float* psf, int prows, int pcols, float K,
                      float* out_norm, float* out_G, float* out_H, float* out_F, float* out_raw, double* mm) {
    if (prows > rows || pcols > cols) return -1;
    size_t n = (size_t)rows * cols;
    float* G = (float*)malloc(sizeof(float) * 2 * n);
    float* H = (float*)malloc(sizeof(float) * 2 * n);
    if (!G || !H) {
        free(G);
        free(H);
        return -2;
    }
    orc_pad_complex(img, rows, cols, (size_t)cols, G, rows, cols);
    orc_pad_complex(psf, prows, pcols, (size_t)pcols, H, rows, cols);
    orc_dft2d(G, rows, cols, 0);
    orc_dft2d(H, rows, cols, 0);
    if (out_G) memcpy(out_G, G, sizeof(float) * 2 * n);
    if (out_H) memcpy(out_H, H, sizeof(float) * 2 * n);
    orc_wiener(G, H, n, K);
    if (out_F) memcpy(out_F, G, sizeof(float) * 2 * n);
    orc_dft2d(G, rows, cols, 1);
    float* re = H; /* reuse */
    for (size_t i = 0; i < n; ++i) re[i] = G[2 * i];
    if (out_raw) memcpy(out_raw, re, sizeof(float) * n);
    orc_normalize_minmax(re, n, mm);
    if (out_norm) memcpy(out_norm, re, sizeof(float) * n);
    free(G);
    free(H);
    return 0;
}

/* ---- utils.hpp:15-24  motionBlurKernel ------------------------------------------
 * S x S zeros, row S/2 = (float)(1.0/S) (:18-19), rotated about (S/2, S/2) by `angle`
 * degrees with getRotationMatrix2D (:20) + warpAffine defaults (:22): bilinear,
 * BORDER_CONSTANT 0.  OpenCV 4.x (un-vendored dependency, pinned empirically to cv2
 * 4.13 by tests/golden/psf_*.npy) evaluates the warp in FIXED POINT:
 *   inverse matrix in double; per column adelta = rint(A00*x*1024), bdelta =
 *   rint(A10*x*1024); per row X0 = rint((A01*y+b0)*1024)+16, Y0 likewise;
 *   X = (X0+adelta)>>5; integer part X>>5, fraction (X&31)/32; 4-tap weights as
 *   fp32 products of fp32 factors; taps accumulated left to right in fp32;
 *   out-of-range taps read 0.
 * out: size*size floats. */
static int orc_rint_sat(double v) { return (int)lrint(v); }

void orc_motion_psf(int size, double angle_deg, float* out) {
    const int S = size;
    float* k = (float*)calloc((size_t)S * S, sizeof(float));
    const int cy = S / 2, cx = S / 2;
    for (int i = 0; i < S; ++i) k[(size_t)cy * S + i] = (float)(1.0 / S);
    /* getRotationMatrix2D(center, angle, 1): center is Point2f */
    double ang = angle_deg * (ORC_PI / 180); /* angle *= CV_PI/180 */
    double alpha = cos(ang) * 1.0, beta = sin(ang) * 1.0;
    double cxf = (double)(float)cx, cyf = (double)(float)cy;
    double M[6];
    M[0] = alpha;
    M[1] = beta;
    M[2] = (1 - alpha) * cxf - beta * cyf;
    M[3] = -beta;
    M[4] = alpha;
    M[5] = beta * cxf + (1 - alpha) * cyf;
    /* warpAffine without WARP_INVERSE_MAP inverts M */
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11;
    M[1] *= -D;
    M[3] *= -D;
    M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5];
    double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1;
    M[5] = b2;
    const int AB_BITS = 10, AB_SCALE = 1 << AB_BITS, INTER_BITS = 5, INTER_TAB = 1 << INTER_BITS;
    const int round_delta = AB_SCALE / INTER_TAB / 2;
    for (int y = 0; y < S; ++y) {
        int X0 = orc_rint_sat((M[1] * y + M[2]) * AB_SCALE) + round_delta;
        int Y0 = orc_rint_sat((M[4] * y + M[5]) * AB_SCALE) + round_delta;
        for (int x = 0; x < S; ++x) {
            int adelta = orc_rint_sat(M[0] * x * AB_SCALE);
            int bdelta = orc_rint_sat(M[3] * x * AB_SCALE);
            int X = (X0 + adelta) >> (AB_BITS - INTER_BITS);
            int Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS);
            int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
            float fx = (float)(X & (INTER_TAB - 1)) * (1.f / INTER_TAB);
            float fy = (float)(Y & (INTER_TAB - 1)) * (1.f / INTER_TAB);
            float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx;
            float w10 = fy * (1.f - fx), w11 = fy * fx;
            float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
            if (sy >= 0 && sy < S) {
                if (sx >= 0 && sx < S) t00 = k[(size_t)sy * S + sx];
                if (sx + 1 >= 0 && sx + 1 < S) t01 = k[(size_t)sy * S + sx + 1];
            }
            if (sy + 1 >= 0 && sy + 1 < S) {
                if (sx >= 0 && sx < S) t10 = k[(size_t)(sy + 1) * S + sx];
                if (sx + 1 >= 0 && sx + 1 < S) t11 = k[(size_t)(sy + 1) * S + sx + 1];
            }
            float acc = t00 * w00;
            acc = acc + t01 * w01;
            acc = acc + t10 * w10;
            acc = acc + t11 * w11;
            out[(size_t)y * S + x] = acc;
        }
    }
    free(k);
}

/* ---- serial.cpp:54 / gpu.cpp:134  convertTo(CV_8U, 255.0) -----------------------
 * u8 = saturate(rint(255*x)), round-half-even. */
void orc_pack_u8(const float* x, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) {
        float v = x[i] * 255.0f;
        long q = lrintf(v);
        out[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
    }
}

/* ---- SURVEY.md 8(d) synthetic input generator (counter hash, iid uniform u8) ---- */
static uint32_t orc_lowbias32(uint32_t v) {
    v ^= v >> 16;
    v *= 0x7feb352dU;
    v ^= v >> 15;
    v *= 0x846ca68bU;
    v ^= v >> 16;
    return v;
}

/* fills out[0..count) with pixels idx0 .. idx0+count-1 of the stream for `seed` */
void orc_synth_u8(uint32_t seed, uint64_t idx0, uint64_t count, uint8_t* out) {
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t idx = idx0 + i;
        uint32_t hi = orc_lowbias32(seed + (uint32_t)(idx >> 32));
        out[i] = (uint8_t)(orc_lowbias32((uint32_t)idx ^ hi) >> 24);
    }
}
