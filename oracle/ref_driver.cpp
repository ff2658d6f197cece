// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" entry points over the reference's OWN CPU implementations, compiled
// unmodified from /root/reference (see oracle/Makefile).  This file contains no
// arithmetic of its own: it wraps raw float buffers in cv::Mat (the shim in
// oracle/cvshim) and forwards to
//   fft_serial::fft_radix2_inplace   /root/reference/fft/fft_serial.cpp:40-68
//   fft_serial::my_dft2D             /root/reference/fft/fft_serial.cpp:113-139
//   fft_serial::wienerDeblur_myfft   /root/reference/fft/fft_serial.cpp:141-261
// and the same three functions of fft_openmp (fft_openmp.cpp) and fft_simd
// (fft_simd.cpp).  The caller pre-pads planes to powers of two exactly as the
// reference drivers do (serial.cpp:36, gpu.cpp:85).
#include "fft/fft.hpp"

#include <complex>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

cv::Mat wrap_f32(const float* src, int rows, int cols) {
    cv::Mat m(rows, cols, CV_32F);
    std::memcpy(m.data, src, sizeof(float) * (size_t)rows * (size_t)cols);
    return m;
}

cv::Mat wrap_c32(const float* src, int rows, int cols) {
    cv::Mat m(rows, cols, CV_32FC2);
    std::memcpy(m.data, src, 2 * sizeof(float) * (size_t)rows * (size_t)cols);
    return m;
}

void unwrap(const cv::Mat& m, float* dst) {
    const size_t rowBytes = (size_t)m.cols * m.elemSize();
    for (int r = 0; r < m.rows; ++r)
        std::memcpy(reinterpret_cast<unsigned char*>(dst) + rowBytes * (size_t)r, m.data + m.step * (size_t)r,
                    rowBytes);
}

template <typename F> void fft1d(F f, float* data, int n, int inverse) {
    std::vector<std::complex<float>> a((size_t)n);
    for (int i = 0; i < n; ++i) a[(size_t)i] = std::complex<float>(data[2 * i], data[2 * i + 1]);
    f(a, inverse != 0);
    for (int i = 0; i < n; ++i) {
        data[2 * i] = a[(size_t)i].real();
        data[2 * i + 1] = a[(size_t)i].imag();
    }
}

}  // namespace

extern "C" {

// mode: 0 = serial, 1 = openmp, 2 = simd
int ref_fft1d(int mode, float* interleaved, int n, int inverse) {
    switch (mode) {
        case 0: fft1d(fft_serial::fft_radix2_inplace, interleaved, n, inverse); return 0;
        case 1: fft1d(fft_openmp::fft_radix2_inplace, interleaved, n, inverse); return 0;
        case 2:
            // fft_simd defines its radix-2 on split re/im vectors (fft_simd.cpp:29); the interleaved
            // entry point of that mode is transform_row_inplace (fft_simd.cpp:178).
            fft_simd::transform_row_inplace(reinterpret_cast<cv::Vec2f*>(interleaved), n, inverse != 0);
            return 0;
    }
    return -1;
}

int ref_dft_naive(float* interleaved, int n, int inverse) {
    fft1d(fft_serial::dft_naive_inplace, interleaved, n, inverse);
    return 0;
}

int ref_dft2d(int mode, float* interleaved, int rows, int cols, int inverse) {
    cv::Mat m = wrap_c32(interleaved, rows, cols);
    switch (mode) {
        case 0: fft_serial::my_dft2D(m, inverse != 0); break;
        case 1: fft_openmp::my_dft2D(m, inverse != 0); break;
        case 2: fft_simd::my_dft2D(m, inverse != 0); break;
        default: return -1;
    }
    unwrap(m, interleaved);
    return 0;
}

// img: rows x cols f32 (already padded to powers of two), psf: prows x pcols f32.
// out: rows x cols f32, min-max normalised (the function's own return value).
int ref_wiener(int mode, const float* img, int rows, int cols, const float* psf, int prows, int pcols, float K,
               float* out) {
    cv::Mat mi = wrap_f32(img, rows, cols);
    cv::Mat mp = wrap_f32(psf, prows, pcols);
    cv::Mat r;
    switch (mode) {
        case 0: r = fft_serial::wienerDeblur_myfft(mi, mp, K); break;
        case 1: r = fft_openmp::wienerDeblur_myfft(mi, mp, K); break;
        case 2: r = fft_simd::wienerDeblur_myfft(mi, mp, K); break;
        default: return -1;
    }
    if (r.rows != rows || r.cols != cols) return -2;
    unwrap(r, out);
    return 0;
}

void ref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
