// TEST / BENCH INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" wrapper over the reference's OWN gpu mode (fft/fft_gpu.cu, compiled unmodified for
// sm_100a by oracle/Makefile) so it can be timed on the same B200 as a side comparison
// ("the kernel to beat").  Its output is NOT a parity oracle (crop-then-normalise,
// stale-input ordering).
#include "fft/fft.hpp"

#include <chrono>
#include <cstring>
#include <vector>

extern "C" {

// planes: n contiguous f32 planes of rows x cols (in place).  which: 0 = RGB_optimized, 1 = RGB_naive.
// Returns wall-clock milliseconds of the call (as gpu.cpp:100-104 measures it).
double ref_gpu_restore(int which, float* planes, int n, int rows, int cols, const float* psf, int prows, int pcols, float K) {
    std::vector<cv::Mat> ch((size_t)n);
    for (int i = 0; i < n; ++i) {
        ch[(size_t)i] = cv::Mat(rows, cols, CV_32F);
        std::memcpy(ch[(size_t)i].data, planes + (size_t)i * rows * cols, sizeof(float) * (size_t)rows * cols);
    }
    cv::Mat p(prows, pcols, CV_32F);
    std::memcpy(p.data, psf, sizeof(float) * (size_t)prows * pcols);
    auto t0 = std::chrono::high_resolution_clock::now();
    if (which == 0)
        fft_gpu::wienerDeblur_RGB_optimized(ch, p, K);
    else
        fft_gpu::wienerDeblur_RGB_naive(ch, p, K);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < n; ++i) std::memcpy(planes + (size_t)i * rows * cols, ch[(size_t)i].data, sizeof(float) * (size_t)rows * cols);
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}
}
