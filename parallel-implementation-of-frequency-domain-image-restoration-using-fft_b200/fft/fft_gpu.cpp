// fft/fft_gpu.cpp -- namespace fft_gpu of fft/fft.hpp implemented on the C ABI (include/fdr_b200.h).
//
// Host logic only (argument marshalling, the reference's stdout profile block, its
// print-and-exit(1) error convention, fft_gpu.cu:45-66); all arithmetic runs in the CUDA library.
#include "fft.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "../../include/fdr_b200.h"

namespace fft_gpu {

#define CHECK_FDR(call)                                                                           \
    do {                                                                                          \
        if ((call) != FDR_OK) {                                                                   \
            std::cerr << "Error: " << __FILE__ << ":" << __LINE__ << ", " << fdr_last_error() << std::endl; \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

// Same text as the reference's Profiler::print (fft_gpu.cu:45-56).
static void print_profile(const std::string& title, const float t[6]) {
    cout << "=== " << title << " Profiling (3 Channels) ===" << endl;
    cout << "[1. Allocation]  Time: " << t[0] << " ms (MallocHost + Malloc)" << endl;
    cout << "[2. H2D Copy]    Time: " << t[1] << " ms (Raw Img + PSF + Twiddle)" << endl;
    cout << "[3. Pre-process] Time: " << t[2] << " ms (Padding + PSF FFT)" << endl;
    cout << "[4. GPU Compute] Time: " << t[3] << " ms (FFT + Filter + IFFT)" << endl;
    cout << "[5. D2H Copy]    Time: " << t[4] << " ms (Result Transfer)" << endl;
    cout << "[6. Post-process]Time: " << t[5] << " ms (Normalize + CPU Copy)" << endl;
    cout << "--------------------------------------------" << endl;
    cout << "Total (Sum)      Time: " << (t[0] + t[1] + t[2] + t[3] + t[4] + t[5]) << " ms" << endl;
    cout << "============================================" << endl;
}

// One plan per call, like the reference, which allocates and frees everything inside the call
// (fft_gpu.cu:304-322, 389-393).  `keep` reuses it across the channels of the call (optimized)
// or rebuilds it per channel (naive, fft_gpu.cu:429-447).
static void restore(vector<Mat>& channels, const Mat& psf, float K, bool per_channel_alloc, const char* title) {
    if (channels.empty()) return;
    const int rows = channels[0].rows, cols = channels[0].cols;
    const int n = (int)channels.size();
    Mat psfc = psf.isContinuous() ? psf : psf.clone();
    float total[6] = {0, 0, 0, 0, 0, 0};
    vector<Mat> out((size_t)n);
    for (int i = 0; i < n; ++i) out[(size_t)i] = Mat(rows, cols, CV_32F);

    auto run = [&](int first, int count) {
        fdr_plan* plan = nullptr;
        CHECK_FDR(fdr_plan_create(&plan, rows, cols, 1, count, 0));
        CHECK_FDR(fdr_plan_set_psf_host(plan, psfc.ptr<float>(0), psfc.rows, psfc.cols, psfc.step, K));
        vector<const float*> in((size_t)count);
        vector<float*> dst((size_t)count);
        for (int i = 0; i < count; ++i) {
            in[(size_t)i] = channels[(size_t)(first + i)].ptr<float>(0);
            dst[(size_t)i] = out[(size_t)(first + i)].ptr<float>(0);
        }
        CHECK_FDR(fdr_restore_planes_host_f32(plan, in.data(), channels[(size_t)first].step, dst.data(),
                                              out[(size_t)first].step, count));
        float t[6];
        CHECK_FDR(fdr_plan_get_profile(plan, t));
        for (int k = 0; k < 6; ++k) total[k] += t[k];
        CHECK_FDR(fdr_plan_destroy(plan));
    };
    // planes of one call share a row stride only if they were allocated alike; fall back to
    // one plane per call otherwise
    bool same_step = true;
    for (int i = 1; i < n; ++i) same_step = same_step && channels[(size_t)i].step == channels[0].step;
    if (per_channel_alloc || !same_step)
        for (int i = 0; i < n; ++i) run(i, 1);
    else
        run(0, n);
    for (int i = 0; i < n; ++i) channels[(size_t)i] = out[(size_t)i];  // fft_gpu.cu:384
    print_profile(title, total);
}

void wienerDeblur_RGB_optimized(vector<Mat>& channels, const Mat& psf, float K) {
    restore(channels, psf, K, false, "FAST (Reuse Memory)");  // title: fft_gpu.cu:387
}

void wienerDeblur_RGB_naive(vector<Mat>& channels, const Mat& psf, float K) {
    restore(channels, psf, K, true, "SLOW (Naive Allocation)");  // title: fft_gpu.cu:510
}

void fft_radix2_kernel(float* data, int n, bool inverse) { CHECK_FDR(fdr_fft_radix2_host(data, n, inverse ? 1 : 0)); }

void dft_naive_kernel(float* data, int n, bool inverse) { CHECK_FDR(fdr_dft_naive_host(data, n, inverse ? 1 : 0)); }

void transform_row_kernel(float* rowPtr, int N, bool inverse) { CHECK_FDR(fdr_transform_rows_host(rowPtr, 1, N, inverse ? 1 : 0)); }

void my_dft2D(Mat& complexMat, bool inverse) {
    CV_Assert(complexMat.type() == CV_32FC2);  // fft_serial.cpp:115
    Mat m = complexMat.isContinuous() ? complexMat : complexMat.clone();
    CHECK_FDR(fdr_dft2d_host(m.ptr<float>(0), m.rows, m.cols, inverse ? 1 : 0));
    if (m.data != complexMat.data)
        for (int r = 0; r < m.rows; ++r) memcpy(complexMat.ptr<unsigned char>(r), m.ptr<unsigned char>(r), (size_t)m.cols * m.elemSize());
}

Mat wienerDeblur_myfft(const Mat& img, const Mat& psf, float K) {
    vector<Mat> one(1, img);
    Mat psfc = psf.isContinuous() ? psf : psf.clone();
    fdr_plan* plan = nullptr;
    CHECK_FDR(fdr_plan_create(&plan, img.rows, img.cols, 1, 1, 0));
    CHECK_FDR(fdr_plan_set_psf_host(plan, psfc.ptr<float>(0), psfc.rows, psfc.cols, psfc.step, K));
    Mat out(img.rows, img.cols, CV_32F);
    const float* in = img.ptr<float>(0);
    float* dst = out.ptr<float>(0);
    CHECK_FDR(fdr_restore_planes_host_f32(plan, &in, img.step, &dst, out.step, 1));
    CHECK_FDR(fdr_plan_destroy(plan));
    return out;
}

}  // namespace fft_gpu
