// utils.hpp -- helpers with the reference's names (utils.hpp:9-52).  motionBlurKernel builds the
// PSF ON THE DEVICE through the C ABI (fdr_motion_psf_host), bit-identical to the OpenCV
// getRotationMatrix2D + warpAffine recipe of the reference (utils.hpp:15-24).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../include/fdr_b200.h"
#include "fft/fft.hpp"
using namespace std::chrono;

inline double getElapsedMs(high_resolution_clock::time_point start, high_resolution_clock::time_point end) {
    return duration<double, std::milli>(end - start).count();
}

inline Mat motionBlurKernel(int size, double angle) {
    Mat k(size, size, CV_32F);
    if (fdr_motion_psf_host(size, angle, k.ptr<float>(0)) != FDR_OK) {
        std::fprintf(stderr, "Error: %s:%d, %s\n", __FILE__, __LINE__, fdr_last_error());
        std::exit(1);
    }
    return k;
}

inline int nextPowerOfTwo(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

inline bool isPowerOfTwo(int n) { return n > 0 && ((n & (n - 1)) == 0); }

inline Mat autoPadToPowerOfTwo(const Mat& src) {
    const int newRows = nextPowerOfTwo(src.rows), newCols = nextPowerOfTwo(src.cols);
    Mat padded = Mat::zeros(newRows, newCols, src.type());
    for (int r = 0; r < src.rows; ++r) std::memcpy(padded.ptr<unsigned char>(r), src.ptr<unsigned char>(r), (size_t)src.cols * src.elemSize());
    return padded;
}
