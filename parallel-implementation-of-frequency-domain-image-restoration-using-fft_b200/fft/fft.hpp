// fft/fft.hpp -- the operator interface of the gpu mode, same names and signatures as the
// reference's fft/fft.hpp:31-45 (namespace fft_gpu), implemented by fft/fft_gpu.cpp on top of
// the C ABI in include/fdr_b200.h.  With OpenCV installed the real cv::Mat is used, so the
// reference's own gpu.cpp compiles against this header unchanged; otherwise compat/cvmat.hpp
// supplies the few cv:: types the interface needs.
//
// Differences from the reference, all additive:
//   - fft_radix2_kernel, my_dft2D are real implementations (empty stubs in fft_gpu.cu:514-515);
//   - dft_naive_kernel, transform_row_kernel, wienerDeblur_myfft are defined (declared but never
//     defined by the reference, fft.hpp:37,39,44);
//   - results follow the SERIAL path (normalise over the padded plane, then crop), not the
//     reference gpu mode's crop-then-normalise and stale-input ordering.
#pragma once
#if __has_include(<opencv2/opencv.hpp>) && !defined(FDR_FORCE_COMPAT_MAT)
#include <opencv2/opencv.hpp>
#else
#include "../compat/cvmat.hpp"
#endif
#include <complex>
#include <vector>
using namespace cv;
using namespace std;

// NOTE on semantics (not on the signatures, which are the reference's): these functions normalise every channel over the PADDED
// plane and then crop, exactly as the serial path does (fft_serial.cpp:246-258) -- the reference's own gpu mode crops first and
// normalises the cropped plane (fft_gpu.cu:361-381).  For images whose sizes are powers of two the two are identical; otherwise the
// result matches serial.cpp / the --impl serial output, which is the parity contract of this path.
namespace fft_gpu {
    void wienerDeblur_RGB_naive(vector<Mat>& channels, const Mat& psf, float K);
    void wienerDeblur_RGB_optimized(vector<Mat>& channels, const Mat& psf, float K);
    void fft_radix2_kernel(float* data, int n, bool inverse);
    void dft_naive_kernel(float* data, int n, bool inverse);
    void transform_row_kernel(float* rowPtr, int N, bool inverse);
    void my_dft2D(Mat& complexMat, bool inverse);
    inline void my_dft2D_forward(Mat& complexMat) { my_dft2D(complexMat, false); }
    inline void my_dft2D_inverse(Mat& complexMat) { my_dft2D(complexMat, true); }
    // Wiener deblur of one (already padded or not) plane; returns the normalised plane
    Mat wienerDeblur_myfft(const Mat& img, const Mat& psf, float K);
}

// Present only when the reference's serial translation unit is linked in (make SERIAL=1);
// gpu.cpp prints the serial baseline line and [Speedup] from it (gpu.cpp:82-91).
namespace fft_serial {
    Mat wienerDeblur_myfft(const Mat& img, const Mat& psf, float K);
}
