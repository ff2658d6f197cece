// Exercises every function the reference's fft/fft.hpp declares in namespace fft_gpu (fft.hpp:31-45)
// through this repository's C++ host layer, and dumps the results for tests/test_gpu_cpp_api.py.
#include <cstdio>
#include <vector>

#include "fft/fft.hpp"
#include "utils.hpp"

static void dump(FILE* f, const float* p, size_t n) { fwrite(p, sizeof(float), n, f); }

int main(int argc, char** argv) {
    if (argc != 2) return 2;
    FILE* f = fopen(argv[1], "wb");
    if (!f) return 3;
    const int N = 64, R = 16, C = 32;
    std::vector<float> a(2 * N), b(2 * 12), m(2 * R * C);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    for (auto& v : m) v = rnd();
    dump(f, a.data(), a.size());
    dump(f, b.data(), b.size());
    dump(f, m.data(), m.size());
    std::vector<float> t = a;
    fft_gpu::fft_radix2_kernel(t.data(), N, false);       // fft.hpp:35
    dump(f, t.data(), t.size());
    t = a;
    fft_gpu::transform_row_kernel(t.data(), N, true);     // fft.hpp:39 (power of two -> radix 2, inverse)
    dump(f, t.data(), t.size());
    t = b;
    fft_gpu::dft_naive_kernel(t.data(), 12, false);       // fft.hpp:37
    dump(f, t.data(), t.size());
    t = b;
    fft_gpu::transform_row_kernel(t.data(), 12, false);   // non power of two -> naive DFT
    dump(f, t.data(), t.size());
    Mat cm(R, C, CV_32FC2);
    memcpy(cm.data, m.data(), m.size() * sizeof(float));
    fft_gpu::my_dft2D_forward(cm);                        // fft.hpp:40-41
    dump(f, cm.ptr<float>(0), m.size());
    fft_gpu::my_dft2D_inverse(cm);                        // fft.hpp:42
    dump(f, cm.ptr<float>(0), m.size());
    // wienerDeblur_myfft on one plane, PSF from motionBlurKernel (built on the device)
    const int H = 40, W = 56;
    Mat img(H, W, CV_32F);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) img.at<float>(y, x) = rnd() + 0.5f;
    Mat psf = motionBlurKernel(9, 30.0);                  // utils.hpp:15-24
    dump(f, img.ptr<float>(0), (size_t)H * W);
    dump(f, psf.ptr<float>(0), 81);
    Mat out = fft_gpu::wienerDeblur_myfft(img, psf, 0.01f);  // fft.hpp:44
    dump(f, out.ptr<float>(0), (size_t)H * W);
    // the vector<Mat> entry points, including a non-continuous ROI channel (fft_gpu.cu:347-348)
    Mat big = Mat::zeros(H + 3, W + 5, CV_32F);
    Mat roi = big(Rect(2, 1, W, H));
    for (int y = 0; y < H; ++y) memcpy(roi.ptr<float>(y), img.ptr<float>(y), W * sizeof(float));
    std::vector<Mat> ch = {img.clone(), roi, img.clone()};
    fft_gpu::wienerDeblur_RGB_optimized(ch, psf, 0.01f);  // fft.hpp:33
    for (auto& c2 : ch) dump(f, c2.ptr<float>(0), (size_t)H * W);
    std::vector<Mat> ch2 = {img.clone(), img.clone()};
    fft_gpu::wienerDeblur_RGB_naive(ch2, psf, 0.01f);     // fft.hpp:32
    for (auto& c2 : ch2) dump(f, c2.ptr<float>(0), (size_t)H * W);
    fclose(f);
    return 0;
}
