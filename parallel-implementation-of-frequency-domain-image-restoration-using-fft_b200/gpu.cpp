// gpu.cpp -- `./gpu <img-path> <psf-length> <psf-angle> [out.png]`, the CLI of the gpu mode.
//
// Keeps the reference driver's contract (gpu.cpp:57-138): same positional arguments, usage
// string, "Cannot read image", exit code -1, K = 0.01, the three timed sections with their
// stdout lines and [Speedup] ratios.  Differences: every section runs on a
// fresh copy of the input (the reference filters `channels` in place three times), the serial
// baseline is printed only when the reference's serial translation unit is linked in
// (-DFDR_WITH_REFERENCE_SERIAL), and an optional 4th argument writes the restored 8-bit image
// (after the same Lab white-balance post stage as the reference, gpu.cpp:123-134, run on the device).
#include <cstdlib>
#include <iostream>

#include "fft/fft.hpp"
#include "utils.hpp"

using namespace cv;
using namespace std;
using namespace std::chrono;

int main(int argc, char** argv) {
    if (argc != 4 && argc != 5) {
        cout << "Usage: ./gpu <img-path> <psf-length> <psf-angle>\n";
        return -1;
    }
    string img_path = argv[1];
    int psf_length = atoi(argv[2]);
    double psf_angle = atof(argv[3]);

    Mat img = imread(img_path, IMREAD_COLOR);
    if (img.empty()) {
        cout << "Cannot read image\n";
        return -1;
    }
    img.convertTo(img, CV_32F);
    img /= 255.0;

    Mat psf = motionBlurKernel(psf_length, psf_angle);
    float K = 0.01f;

    vector<Mat> input;
    split(img, input);

    double serial_time = 0.0;
#ifdef FDR_WITH_REFERENCE_SERIAL
    {
        vector<Mat> serial_channels = input;
        auto t0 = high_resolution_clock::now();
        for (int i = 0; i < 3; i++) {
            Mat channel = autoPadToPowerOfTwo(serial_channels[i]);
            serial_channels[i] = fft_serial::wienerDeblur_myfft(channel, psf, K);
            serial_channels[i] = serial_channels[i](Rect(0, 0, img.cols, img.rows));
        }
        auto t1 = high_resolution_clock::now();
        serial_time = getElapsedMs(t0, t1);
        cout << "Deblurring 3 channels took(serial): " << serial_time << " ms\n";
    }
#endif

    int ndev = 0;
    if (fdr_device_count(&ndev) != FDR_OK || ndev < 1) {
        cerr << "Error: " << __FILE__ << ":" << __LINE__ << ", " << fdr_last_error() << endl;
        return 1;
    }

    vector<Mat> channels = input;
    fft_gpu::wienerDeblur_RGB_optimized(channels, psf, K);  // untimed warm-up (gpu.cpp:96)

    channels = input;
    auto t_start = high_resolution_clock::now();
    fft_gpu::wienerDeblur_RGB_optimized(channels, psf, K);
    auto t_end = high_resolution_clock::now();
    double gpu_time = getElapsedMs(t_start, t_end);
    cout << "Deblurring 3 channels took(gpu[optimize]): " << gpu_time << " ms\n";
    if (serial_time > 0) printf("[Speedup] %.2fx ms\n", serial_time / gpu_time);
    vector<Mat> result = channels;

    channels = input;
    t_start = high_resolution_clock::now();
    fft_gpu::wienerDeblur_RGB_naive(channels, psf, K);
    t_end = high_resolution_clock::now();
    gpu_time = getElapsedMs(t_start, t_end);
    cout << "Deblurring 3 channels took(gpu): " << gpu_time << " ms\n";
    if (serial_time > 0) printf("[Speedup] %.2fx ms\n", serial_time / gpu_time);

    if (argc == 5) {
        // gpu.cpp:123-134: Lab white balance against the blurred input, then 8-bit -- on the device
        Mat out8(img.rows, img.cols, CV_8UC3);
        vector<Mat> rc(3), oc(3);
        const float* rp[3];
        const float* op[3];
        for (int i = 0; i < 3; ++i) {
            rc[i] = result[i].isContinuous() ? result[i] : result[i].clone();
            oc[i] = input[i].isContinuous() ? input[i] : input[i].clone();
            rp[i] = rc[i].ptr<float>(0);
            op[i] = oc[i].ptr<float>(0);
        }
        if (fdr_white_balance_pack_host(rp, op, img.rows, img.cols, out8.ptr<unsigned char>(0)) != FDR_OK) {
            cerr << "Error: " << __FILE__ << ":" << __LINE__ << ", " << fdr_last_error() << endl;
            return 1;
        }
        if (!imwrite(argv[4], out8)) {
            cout << "Cannot write image\n";
            return -1;
        }
    }
    waitKey(0);
    return 0;
}
