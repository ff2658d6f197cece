// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal header-only stand-in for <opencv2/opencv.hpp>, written for this repo so
// that the reference's CPU translation units (fft/fft_serial.cpp, fft_openmp.cpp,
// fft_simd.cpp and utils.hpp under /root/reference) compile UNMODIFIED, straight
// from their read-only location, in a container that has no OpenCV C++ SDK.
// Only the cv:: surface those files touch is provided.  Arithmetic follows
// OpenCV's documented element-wise fp32 semantics (one rounding per operation,
// NORM_MINMAX scale/shift derived in double and applied in float).
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_Assert(expr)                                                                   \
    do {                                                                                  \
        if (!(expr)) {                                                                    \
            std::fprintf(stderr, "CV_Assert failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__); \
            std::abort();                                                                 \
        }                                                                                 \
    } while (0)

namespace cv {

typedef unsigned char uchar;

enum { BORDER_CONSTANT = 0 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_L2SQR = 5, NORM_MINMAX = 32 };
enum { INTER_LINEAR = 1 };

struct Vec2f {
    float val[2];
    Vec2f() : val{0.f, 0.f} {}
    Vec2f(float a, float b) : val{a, b} {}
    float& operator[](int i) { return val[i]; }
    const float& operator[](int i) const { return val[i]; }
};

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
    operator Point2f() const { return Point2f((float)x, (float)y); }
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double v0) : val{v0, 0, 0, 0} {}
    static Scalar all(double v) {
        Scalar s;
        s.val[0] = s.val[1] = s.val[2] = s.val[3] = v;
        return s;
    }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

inline int shim_channels(int type) { return (type >> 3) + 1; }
inline int shim_elem_size(int type) {
    int depth = type & 7;
    int bytes = (depth == CV_8U) ? 1 : 4;
    return bytes * shim_channels(type);
}

// Reference-counted dense matrix; copies alias, ROI views alias, clone() deep-copies.
class Mat {
public:
    int rows, cols;
    uchar* data;
    const uchar* datastart;  // whole allocation (fft_mpi.cpp:363 copies [datastart, dataend))
    const uchar* dataend;
    size_t step;  // bytes between consecutive rows

    Mat() : rows(0), cols(0), data(nullptr), datastart(nullptr), dataend(nullptr), step(0), type_(CV_32F) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_ && step == (size_t)c * shim_elem_size(type))
            return;
        rows = r;
        cols = c;
        type_ = type;
        step = (size_t)c * shim_elem_size(type);
        size_t bytes = step * (size_t)r;
        store_ = std::shared_ptr<uchar>(static_cast<uchar*>(std::malloc(bytes ? bytes : 1)), std::free);
        data = store_.get();
        datastart = data;
        dataend = data + bytes;
    }

    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        std::memset(m.data, 0, m.step * (size_t)r);
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

    int type() const { return type_; }
    int channels() const { return shim_channels(type_); }
    size_t elemSize() const { return (size_t)shim_elem_size(type_); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * (size_t)cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step * (size_t)r); }
    template <typename T> const T* ptr(int r = 0) const {
        return reinterpret_cast<const T*>(data + step * (size_t)r);
    }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

    Mat operator()(const Rect& roi) const {
        CV_Assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
        Mat v;
        v.rows = roi.height;
        v.cols = roi.width;
        v.type_ = type_;
        v.step = step;
        v.store_ = store_;
        v.datastart = datastart;
        v.dataend = dataend;
        v.data = data + step * (size_t)roi.y + elemSize() * (size_t)roi.x;
        return v;
    }

    Mat clone() const {
        Mat m(rows, cols, type_);
        size_t rowBytes = (size_t)cols * elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + m.step * (size_t)r, data + step * (size_t)r, rowBytes);
        return m;
    }

    Mat mul(const Mat& o) const;

private:
    int type_;
    std::shared_ptr<uchar> store_;
};

namespace shim_detail {
template <typename F> inline Mat binary_f32(const Mat& a, const Mat& b, F f) {
    CV_Assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat out(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) {
        const float* pa = a.ptr<float>(r);
        const float* pb = b.ptr<float>(r);
        float* po = out.ptr<float>(r);
        for (int c = 0; c < a.cols; ++c) po[c] = f(pa[c], pb[c]);
    }
    return out;
}
template <typename F> inline Mat unary_f32(const Mat& a, F f) {
    CV_Assert(a.type() == CV_32F);
    Mat out(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) {
        const float* pa = a.ptr<float>(r);
        float* po = out.ptr<float>(r);
        for (int c = 0; c < a.cols; ++c) po[c] = f(pa[c]);
    }
    return out;
}
}  // namespace shim_detail

inline Mat Mat::mul(const Mat& o) const {
    return shim_detail::binary_f32(*this, o, [](float x, float y) { return x * y; });
}
inline Mat operator+(const Mat& a, const Mat& b) {
    return shim_detail::binary_f32(a, b, [](float x, float y) { return x + y; });
}
inline Mat operator-(const Mat& a, const Mat& b) {
    return shim_detail::binary_f32(a, b, [](float x, float y) { return x - y; });
}
inline Mat operator/(const Mat& a, const Mat& b) {
    return shim_detail::binary_f32(a, b, [](float x, float y) { return x / y; });
}
inline Mat operator-(const Mat& a) {
    return shim_detail::unary_f32(a, [](float x) { return -x; });
}
inline Mat operator+(const Mat& a, const Scalar& s) {
    const float k = (float)s.val[0];
    return shim_detail::unary_f32(a, [k](float x) { return x + k; });
}
// m /= s : OpenCV evaluates it as m * (1/s) through convertTo (one rounding of the float factor)
inline Mat& operator/=(Mat& a, double s) {
    const float k = (float)(1.0 / s);
    for (int r = 0; r < a.rows; ++r) {
        float* p = a.ptr<float>(r);
        for (int c = 0; c < a.cols * a.channels(); ++c) p[c] = p[c] * k;
    }
    return a;
}
inline Mat operator*(const Mat& a, double s) {
    const float k = (float)s;
    return shim_detail::unary_f32(a, [k](float x) { return x * k; });
}

inline int getOptimalDFTSize(int n) {
    // every reference driver pre-pads to a power of two (serial.cpp:36), where OpenCV's
    // 2^a*3^b*5^c table returns n itself.
    CV_Assert(n > 0 && (n & (n - 1)) == 0);
    return n;
}

inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                           const Scalar& value = Scalar()) {
    CV_Assert(borderType == BORDER_CONSTANT && src.type() == CV_32F);
    CV_Assert(top >= 0 && bottom >= 0 && left >= 0 && right >= 0);
    Mat out(src.rows + top + bottom, src.cols + left + right, CV_32F);
    const float fill = (float)value.val[0];
    for (int r = 0; r < out.rows; ++r) {
        float* po = out.ptr<float>(r);
        for (int c = 0; c < out.cols; ++c) po[c] = fill;
    }
    for (int r = 0; r < src.rows; ++r)
        std::memcpy(out.ptr<float>(r + top) + left, src.ptr<float>(r), sizeof(float) * (size_t)src.cols);
    dst = out;
}

inline void merge(const Mat* mv, size_t count, Mat& dst) {
    CV_Assert(count >= 1);
    const int ch = (int)count;
    Mat out(mv[0].rows, mv[0].cols, CV_32F + ((ch - 1) << 3));
    for (int k = 0; k < ch; ++k) {
        CV_Assert(mv[k].type() == CV_32F && mv[k].rows == out.rows && mv[k].cols == out.cols);
        for (int r = 0; r < out.rows; ++r) {
            const float* ps = mv[k].ptr<float>(r);
            float* po = out.ptr<float>(r);
            for (int c = 0; c < out.cols; ++c) po[(size_t)c * ch + k] = ps[c];
        }
    }
    dst = out;
}
inline void merge(const std::vector<Mat>& mv, Mat& dst) { merge(mv.data(), mv.size(), dst); }

inline void split(const Mat& src, Mat* mv) {
    const int ch = src.channels();
    for (int k = 0; k < ch; ++k) {
        Mat plane(src.rows, src.cols, CV_32F);
        for (int r = 0; r < src.rows; ++r) {
            const float* ps = src.ptr<float>(r);
            float* po = plane.ptr<float>(r);
            for (int c = 0; c < src.cols; ++c) po[c] = ps[(size_t)c * ch + k];
        }
        mv[k] = plane;
    }
}
inline void split(const Mat& src, std::vector<Mat>& mv) {
    mv.resize((size_t)src.channels());
    split(src, mv.data());
}

inline void transpose(const Mat& src, Mat& dst) {
    const size_t es = src.elemSize();
    Mat out(src.cols, src.rows, src.type());
    const int B = 32;
    for (int r0 = 0; r0 < src.rows; r0 += B)
        for (int c0 = 0; c0 < src.cols; c0 += B)
            for (int r = r0; r < std::min(r0 + B, src.rows); ++r)
                for (int c = c0; c < std::min(c0 + B, src.cols); ++c)
                    std::memcpy(out.data + out.step * (size_t)c + es * (size_t)r,
                                src.data + src.step * (size_t)r + es * (size_t)c, es);
    dst = out;
}

inline void magnitude(const Mat& x, const Mat& y, Mat& mag) {
    mag = shim_detail::binary_f32(x, y, [](float a, float b) { return std::sqrt(a * a + b * b); });
}

// NORM_MINMAX only: scale/shift in double, applied per element in float
// (OpenCV: minMaxIdx -> convertTo(dst, type, scale, shift)).  cv2 4.13 applies it as ONE
// fused multiply-add (v_fma on AVX2 hosts; checked bit-for-bit against cv2.normalize by
// tests/golden/make_golden.py), so fmaf is used here rather than a separate mul and add.
inline void normalize(const Mat& src, Mat& dst, double alpha, double beta, int normType) {
    CV_Assert(normType == NORM_MINMAX && src.type() == CV_32F);
    double smin = DBL_MAX, smax = -DBL_MAX;
    for (int r = 0; r < src.rows; ++r) {
        const float* p = src.ptr<float>(r);
        for (int c = 0; c < src.cols; ++c) {
            smin = std::min(smin, (double)p[c]);
            smax = std::max(smax, (double)p[c]);
        }
    }
    const double dmin = std::min(alpha, beta), dmax = std::max(alpha, beta);
    // OpenCV 4.x, rtype == CV_32F: scale is rounded to float BEFORE the shift is derived
    // (scale = (float)scale; shift = (float)dmin - (float)(smin*scale)); checked against
    // cv2.normalize on 300 random planes and tests/golden/normalize_case.npz.
    double scale = (dmax - dmin) * ((smax - smin) > DBL_EPSILON ? 1. / (smax - smin) : 0.);
    scale = (double)(float)scale;
    const double shift = (double)((float)dmin - (float)(smin * scale));
    const float a = (float)scale, b = (float)shift;
    Mat out(src.rows, src.cols, CV_32F);
    for (int r = 0; r < src.rows; ++r) {
        const float* p = src.ptr<float>(r);
        float* po = out.ptr<float>(r);
        for (int c = 0; c < src.cols; ++c) po[c] = std::fmaf(p[c], a, b);
    }
    dst = out;
}

// Declared only: referenced by inline helpers in the reference's utils.hpp that the
// oracle never calls (the PSF and the Lab stage come from Python cv2 instead).
Mat getRotationMatrix2D(Point2f center, double angle, double scale);
void warpAffine(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags = INTER_LINEAR,
                int borderMode = BORDER_CONSTANT, const Scalar& borderValue = Scalar());
Scalar mean(const Mat& src);
void min(const Mat& src1, double s, Mat& dst);
void max(const Mat& src1, double s, Mat& dst);

}  // namespace cv
