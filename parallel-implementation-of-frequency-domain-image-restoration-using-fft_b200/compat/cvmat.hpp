// compat/cvmat.hpp -- the small part of cv:: that the gpu-mode host code touches, for builds
// on machines without the OpenCV SDK (this image has none).  When <opencv2/opencv.hpp> IS
// available it is used instead (see fft/fft.hpp), so the reference's drivers link against this
// repository's fft_gpu unchanged.
//
// Provided: Mat (CV_8U / CV_32F, 1-4 channels, shared storage, ROI views), Size, Rect, Point,
// Scalar, split, merge, imread / imwrite for 8-bit PNG (zlib), convertTo, operator/=.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr)                                                                        \
    do {                                                                                       \
        if (!(expr)) {                                                                         \
            std::fprintf(stderr, "CV_Assert failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__); \
            std::abort();                                                                      \
        }                                                                                      \
    } while (0)

namespace cv {

typedef unsigned char uchar;
enum { IMREAD_COLOR = 1 };

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Vec2f {
    float val[2];
    float& operator[](int i) { return val[i]; }
    const float& operator[](int i) const { return val[i]; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }

    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        step = (size_t)c * elemSize();
        store_ = std::shared_ptr<uchar>(static_cast<uchar*>(std::malloc(std::max<size_t>(step * r, 1))), std::free);
        data = store_.get();
    }
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        std::memset(m.data, 0, m.step * (size_t)r);
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : 4; }
    size_t elemSize() const { return elemSize1() * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step * (size_t)r); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + step * (size_t)r); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

    Mat operator()(const Rect& roi) const {
        CV_Assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
        Mat v = *this;
        v.rows = roi.height;
        v.cols = roi.width;
        v.data = data + step * (size_t)roi.y + elemSize() * (size_t)roi.x;
        return v;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + m.step * r, data + step * r, (size_t)cols * elemSize());
        return m;
    }
    // dst = saturate(src * alpha + beta) converted to rtype's depth (u8 <-> f32 only)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        const int ddepth = rtype < 0 ? depth() : (rtype & 7);
        Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
        const float a = (float)alpha, b = (float)beta;
        const int n = cols * channels();
        for (int r = 0; r < rows; ++r) {
            for (int i = 0; i < n; ++i) {
                const float s = depth() == CV_8U ? (float)ptr<uchar>(r)[i] : ptr<float>(r)[i];
                const float v = std::fmaf(s, a, b);
                if (ddepth == CV_8U) {
                    long q = std::lrintf(v);
                    out.ptr<uchar>(r)[i] = (uchar)(q < 0 ? 0 : (q > 255 ? 255 : q));
                } else {
                    out.ptr<float>(r)[i] = v;
                }
            }
        }
        dst = out;
    }
    Mat& operator/=(double s) {  // OpenCV evaluates m /= s as m * (1/s)
        Mat t;
        convertTo(t, -1, 1.0 / s, 0.0);
        *this = t;
        return *this;
    }

private:
    int type_ = CV_32F;
    std::shared_ptr<uchar> store_;
};

inline void split(const Mat& src, std::vector<Mat>& mv) {
    const int ch = src.channels();
    mv.assign(ch, Mat());
    for (int k = 0; k < ch; ++k) {
        Mat pl(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
        const size_t es = src.elemSize1();
        for (int r = 0; r < src.rows; ++r)
            for (int c = 0; c < src.cols; ++c)
                std::memcpy(pl.data + pl.step * r + es * c, src.data + src.step * r + es * ((size_t)c * ch + k), es);
        mv[k] = pl;
    }
}
inline void merge(const std::vector<Mat>& mv, Mat& dst) {
    const int ch = (int)mv.size();
    CV_Assert(ch >= 1);
    Mat out(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), ch));
    const size_t es = mv[0].elemSize1();
    for (int k = 0; k < ch; ++k)
        for (int r = 0; r < out.rows; ++r)
            for (int c = 0; c < out.cols; ++c)
                std::memcpy(out.data + out.step * r + es * ((size_t)c * ch + k), mv[k].data + mv[k].step * r + es * c, es);
    dst = out;
}

// ---- minimal PNG I/O (8-bit gray / RGB / RGBA, non-interlaced), BGR channel order like OpenCV ----
namespace pngdetail {
inline uint32_t be32(const uchar* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace pngdetail

inline Mat imread(const std::string& path, int = IMREAD_COLOR) {
    using namespace pngdetail;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return Mat();
    std::vector<uchar> buf;
    uchar tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const uchar sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8) != 0) return Mat();
    uint32_t W = 0, H = 0;
    int bitdepth = 0, ctype = 0, interlace = 0;
    std::vector<uchar> idat;
    size_t pos = 8;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const char* tag = reinterpret_cast<const char*>(&buf[pos + 4]);
        if (pos + 12 + len > buf.size()) return Mat();
        const uchar* d = &buf[pos + 8];
        if (!std::memcmp(tag, "IHDR", 4)) {
            W = be32(d);
            H = be32(d + 4);
            bitdepth = d[8];
            ctype = d[9];
            interlace = d[12];
        } else if (!std::memcmp(tag, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(tag, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!W || !H || bitdepth != 8 || interlace != 0) return Mat();
    const int spp = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!spp) return Mat();
    const size_t stride = (size_t)W * spp;
    std::vector<uchar> raw((stride + 1) * H);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return Mat();
    std::vector<uchar> prev(stride, 0), cur(stride);
    Mat img((int)H, (int)W, CV_8UC3);
    for (uint32_t y = 0; y < H; ++y) {
        const uchar* line = &raw[(stride + 1) * y];
        const int ft = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)spp ? cur[i - spp] : 0, b = prev[i], c = i >= (size_t)spp ? prev[i - spp] : 0;
            int v = line[1 + i];
            switch (ft) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
            }
            cur[i] = (uchar)v;
        }
        uchar* o = img.ptr<uchar>((int)y);
        for (uint32_t x = 0; x < W; ++x) {
            const uchar* px = &cur[(size_t)x * spp];
            const uchar r = px[0], g = spp >= 3 ? px[1] : px[0], b = spp >= 3 ? px[2] : px[0];
            o[3 * x] = b;
            o[3 * x + 1] = g;
            o[3 * x + 2] = r;
        }
        prev.swap(cur);
    }
    return img;
}

inline bool imwrite(const std::string& path, const Mat& img) {
    CV_Assert(img.depth() == CV_8U && (img.channels() == 3 || img.channels() == 1));
    const int spp = img.channels();
    const size_t stride = (size_t)img.cols * spp;
    std::vector<uchar> raw((stride + 1) * img.rows);
    for (int y = 0; y < img.rows; ++y) {
        uchar* line = &raw[(stride + 1) * y];
        line[0] = 0;
        const uchar* s = img.ptr<uchar>(y);
        for (int x = 0; x < img.cols; ++x) {
            if (spp == 3) {
                line[1 + 3 * x] = s[3 * x + 2];
                line[2 + 3 * x] = s[3 * x + 1];
                line[3 + 3 * x] = s[3 * x];
            } else {
                line[1 + x] = s[x];
            }
        }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uchar> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto chunk = [&](const char* tag, const uchar* d, uint32_t len) {
        uchar hdr[8] = {(uchar)(len >> 24), (uchar)(len >> 16), (uchar)(len >> 8), (uchar)len, (uchar)tag[0], (uchar)tag[1], (uchar)tag[2], (uchar)tag[3]};
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(d, 1, len, f);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, d, len);
        uchar c4[4] = {(uchar)(crc >> 24), (uchar)(crc >> 16), (uchar)(crc >> 8), (uchar)crc};
        std::fwrite(c4, 1, 4, f);
    };
    static const uchar sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(sig, 1, 8, f);
    uchar ihdr[13] = {(uchar)(img.cols >> 24), (uchar)(img.cols >> 16), (uchar)(img.cols >> 8), (uchar)img.cols,
                      (uchar)(img.rows >> 24), (uchar)(img.rows >> 16), (uchar)(img.rows >> 8), (uchar)img.rows,
                      8, (uchar)(spp == 3 ? 2 : 0), 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    std::fclose(f);
    return true;
}

inline int waitKey(int = 0) { return -1; }  // headless: gpu.cpp:137 calls waitKey(0) with imshow commented out

}  // namespace cv
