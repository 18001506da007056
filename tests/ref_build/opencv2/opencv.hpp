// tests/ref_build/opencv2/opencv.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for the slice of OpenCV the reference's visualisation touches, so that the reference's own
// translation units can be compiled HERE, unmodified and in place (OpenCV's C++ headers are not in this
// image):
//   * RP/test/test_recursive_patchwork.cpp and RP/src/main.cpp include "visualization.hpp", whose
//     declarations use cv::Mat, cv::Scalar and cv::Point2i (RP/include/visualization.hpp:16-74);
//   * RP/src/visualization.cpp:18-113 (createBEVImage, createGroundNonGroundImage) is the pixel oracle of
//     rpw_bev_image: it needs cv::Mat(rows, cols, CV_8UC3, Scalar), Mat::at<cv::Vec3b>(y, x), cv::Vec3b;
//   * the rest of that file (imwrite, namedWindow, imshow, waitKey, circle) only has to link; imwrite
//     writes a real (stored, uncompressed) PNG so that the reference CLI's output can be read back.
// Semantics kept: row-major 8-bit BGR storage, saturating Scalar -> uchar fill, at<>() addressing.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#define CV_8UC3 16

namespace cv {

typedef unsigned char uchar;

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

struct Vec3b {
    uchar val[3];
    Vec3b() : val{0, 0, 0} {}
    Vec3b(uchar a, uchar b, uchar c) : val{a, b, c} {}
    uchar& operator[](int i) { return val[i]; }
    const uchar& operator[](int i) const { return val[i]; }
};
static_assert(sizeof(Vec3b) == 3, "Vec3b must be three packed bytes");

struct Point2i {
    int x, y;
    Point2i() : x(0), y(0) {}
    Point2i(int px, int py) : x(px), y(py) {}
};
typedef Point2i Point;

inline uchar saturate_u8(double v) {
    const long r = std::lrint(v);
    return (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type, const Scalar& fill = Scalar()) : rows(r), cols(c), type_(type), buf_((size_t)r * c * 3) {
        const uchar b = saturate_u8(fill[0]), g = saturate_u8(fill[1]), rr = saturate_u8(fill[2]);
        for (size_t i = 0; i < (size_t)r * c; ++i) { buf_[3 * i] = b; buf_[3 * i + 1] = g; buf_[3 * i + 2] = rr; }
    }
    template <typename T> T& at(int y, int x) { return *reinterpret_cast<T*>(&buf_[((size_t)y * cols + x) * sizeof(T)]); }
    template <typename T> const T& at(int y, int x) const { return *reinterpret_cast<const T*>(&buf_[((size_t)y * cols + x) * sizeof(T)]); }
    bool empty() const { return buf_.empty(); }
    int type() const { return type_; }
    uchar* data() { return buf_.data(); }
    const uchar* data() const { return buf_.data(); }

private:
    int type_ = CV_8UC3;
    std::vector<uchar> buf_;
};

enum { WINDOW_AUTOSIZE = 1 };
inline void namedWindow(const std::string&, int = WINDOW_AUTOSIZE) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

// filled disc (thickness < 0) or nothing else the reference asks for
inline void circle(Mat& img, Point2i c, int radius, const Scalar& color, int /*thickness*/ = 1) {
    const Vec3b v(saturate_u8(color[0]), saturate_u8(color[1]), saturate_u8(color[2]));
    for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
            const int x = c.x + dx, y = c.y + dy;
            if (dx * dx + dy * dy <= radius * radius && x >= 0 && x < img.cols && y >= 0 && y < img.rows) img.at<Vec3b>(y, x) = v;
        }
}

// PNG writer: 8-bit RGB, zlib stream of stored (uncompressed) deflate blocks.
namespace standin_png {
inline uint32_t crc32(const uchar* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
inline void put32(std::vector<uchar>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
inline void chunk(std::vector<uchar>& out, const char* tag, const std::vector<uchar>& body) {
    put32(out, (uint32_t)body.size());
    std::vector<uchar> t(tag, tag + 4);
    t.insert(t.end(), body.begin(), body.end());
    out.insert(out.end(), t.begin(), t.end());
    put32(out, crc32(t.data(), t.size()));
}
}  // namespace standin_png

inline bool imwrite(const std::string& filename, const Mat& img) {
    using namespace standin_png;
    if (img.empty()) return false;
    std::vector<uchar> raw;
    raw.reserve((size_t)img.rows * (img.cols * 3 + 1));
    for (int y = 0; y < img.rows; ++y) {
        raw.push_back(0);  // filter: none
        for (int x = 0; x < img.cols; ++x) {
            const Vec3b& p = img.at<Vec3b>(y, x);
            raw.push_back(p[2]); raw.push_back(p[1]); raw.push_back(p[0]);  // BGR -> RGB
        }
    }
    std::vector<uchar> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uchar c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t len = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + len >= raw.size() ? 1 : 0);
        z.push_back(len & 0xFF); z.push_back(len >> 8); z.push_back(~len & 0xFF); z.push_back((~len >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        if (raw.empty()) break;
    }
    put32(z, (b << 16) | a);
    std::vector<uchar> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uchar> ihdr;
    put32(ihdr, (uint32_t)img.cols); put32(ihdr, (uint32_t)img.rows);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

}  // namespace cv
