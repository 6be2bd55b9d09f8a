// oracle/ref_shim/opencv2/core.hpp — TEST INFRASTRUCTURE ONLY.
//
// The reference (careylab/LocoMouse_cpp) cannot be built in this image: it needs the OpenCV C++ SDK.  Two
// pieces of its detection path, however, use OpenCV only for VALUE TYPES: Candidates/Candidates.cpp, the free
// functions nmsMax / peakClustering / vecmovingaverage (LocoMouse_Core/LocoMouse_class.cpp:1559-1905), the
// firstLastOverT template (LocoMouse_class.hpp:411-442) and LocoMouse::imadjust (3204-3242, which only fills a
// 256-entry table and applies cv::LUT).  This header supplies
// just those types (Point_, Size_, Rect_, a header-only Mat view, saturate_cast, CV_Assert and inert
// FileStorage stubs) with OpenCV's documented semantics, so that oracle/Makefile can compile the reference's
// OWN source lines, from where they lie under /root/reference, into oracle/_ref/libref_nms.so.  That library
// pins the oracle's restatement of those loops against the real reference code (tests/test_oracle_vs_reference.py).
// No algorithm lives here.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <sys/types.h>
#include <vector>

typedef unsigned char uchar;

#define CV_Assert(expr)                                                        \
    do {                                                                       \
        if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr);     \
    } while (0)
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {

// saturate_cast<int>(double) == cvRound: round half to even
template <typename T>
inline T saturate_cast(double v) { return (T)v; }
template <>
inline int saturate_cast<int>(double v) { return (int)std::lrint(v); }
template <typename T>
inline T saturate_cast(float v) { return saturate_cast<T>((double)v); }
template <typename T>
inline T saturate_cast(int v) { return (T)v; }

template <typename T>
class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U>
    operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
    Point_ &operator+=(const Point_ &o) { x += o.x; y += o.y; return *this; }
    Point_ &operator-=(const Point_ &o) { x -= o.x; y -= o.y; return *this; }
};
template <typename T>
inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T>
inline Point_<T> operator*(const Point_<T> &a, double b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T>
inline Point_<T> operator/(const Point_<T> &a, double b) { return Point_<T>(saturate_cast<T>(a.x / b), saturate_cast<T>(a.y / b)); }
template <typename T>
inline std::ostream &operator<<(std::ostream &o, const Point_<T> &p) { return o << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point;

template <typename T>
class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;

template <typename T>
class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    T area() const { return width * height; }
    Rect_ &operator+=(const Point_<T> &p) { x += p.x; y += p.y; return *this; }
    Rect_ &operator-=(const Point_<T> &p) { x -= p.x; y -= p.y; return *this; }
};
// intersection; empty -> Rect() (all zeros), as cv::Rect_::operator&=
template <typename T>
inline Rect_<T> operator&(const Rect_<T> &a, const Rect_<T> &b) {
    const T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    const T w = std::min(a.x + a.width, b.x + b.width) - x1, h = std::min(a.y + a.height, b.y + b.height) - y1;
    if (w <= 0 || h <= 0) return Rect_<T>();
    return Rect_<T>(x1, y1, w, h);
}
typedef Rect_<int> Rect;

// row-major matrix header: a view on caller memory or an owned buffer; only what nmsMax / peakClustering /
// firstLastOverT / imadjust touch (data, rows, cols, type(), ptr<T>())
class Mat {
    std::vector<uchar> own_;
    int type_;

public:
    uchar *data;
    int rows, cols;
    size_t step;
    Mat() : type_(0), data(nullptr), rows(0), cols(0), step(0) {}
    Mat(int r, int c, int type, void *d, size_t step_bytes) : type_(type), data((uchar *)d), rows(r), cols(c), step(step_bytes) {}
    Mat(int r, int c, int type) : type_(type), rows(r), cols(c) {
        const size_t esz = (type == 0) ? 1 : 4;  // CV_8U : CV_32F / CV_32S
        step = (size_t)c * esz;
        own_.assign((size_t)r * step, 0);
        data = own_.data();
    }
    Mat(const Mat &o) : own_(o.own_), type_(o.type_), data(o.own_.empty() ? o.data : own_.data()), rows(o.rows), cols(o.cols), step(o.step) {}
    Mat &operator=(const Mat &o) {
        own_ = o.own_;
        type_ = o.type_;
        rows = o.rows;
        cols = o.cols;
        step = o.step;
        data = o.own_.empty() ? o.data : own_.data();
        return *this;
    }
    int type() const { return type_; }
    template <typename T>
    const T *ptr(int r) const { return (const T *)(data + (size_t)r * step); }
    template <typename T>
    T *ptr(int r) { return (T *)(data + (size_t)r * step); }
};
// cv::LUT for single-channel 8-bit data: dst(i) = lut(src(i))
inline void LUT(const Mat &src, const Mat &lut, Mat &dst) {
    Mat out(src.rows, src.cols, 0);
    const uchar *t = lut.ptr<uchar>(0);
    for (int r = 0; r < src.rows; ++r) {
        const uchar *s = src.ptr<uchar>(r);
        uchar *d = out.ptr<uchar>(r);
        for (int c = 0; c < src.cols; ++c) d[c] = t[s[c]];
    }
    dst = out;
}

// inert persistence stubs: Candidates.cpp's YAML (de)serialisers must compile, they are never called
class FileNode;
class FileNodeIterator {
public:
    FileNode operator*() const;
    FileNodeIterator &operator++() { return *this; }
    bool operator!=(const FileNodeIterator &) const { return false; }
};
class FileNode {
public:
    FileNode operator[](const char *) const { return FileNode(); }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    bool empty() const { return true; }
    FileNodeIterator begin() const { return FileNodeIterator(); }
    FileNodeIterator end() const { return FileNodeIterator(); }
};
inline FileNode FileNodeIterator::operator*() const { return FileNode(); }
class FileStorage {};
template <typename T>
inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }
template <typename T>
inline void operator>>(const FileNode &, T &) {}

}  // namespace cv
