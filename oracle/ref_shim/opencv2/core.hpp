// oracle/ref_shim/opencv2/core.hpp — TEST INFRASTRUCTURE ONLY.
//
// The reference (careylab/LocoMouse_cpp) cannot be built in this image: it needs the OpenCV C++ SDK.  Two
// pieces of its detection path, however, use OpenCV only for VALUE TYPES: Candidates/Candidates.cpp, the free
// functions nmsMax / peakClustering / vecmovingaverage (LocoMouse_Core/LocoMouse_class.cpp:1559-1905), the
// firstLastOverT template (LocoMouse_class.hpp:411-442) and LocoMouse::imadjust (3204-3242, which only fills a
// 256-entry table and applies cv::LUT).  This header supplies
// just those types (Point_, Size_, Rect_, a header-only Mat with OpenCV's sharing semantics, saturate_cast,
// CV_Assert, inert FileStorage stubs) plus the ELEMENTWISE primitives the compiled code calls (compare, normalize(MINMAX) of a
// mask, convertTo, the alpha*A+beta expression, reduce(SUM / MAX), saturating subtract, threshold (binary / inverse), sum,
// LUT, repeat, >, &, setTo, moments), each a few lines with OpenCV's documented semantics and each pinned against the real
// OpenCV (cv2) in tests/test_oracle_vs_cv2.py (test_reference_shim_primitives_*).  OpenCV ALGORITHMS are not implemented
// here: connectedComponentsWithStats, the image normalisation of readFrame and flip call back into the real OpenCV, and
// filter2D hands out injected maps (the correlation is pinned against cv2.filter2D directly).  With it oracle/Makefile
// compiles the reference's OWN source lines, from where they lie under /root/reference, into oracle/_ref/libref_nms.so,
// which pins the oracle's restatement -- control flow in particular -- against the real reference code
// (tests/test_oracle_vs_reference.py, tests/test_cost_builders.py).
// No reference algorithm (NMS, clustering, pairing, velocity criterion, tail segmentation, cost builders) lives here.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <memory>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <sys/types.h>
#include <vector>

typedef unsigned char uchar;

#define CV_Assert(expr)                                                        \
    do {                                                                       \
        if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr);     \
    } while (0)
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
class Rect_;
template <typename T>
class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    T dot(const Point_ &p) const { return saturate_cast<T>(x * p.x + y * p.y); }   // cv::Point_::dot
    bool inside(const Rect_<T> &r) const;                                            // cv::Rect_::contains: half-open
    template <typename U>
    operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
    Point_ &operator+=(const Point_ &o) { x += o.x; y += o.y; return *this; }
    Point_ &operator-=(const Point_ &o) { x -= o.x; y -= o.y; return *this; }
};
template <typename T>
inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T>
inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y)); }
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
inline bool operator==(const Size_<T> &a, const Size_<T> &b) { return a.width == b.width && a.height == b.height; }
template <typename T>
inline std::ostream &operator<<(std::ostream &o, const Size_<T> &s) { return o << "[" << s.width << " x " << s.height << "]"; }

template <typename T>
class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    T area() const { return width * height; }
    Size_<T> size() const { return Size_<T>(width, height); }
    Rect_ &operator+=(const Point_<T> &p) { x += p.x; y += p.y; return *this; }
    Rect_ &operator-=(const Point_<T> &p) { x -= p.x; y -= p.y; return *this; }
};
template <typename T>
inline bool Point_<T>::inside(const Rect_<T> &r) const { return r.x <= x && x < r.x + r.width && r.y <= y && y < r.y + r.height; }
// intersection; empty -> Rect() (all zeros), as cv::Rect_::operator&=
template <typename T>
inline Rect_<T> operator&(const Rect_<T> &a, const Rect_<T> &b) {
    const T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    const T w = std::min(a.x + a.width, b.x + b.width) - x1, h = std::min(a.y + a.height, b.y + b.height) - y1;
    if (w <= 0 || h <= 0) return Rect_<T>();
    return Rect_<T>(x1, y1, w, h);
}
typedef Rect_<int> Rect;
template <typename T>
inline std::ostream &operator<<(std::ostream &o, const Rect_<T> &r) { return o << "[" << r.width << " x " << r.height << " from (" << r.x << ", " << r.y << ")]"; }

// Row-major single-channel matrix with OpenCV's sharing semantics (copies and ROIs alias the same buffer): what
// nmsMax / peakClustering / firstLastOverT / imadjust / xDist / matchingWithVelocityConstraint / matchViews /
// checkVelCriterion touch.  Depth codes as in OpenCV.
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32SC1 4
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_16U 2
#define CV_16UC1 2
enum { NORM_MINMAX = 32, THRESH_BINARY = 0, THRESH_BINARY_INV = 1, CV_REDUCE_SUM = 0, CV_REDUCE_MAX = 2, CMP_EQ = 0, CC_STAT_AREA = 4, BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };

class Mat;
// the one lazy expression the path relies on: alpha * A + beta, folded like cv::MatOp_AddEx
struct MatExpr {
    const Mat *a;
    double alpha, beta;
};

class Mat {
    std::shared_ptr<std::vector<uchar> > buf_;
    int type_;

public:
    uchar *data;
    int rows, cols;
    size_t step;
    Mat() : type_(0), data(nullptr), rows(0), cols(0), step(0) {}
    Mat(int r, int c, int type, void *d, size_t step_bytes) : type_(type), data((uchar *)d), rows(r), cols(c), step(step_bytes) {}
    Mat(int r, int c, int type) : type_(type), rows(r), cols(c) {
        step = (size_t)c * elemSize();
        buf_.reset(new std::vector<uchar>((size_t)r * step + 1, 0));
        data = buf_->data();
    }
    static size_t elemSizeOf(int type) { return type == CV_8U ? 1 : (type == CV_16U ? 2 : (type == CV_64F ? 8 : 4)); }
    size_t elemSize() const { return elemSizeOf(type_); }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size_<int> size() const { return Size_<int>(cols, rows); }
    template <typename T>
    const T *ptr(int r) const { return (const T *)(data + (size_t)r * step); }
    template <typename T>
    T *ptr(int r) { return (T *)(data + (size_t)r * step); }
    Mat &assign_scaled_u8(const MatExpr &e);
    Mat colRange(int a, int b) const { return (*this)(Rect_<int>(a, 0, b - a, rows)); }
    Mat rowRange(int a, int b) const { return (*this)(Rect_<int>(0, a, cols, b - a)); }
    Mat &setTo(double v) {  // in place, all elements (8-bit)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) ptr<uchar>(r)[c] = (uchar)v;
        return *this;
    }
    void copyTo(Mat &dst) const {
        Mat out;
        convertTo(out, type_);
        dst = out;
    }
    Mat clone() const {
        Mat out;
        convertTo(out, type_);
        return out;
    }
    Mat reshape(int /*cn*/, int /*rows*/) const { return *this; }  // the reference discards the result (class.cpp:1352)
    template <typename T>
    T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T>
    const T &at(int r, int c) const { return ptr<T>(r)[c]; }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size_<int> sz, int type) { return Mat(sz.height, sz.width, type); }
    static Mat ones(int r, int c, int type) {  // only the depths the path uses
        Mat m(r, c, type);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) {
                if (type == CV_32S) m.ptr<int>(i)[j] = 1;
                else if (type == CV_8U) m.ptr<uchar>(i)[j] = 1;
                else if (type == CV_32F) m.ptr<float>(i)[j] = 1.f;
                else m.ptr<double>(i)[j] = 1.0;
            }
        return m;
    }
    Mat operator()(const Rect_<int> &r) const {  // ROI view; out-of-range rectangles are an error as in OpenCV
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw std::runtime_error("cv::Mat ROI out of range");
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    // dst = saturate_cast<dst type>(src * alpha + beta); only the conversions the path uses
    void convertTo(Mat &dst, int type, double alpha = 1.0, double beta = 0.0) const {
        Mat out(rows, cols, type);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                double v = type_ == CV_8U ? (double)ptr<uchar>(r)[c] : type_ == CV_32S ? (double)ptr<int>(r)[c]
                         : type_ == CV_32F ? (double)ptr<float>(r)[c] : ptr<double>(r)[c];
                v = v * alpha + beta;
                if (type == CV_64F) out.ptr<double>(r)[c] = v;
                else if (type == CV_32F) out.ptr<float>(r)[c] = (float)v;
                else if (type == CV_32S) out.ptr<int>(r)[c] = (int)std::lrint(v);
                else { long q = std::lrint(v); out.ptr<uchar>(r)[c] = (uchar)(q < 0 ? 0 : (q > 255 ? 255 : q)); }
            }
        dst = out;
    }
    Mat &setTo(double v, const Mat &mask) {  // in place (ROI views write through to the parent), where mask != 0
        if (mask.rows != rows || mask.cols != cols) throw std::runtime_error("cv::Mat::setTo: mask size mismatch");
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                if (mask.ptr<uchar>(r)[c]) {
                    if (type_ == CV_32F) ptr<float>(r)[c] = (float)v;
                    else ptr<uchar>(r)[c] = (uchar)v;
                }
        return *this;
    }
    Mat &operator=(const MatExpr &e) {
        if (e.a->type() == CV_8U) return assign_scaled_u8(e);  // 8-bit -> 8-bit with scale: executed by the real OpenCV
        Mat out;
        e.a->convertTo(out, e.a->type(), e.alpha, e.beta);
        *this = out;
        return *this;
    }
};
inline MatExpr operator/(const Mat &a, double s) { MatExpr e = {&a, 1.0 / s, 0.0}; return e; }
// (A - s) and (expr / s), folded like cv::MatOp_AddEx (matop.cpp: subtract -> beta -= s; divide -> multiply(e, 1. / s))
inline MatExpr operator-(const Mat &a, double s) { MatExpr e = {&a, 1.0, -s}; return e; }
inline MatExpr operator/(const MatExpr &x, double s) { MatExpr e = {x.a, x.alpha * (1.0 / s), x.beta * (1.0 / s)}; return e; }
inline MatExpr operator-(double c, const MatExpr &e) { MatExpr r = {e.a, -e.alpha, c - e.beta}; return r; }
inline MatExpr operator-(int c, const MatExpr &e) { return (double)c - e; }
// A <= s  ->  8-bit mask, 255 where true
inline Mat operator<=(const Mat &a, double s) {
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) out.ptr<uchar>(r)[c] = ((double)a.ptr<int>(r)[c] <= s) ? 255 : 0;
    return out;
}
// cv::normalize(src, dst, alpha, beta, NORM_MINMAX, -1) on 8-bit data: scale = (beta - alpha) / (max - min), 0 when the
// image is constant (so a constant image becomes all alpha: SURVEY Q7)
inline void shim_u8_op(int op, const Mat &src, Mat &dst);
inline void normalize(const Mat &src, Mat &dst, double alpha, double beta, int /*NORM_MINMAX*/, int /*dtype*/) {
    if (beta == 255) {  // readFrame's image normalisation: executed by the real OpenCV
        shim_u8_op(0, src, dst);
        return;
    }
    double mn = 1e300, mx = -1e300;
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            mn = std::min(mn, (double)src.ptr<uchar>(r)[c]);
            mx = std::max(mx, (double)src.ptr<uchar>(r)[c]);
        }
    const double scale = (beta - alpha) * ((mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0);
    const double shift = alpha - mn * scale;
    Mat out;
    src.convertTo(out, CV_8U, scale, shift);
    dst = out;
}
inline Mat &shim_last_reduce_i32() { static Mat m; return m; }
// cv::reduce(src, dst, dim, CV_REDUCE_SUM, CV_32FC1) for 8-bit sources: dim 0 -> one row, dim 1 -> one column
inline void reduce(const Mat &src, Mat &dst, int dim, int rtype, int dtype = -1) {
    if (rtype == CV_REDUCE_MAX) {  // 8-bit maximum, same depth (dtype -1): only dim 0 is used
        Mat out(dim == 0 ? 1 : src.rows, dim == 0 ? src.cols : 1, CV_8U);
        for (int r = 0; r < src.rows; ++r)
            for (int c = 0; c < src.cols; ++c) {
                uchar &o = dim == 0 ? out.ptr<uchar>(0)[c] : out.ptr<uchar>(r)[0];
                o = std::max(o, src.ptr<uchar>(r)[c]);
            }
        dst = out;
        return;
    }
    if (dtype == CV_32SC1) {  // 8-bit -> 32-bit signed sums (computeMouseBox, LocoMouse_class.cpp:964-968; pinned vs cv2.reduce)
        Mat out(dim == 0 ? 1 : src.rows, dim == 0 ? src.cols : 1, CV_32S);
        for (int r = 0; r < src.rows; ++r)
            for (int c = 0; c < src.cols; ++c) {
                if (dim == 0) out.ptr<int>(0)[c] += (int)src.ptr<uchar>(r)[c];
                else out.ptr<int>(r)[0] += (int)src.ptr<uchar>(r)[c];
            }
        dst = out;
        shim_last_reduce_i32() = out;  // test tap: the harness reads back the sums the reference's code formed
        return;
    }
    Mat out(dim == 0 ? 1 : src.rows, dim == 0 ? src.cols : 1, CV_32F);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            if (dim == 0) out.ptr<float>(0)[c] += (float)src.ptr<uchar>(r)[c];
            else out.ptr<float>(r)[0] += (float)src.ptr<uchar>(r)[c];
        }
    dst = out;
}
struct NoArray {};
inline NoArray noArray() { return NoArray(); }
inline void subtract(const Mat &a, const Mat &b, Mat &dst, const NoArray &, int /*CV_8UC1*/) {  // saturating 8-bit a - b
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            const int d = (int)a.ptr<uchar>(r)[c] - (int)b.ptr<uchar>(r)[c];
            out.ptr<uchar>(r)[c] = (uchar)(d < 0 ? 0 : d);
        }
    dst = out;
}
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type) {  // THRESH_BINARY / THRESH_BINARY_INV
    // cv::threshold(I, I, ...) works in place: views of the same buffer (computeMouseBox's I_side_view / I_bottom_view) must
    // see the result, so an output that already has the source's size and type is written through
    const bool inplace = dst.data && dst.rows == src.rows && dst.cols == src.cols && dst.type() == src.type();
    Mat out(src.rows, src.cols, src.type());  // same depth as the source: 8-bit or 32-bit float
    const bool inv = type == THRESH_BINARY_INV;
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            if (src.type() == CV_32F) out.ptr<float>(r)[c] = ((src.ptr<float>(r)[c] > (float)thresh) != inv) ? (float)maxval : 0.f;
            else out.ptr<uchar>(r)[c] = (((double)src.ptr<uchar>(r)[c] > thresh) != inv) ? (uchar)maxval : 0;
        }
    if (inplace) {
        for (int r = 0; r < src.rows; ++r)
            std::memcpy(dst.ptr<uchar>(r), out.ptr<uchar>(r), (size_t)src.cols * src.elemSize());
    } else {
        dst = out;
    }
    return thresh;
}
// ---- additions for the tail code (detectLineCandidates / selectLargestRegion, LocoMouse_class.cpp:2558-2767) -----------
inline Mat operator-(const Mat &a) {  // -Mat::ones(...): CV_32S only
    Mat out(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) out.ptr<int>(r)[c] = -a.ptr<int>(r)[c];
    return out;
}
inline Mat operator>(const Mat &a, double s) {  // CV_32F > s  ->  8-bit mask, 255 where true
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) out.ptr<uchar>(r)[c] = ((double)a.ptr<float>(r)[c] > s) ? 255 : 0;
    return out;
}
inline Mat operator&(const Mat &a, const Mat &b) {  // bitwise and of two 8-bit images
    if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("cv::bitwise_and: size mismatch");
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) out.ptr<uchar>(r)[c] = a.ptr<uchar>(r)[c] & b.ptr<uchar>(r)[c];
    return out;
}
inline Mat repeat(const Mat &src, int ny, int nx) {  // 8-bit
    Mat out(src.rows * ny, src.cols * nx, CV_8U);
    for (int r = 0; r < out.rows; ++r)
        for (int c = 0; c < out.cols; ++c) out.ptr<uchar>(r)[c] = src.ptr<uchar>(r % src.rows)[c % src.cols];
    return out;
}
inline void compare(const Mat &src, int value, Mat &dst, int /*CMP_EQ*/) {  // CV_16U labels == value -> 255 / 0
    Mat out(src.rows, src.cols, CV_8U);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) out.ptr<uchar>(r)[c] = ((int)src.ptr<unsigned short>(r)[c] == value) ? 255 : 0;
    dst = out;
}
struct Moments {
    double m00, m10, m01;
    Moments() : m00(0), m10(0), m01(0) {}
};
inline Moments moments(const Mat &img, bool binary) {  // raw spatial moments up to order 1 of an 8-bit image
    Moments M;
    for (int r = 0; r < img.rows; ++r)
        for (int c = 0; c < img.cols; ++c) {
            const double v = binary ? (img.ptr<uchar>(r)[c] != 0 ? 1.0 : 0.0) : (double)img.ptr<uchar>(r)[c];
            M.m00 += v;
            M.m10 += v * c;
            M.m01 += v * r;
        }
    return M;
}
// cv::connectedComponentsWithStats and cv::filter2D are ALGORITHMS: the shim does not implement them.  The test harness
// installs callbacks: the labelling runs in the REAL OpenCV (cv2, through ctypes), the filter outputs are injected.
typedef int (*shim_cc_fn)(const uchar *img, int rows, int cols, int connectivity, unsigned short *labels, int *areas, int cap);
typedef void (*shim_filter_fn)(float *dst, int rows, int cols, int id);  // id: kernel(0, 0) of a tagged 1 x 1 kernel, else -1
inline shim_cc_fn &shim_cc_callback() { static shim_cc_fn f = nullptr; return f; }
inline shim_filter_fn &shim_filter_callback() { static shim_filter_fn f = nullptr; return f; }
inline int connectedComponentsWithStats(const Mat &img, Mat &labels, Mat &stats, Mat & /*centroids*/, int connectivity, int /*CV_16U*/) {
    if (!shim_cc_callback()) throw std::runtime_error("connectedComponentsWithStats: no callback installed");
    std::vector<uchar> flat((size_t)img.rows * img.cols);
    for (int r = 0; r < img.rows; ++r)
        for (int c = 0; c < img.cols; ++c) flat[(size_t)r * img.cols + c] = img.ptr<uchar>(r)[c];
    std::vector<unsigned short> lab(flat.size());
    std::vector<int> areas(65536);
    const int n = shim_cc_callback()(flat.data(), img.rows, img.cols, connectivity, lab.data(), areas.data(), (int)areas.size());
    // one spare zero row: largestBWAreaObject reads stats row 1 before it knows that there is a foreground label
    // (LocoMouse_class.cpp:932); in the real OpenCV that read is out of bounds and its value is never used
    Mat L(img.rows, img.cols, CV_16U), S((n > 0 ? n : 1) + 1, 5, CV_32S);
    for (int r = 0; r < img.rows; ++r)
        for (int c = 0; c < img.cols; ++c) L.ptr<unsigned short>(r)[c] = lab[(size_t)r * img.cols + c];
    for (int i = 0; i < n; ++i) S.ptr<int>(i)[CC_STAT_AREA] = areas[i];
    labels = L;
    stats = S;
    return n;
}
// 8-bit -> 8-bit filter2D with a float kernel (LocoMouse_TM::computeMouseBox_DD's disk filter, LocoMouse_TM.cpp:216) and
// cv::floodFill (imfill, LocoMouse_TM.cpp:259) are ALGORITHMS too: both run in the REAL OpenCV through callbacks
// (cv2.filter2D(src, CV_8U, kernel, anchor (-1,-1), delta 0, BORDER_REPLICATE); cv2.floodFill(img, None, (x, y), value)).
typedef void (*shim_filter_u8_fn)(const uchar *src, uchar *dst, int rows, int cols, const float *kernel, int krows, int kcols);
typedef void (*shim_flood_fn)(uchar *img, int rows, int cols, int x, int y, int value);
inline shim_filter_u8_fn &shim_filter_u8_callback() { static shim_filter_u8_fn f = nullptr; return f; }
inline shim_flood_fn &shim_flood_callback() { static shim_flood_fn f = nullptr; return f; }
inline void floodFill(Mat &img, Point_<int> seed, double value) {
    if (!shim_flood_callback()) throw std::runtime_error("floodFill: no callback installed");
    std::vector<uchar> flat((size_t)img.rows * img.cols);
    for (int r = 0; r < img.rows; ++r) std::memcpy(&flat[(size_t)r * img.cols], img.ptr<uchar>(r), (size_t)img.cols);
    shim_flood_callback()(flat.data(), img.rows, img.cols, seed.x, seed.y, (int)value);
    for (int r = 0; r < img.rows; ++r) std::memcpy(img.ptr<uchar>(r), &flat[(size_t)r * img.cols], (size_t)img.cols);
}
inline void bitwise_not(const Mat &src, Mat &dst) {  // 8-bit
    Mat out(src.rows, src.cols, CV_8U);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) out.ptr<uchar>(r)[c] = (uchar)~src.ptr<uchar>(r)[c];
    dst = out;
}
inline Mat operator|(const Mat &a, const Mat &b) {  // bitwise or of two 8-bit images
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) out.ptr<uchar>(r)[c] = a.ptr<uchar>(r)[c] | b.ptr<uchar>(r)[c];
    return out;
}
inline Mat &operator|=(Mat &a, const Mat &b) {  // in place
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) a.ptr<uchar>(r)[c] |= b.ptr<uchar>(r)[c];
    return a;
}
inline void filter2D(const Mat &src, Mat &dst, int ddepth, const Mat &k, Point_<int> /*anchor*/, double /*delta*/, int /*border*/) {
    if (ddepth == CV_8U && src.type() == CV_8U && !k.empty() && (k.rows > 1 || k.cols > 1)) {
        if (!shim_filter_u8_callback()) throw std::runtime_error("filter2D (8-bit): no callback installed");
        std::vector<uchar> in((size_t)src.rows * src.cols), o8(in.size());
        for (int r = 0; r < src.rows; ++r) std::memcpy(&in[(size_t)r * src.cols], src.ptr<uchar>(r), (size_t)src.cols);
        std::vector<float> kk((size_t)k.rows * k.cols);
        for (int r = 0; r < k.rows; ++r)
            for (int c = 0; c < k.cols; ++c) kk[(size_t)r * k.cols + c] = k.type() == CV_64F ? (float)k.ptr<double>(r)[c] : k.ptr<float>(r)[c];
        shim_filter_u8_callback()(in.data(), o8.data(), src.rows, src.cols, kk.data(), k.rows, k.cols);
        Mat out(src.rows, src.cols, CV_8U);
        for (int r = 0; r < src.rows; ++r) std::memcpy(out.ptr<uchar>(r), &o8[(size_t)r * src.cols], (size_t)src.cols);
        dst = out;
        return;
    }
    if (!shim_filter_callback()) throw std::runtime_error("filter2D: no callback installed");
    Mat out(src.rows, src.cols, CV_32F);
    std::vector<float> flat((size_t)src.rows * src.cols);
    shim_filter_callback()(flat.data(), src.rows, src.cols, k.empty() ? -1 : (int)k.ptr<float>(0)[0]);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) out.ptr<float>(r)[c] = flat[(size_t)r * src.cols + c];
    dst = out;
}

// cv::medianBlur is an ALGORITHM: it runs in the REAL OpenCV through a callback (cv2.medianBlur), in place like the reference's call
typedef void (*shim_median_fn)(const uchar *src, uchar *dst, int rows, int cols, int ksize);
inline shim_median_fn &shim_median_callback() { static shim_median_fn f = nullptr; return f; }
inline void medianBlur(const Mat &src, Mat &dst, int ksize) {
    if (!shim_median_callback()) throw std::runtime_error("medianBlur: no callback installed");
    std::vector<uchar> in((size_t)src.rows * src.cols), out(in.size());
    for (int r = 0; r < src.rows; ++r) std::memcpy(&in[(size_t)r * src.cols], src.ptr<uchar>(r), (size_t)src.cols);
    shim_median_callback()(in.data(), out.data(), src.rows, src.cols, ksize);
    if (!(dst.data && dst.rows == src.rows && dst.cols == src.cols)) dst = Mat(src.rows, src.cols, CV_8U);
    for (int r = 0; r < src.rows; ++r) std::memcpy(dst.ptr<uchar>(r), &out[(size_t)r * src.cols], (size_t)src.cols);
}

// ---- additions for readFrame / correctImage (LocoMouse_class.cpp:1273-1406) --------------------------------------------
// normalize(F, F, 0, 255, NORM_MINMAX, CV_8UC1) on image data and flip are NOT implemented here: they run in the REAL
// OpenCV through a callback (op 0 = cv2.normalize(src, None, 0, 255, NORM_MINMAX, CV_8U), op 1 = cv2.flip(src, 1)).
typedef void (*shim_u8_op_fn)(int op, const uchar *src, uchar *dst, int rows, int cols);
inline shim_u8_op_fn &shim_u8_op_callback() { static shim_u8_op_fn f = nullptr; return f; }
inline void shim_u8_op(int op, const Mat &src, Mat &dst) {
    if (!shim_u8_op_callback()) throw std::runtime_error("no callback installed for normalize / flip");
    std::vector<uchar> in((size_t)src.rows * src.cols), out(in.size());
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) in[(size_t)r * src.cols + c] = src.ptr<uchar>(r)[c];
    shim_u8_op_callback()(op, in.data(), out.data(), src.rows, src.cols);
    Mat o(src.rows, src.cols, CV_8U);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) o.ptr<uchar>(r)[c] = out[(size_t)r * src.cols + c];
    dst = o;
}
inline void flip(const Mat &src, Mat &dst, int /*1: around the y axis*/) {
    Mat tmp;
    shim_u8_op(1, src, tmp);
    // cv::flip(I, I, 1) works in place: the caller's buffer (a view of I_PAD in the reference) must receive the pixels
    if (dst.data && dst.rows == src.rows && dst.cols == src.cols) {
        for (int r = 0; r < tmp.rows; ++r)
            for (int c = 0; c < tmp.cols; ++c) dst.ptr<uchar>(r)[c] = tmp.ptr<uchar>(r)[c];
    } else {
        dst = tmp;
    }
}
inline void extractChannel(const Mat &src, Mat &dst, int /*0*/) { dst = src; }  // the injected frames are single-channel
inline void subtract(const Mat &a, const Mat &b, Mat &dst) {  // saturating 8-bit a - b (3-argument form)
    if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("cv::subtract: size mismatch");
    Mat out(a.rows, a.cols, CV_8U);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            const int d = (int)a.ptr<uchar>(r)[c] - (int)b.ptr<uchar>(r)[c];
            out.ptr<uchar>(r)[c] = (uchar)(d < 0 ? 0 : d);
        }
    dst = out;
}
// 8-bit -> 8-bit Mat::convertTo(alpha, beta) (what `Iout = (Iout - r0) / (r1 - r0)` evaluates to): the real OpenCV's
// cv2.convertScaleAbs through a callback (same float multiply-add + rounding kernels; every value here is >= -0.5, so the
// absolute value changes nothing), except OpenCV's own shortcut: alpha == 1 and beta == 0 copies.
typedef void (*shim_scale_fn)(const uchar *src, uchar *dst, int rows, int cols, double alpha, double beta);
inline shim_scale_fn &shim_scale_callback() { static shim_scale_fn f = nullptr; return f; }
inline Mat &Mat::assign_scaled_u8(const MatExpr &e) {
    const Mat &a = *e.a;
    std::vector<uchar> in((size_t)a.rows * a.cols), out(in.size());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) in[(size_t)r * a.cols + c] = a.ptr<uchar>(r)[c];
    if (std::fabs(e.alpha - 1) < 2.220446049250313e-16 && std::fabs(e.beta) < 2.220446049250313e-16) {
        out = in;
    } else {
        if (!shim_scale_callback()) throw std::runtime_error("no callback installed for the scaled 8-bit conversion");
        shim_scale_callback()(in.data(), out.data(), a.rows, a.cols, e.alpha, e.beta);
    }
    // assignment to an existing Mat of the same size and type reuses its buffer (Mat::create is a no-op): ROI views stay views
    const bool inplace = data && rows == a.rows && cols == a.cols && type_ == CV_8U;
    if (!inplace) {
        Mat o(a.rows, a.cols, CV_8U);
        *this = o;
    }
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) ptr<uchar>(r)[c] = out[(size_t)r * cols + c];
    return *this;
}
// cv::calcHist for one 8-bit image, 256 unit bins on [0, 256): float counts in a 256 x 1 matrix
inline void calcHist(const Mat *images, int /*nimages*/, const int * /*channels*/, const Mat & /*mask*/, Mat &hist, int /*dims*/,
                     const int *histSize, const float ** /*ranges*/) {
    Mat h(*histSize, 1, CV_32F);
    for (int r = 0; r < images->rows; ++r)
        for (int c = 0; c < images->cols; ++c) h.ptr<float>(images->ptr<uchar>(r)[c])[0] += 1.f;
    hist = h;
}

class VideoCapture {  // hands out the injected frame
public:
    Mat next;
    VideoCapture &operator>>(Mat &f) {
        f = next;
        return *this;
    }
};

inline std::ostream &operator<<(std::ostream &o, const Mat &m) { return o << "Mat(" << m.rows << " x " << m.cols << ")"; }  // debug prints only
struct Scalar {
    double v[4];
    double operator()(int i) const { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
inline Scalar sum(const Mat &m) {  // double accumulation, as cv::sum
    Scalar s = {{0, 0, 0, 0}};
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c) s.v[0] += m.type() == CV_32F ? (double)m.ptr<float>(r)[c] : (double)m.ptr<uchar>(r)[c];
    return s;
}
// cv::LUT for single-channel 8-bit data: dst(i) = lut(src(i))
inline void LUT(const Mat &src, const Mat &lut, Mat &dst) {
    Mat out(src.rows, src.cols, CV_8U);
    const uchar *t = lut.ptr<uchar>(0);
    for (int r = 0; r < src.rows; ++r) {
        const uchar *s = src.ptr<uchar>(r);
        uchar *d = out.ptr<uchar>(r);
        for (int c = 0; c < src.cols; ++c) d[c] = t[s[c]];
    }
    dst = out;
}
template <typename T>
inline Rect_<T> operator+(const Rect_<T> &r, const Point_<T> &p) { return Rect_<T>(r.x + p.x, r.y + p.y, r.width, r.height); }

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
