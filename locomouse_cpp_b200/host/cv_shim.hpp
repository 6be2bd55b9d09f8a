// cv_shim.hpp — the handful of OpenCV core value types that appear in the reference's class and
// Candidates interfaces (cv::Point_, cv::Size_, cv::Rect_; LocoMouse_class.hpp:188-236,
// Candidates/Candidates.hpp:16-105).  The OpenCV C++ SDK is not part of this image, and pixel data
// never crosses this layer as cv::Mat any more (frames live in HBM), so these PODs are all the host
// mirror needs.  When the real <opencv2/core.hpp> is available, define LM_USE_OPENCV and the real
// types are used instead (layout compatible: public x/y/width/height members).
#pragma once
#ifdef LM_USE_OPENCV
#include <opencv2/core.hpp>
#else
#include <ostream>
#include <vector>
#ifndef CV_32SC1
#define CV_32SC1 4
#endif
namespace cv {
template <typename T>
struct Point_ {
    T x{}, y{};
    Point_() = default;
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_ operator+(const Point_ &o) const { return Point_(x + o.x, y + o.y); }
    bool operator==(const Point_ &o) const { return x == o.x && y == o.y; }
};
template <typename T>
struct Size_ {
    T width{}, height{};
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
template <typename T>
struct Rect_ {
    T x{}, y{}, width{}, height{};
    Rect_() = default;
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_ operator+(const Point_<T> &p) const { return Rect_(x + p.x, y + p.y, width, height); }
    T area() const { return width * height; }
};
// The only cv::Mat the host stages that follow detection exchange: 32-bit signed label / track matrices (match2nd's
// result, TRACK_INDEX_*, the exported N_FRAMES x 3 tracks).  Owning, continuous, row-major.
class Mat {
    std::vector<int> v_;

public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int /*type: CV_32SC1*/, int fill = 0) : v_((size_t)(r > 0 ? r : 0) * (size_t)(c > 0 ? c : 0), fill), rows(r), cols(c) {}
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, 0); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, 1); }
    bool empty() const { return v_.empty(); }
    bool isContinuous() const { return true; }
    int type() const { return CV_32SC1; }
    template <typename T>
    T *ptr(int r = 0) { return reinterpret_cast<T *>(v_.data() + (size_t)r * cols); }
    template <typename T>
    const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(v_.data() + (size_t)r * cols); }
    template <typename T>
    T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T>
    const T &at(int r, int c) const { return ptr<T>(r)[c]; }
    void copyTo(Mat &dst) const { dst = *this; }
};
using Point = Point_<int>;
using Size = Size_<int>;
using Rect = Rect_<int>;
template <typename T>
std::ostream &operator<<(std::ostream &o, const Point_<T> &p) { return o << "[" << p.x << ", " << p.y << "]"; }
template <typename T>
std::ostream &operator<<(std::ostream &o, const Rect_<T> &r) {
    return o << "[" << r.width << " x " << r.height << " from (" << r.x << ", " << r.y << ")]";
}
}  // namespace cv
#endif
