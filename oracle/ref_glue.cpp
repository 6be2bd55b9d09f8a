// oracle/ref_glue.cpp — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own nmsMax /
// peakClustering (LocoMouse_Core/LocoMouse_class.cpp:1610-1905, extracted verbatim at build time into
// oracle/_ref/ref_nms_body.inc by oracle/Makefile — never committed) and its Candidate / P22D classes
// (Candidates/Candidates.cpp, compiled from /root/reference directly).  See oracle/ref_shim/opencv2/core.hpp.
#include <map>
#include <algorithm>
#include "Candidates.hpp"  // the reference's header, found through -I/root/reference/Candidates

#include <cmath>
#include <cstdint>
#include <functional>
#include <numeric>

// stand-in for the class whose member LocoMouse::imadjust is compiled below (the real declaration,
// LocoMouse_class.hpp:170-350, drags in VideoCapture / FileStorage members that the method never touches)
class LocoMouse {
public:
    void imadjust(const cv::Mat &Iin, cv::Mat &Iout, double low_in, double high_in, double low_out, double high_out);
};

#include "_ref/ref_nms_body.inc"      // vecmovingaverage, nmsMax, peakClustering   (LocoMouse_class.cpp:1559-1905)
#include "_ref/ref_hpp_body.inc"      // template firstLastOverT                     (LocoMouse_class.hpp:411-442)
#include "_ref/ref_imadjust_body.inc" // LocoMouse::imadjust                         (LocoMouse_class.cpp:3204-3242)

extern "C" {

struct ref_cand {
    int x, y;
    double s;
};

static int emit(const std::vector<Candidate> &c, ref_cand *out, int cap) {
    for (size_t i = 0; i < c.size() && (int)i < cap; ++i) {
        out[i].x = c[i].point().x;
        out[i].y = c[i].point().y;
        out[i].s = c[i].score();
    }
    return (int)c.size();
}

int ref_nms_max(const float *scores, int rows, int cols, int box_w, int box_h, double overlap, ref_cand *out, int cap) {
    cv::Mat m(rows, cols, CV_32F, (void *)scores, (size_t)cols * sizeof(float));
    return emit(nmsMax(m, cv::Size(box_w, box_h), overlap), out, cap);
}

int ref_peak_clustering(const float *scores, int rows, int cols, int box_w, int box_h, int method, double overlap, ref_cand *out,
                        int cap) {
    cv::Mat m(rows, cols, CV_32F, (void *)scores, (size_t)cols * sizeof(float));
    return emit(peakClustering(m, cv::Size(box_w, box_h), method, overlap, false), out, cap);
}

// P22D as matchViews builds it (class.cpp:1160-1251): constructed from (bottom, first side or the (-1,-1,-1) sentinel), further
// side candidates added; returns number_of_candidates() and the stored side entries.
int ref_p22d(int xb, int yb, double sb, int n_side, const int *ys, const double *ss, int *out_y, double *out_s, int cap) {
    P22D p = n_side == 0 ? P22D(Candidate(xb, yb, sb), Candidate(-1, -1, -1)) : P22D(Candidate(xb, yb, sb), Candidate(xb, ys[0], ss[0]));
    for (int i = 1; i < n_side; ++i) p.add_side_candidate(Candidate(xb, ys[i], ss[i]));
    const int n = p.number_of_candidates();
    for (int i = 0; i < n && i < cap; ++i) {
        out_y[i] = p.y_side_coord((uint)i);
        out_s[i] = p.score_side((uint)i);
    }
    return n;
}

void ref_vecmovingaverage(const double *v, int n, int window, unsigned int *out) {
    std::vector<double> a(v, v + n);
    std::vector<unsigned int> o(n, 0u);
    vecmovingaverage(a, o, window);
    for (int i = 0; i < n; ++i) out[i] = o[i];
}

void ref_first_last_over_t(const float *values, unsigned int L, int th, int *first_last) {
    cv::Mat m(1, (int)L, CV_32F, (void *)values, (size_t)L * sizeof(float));
    firstLastOverT<int>(m, L, first_last, th);
}

void ref_imadjust_lut(double low_in, double high_in, double low_out, double high_out, unsigned char *lut) {
    cv::Mat ramp(1, 256, CV_8U), out(1, 256, CV_8U);
    for (int i = 0; i < 256; ++i) ramp.ptr<uchar>(0)[i] = (uchar)i;
    LocoMouse L;
    L.imadjust(ramp, out, low_in, high_in, low_out, high_out);
    for (int i = 0; i < 256; ++i) lut[i] = out.ptr<uchar>(0)[i];
}

int ref_default_candidate(int *x, int *y, double *s) {
    Candidate c;
    *x = c.point().x;
    *y = c.point().y;
    *s = c.score();
    P22D p;
    return p.number_of_candidates() * 1000 + p.y_side_coord(0);  // 0 candidates, sentinel -1  ->  -1
}
}
