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

#include <fstream>

// Stand-ins for the declarations the compiled member functions need (the real ones, LocoMouse_class.hpp:110-350, drag in
// VideoCapture / FileStorage / MyMat members that these methods never touch).  Only data the methods read is kept.
struct LocoMouse_Parameters_Stub {
    bool LM_DEBUG = false;
    unsigned int N_tail_points = 15;
    int conn_comp_connectivity = 8;
    int median_filter_size = 11;   // LocoMouse_class.hpp:54-55
    int min_pixel_visible = 1;
};
typedef LocoMouse_Parameters_Stub LocoMouse_Parameters;  // computeMouseBox takes `const LocoMouse_Parameters &`
double medianvec(std::vector<double> &v, const int N);    // LocoMouse_class.hpp:356-358
double stdvec(std::vector<double> &v, int N);
class LocoMouse_Feature {  // sizes + velocity-matching boxes; boxes by the formula of LocoMouse_class.cpp:2954-2969
    cv::Size size_b, size_s;
    cv::Rect match_rect_b, match_rect_s;
    int tag_;  // >= 0: detector_*() return 1 x 1 kernels holding tag (bottom) / tag + 2 (side), read by the injected filter2D

public:
    LocoMouse_Feature(int tw_b, int th_b, int tw_s, int th_s, int tag = -1) : size_b(tw_b, th_b), size_s(tw_s, th_s), tag_(tag) {
        int new_b_w = round(((double)tw_b) / 2), new_b_h = round(((double)th_b) / 2);
        int new_t_w = round(((double)tw_s) / 2), new_t_h = round(((double)th_s) / 2);
        match_rect_b = cv::Rect(-(new_b_w / 2), -(new_b_h / 2), new_b_w, new_b_h);
        match_rect_s = cv::Rect(-(new_t_w / 2), -(new_t_h / 2), new_t_w, new_t_h);
    }
    cv::Size size_bottom() const { return size_b; }
    cv::Size size_side() const { return size_s; }
    cv::Rect match_box_bottom() const { return match_rect_b; }
    cv::Rect match_box_side() const { return match_rect_s; }
    // the detection code passes these to filter2D, whose outputs the harness injects
    cv::Mat tagged(int id) const {
        if (tag_ < 0) return cv::Mat();
        cv::Mat k(1, 1, CV_32F);
        k.ptr<float>(0)[0] = (float)id;
        return k;
    }
    cv::Mat detector_bottom() const { return tagged(tag_); }
    cv::Mat detector_side() const { return tagged(tag_ + 2); }
    double rho_bottom() const { return 0.0; }
    double rho_side() const { return 0.0; }
};
struct LocoMouse_Model_Stub {
    LocoMouse_Feature paw, snout, tail;
    LocoMouse_Model_Stub() : paw(30, 30, 30, 30, 0), snout(30, 30, 30, 30, 1), tail(30, 30, 30, 30) {}
};
// value type with the reference's interface (LocoMouse_class.hpp:33-45; ctor LocoMouse_class.cpp:3196-3202)
class LocoMouse_LocationPrior {
    cv::Point_<double> pos_;
    double md_;
    cv::Rect_<double> area_;

public:
    LocoMouse_LocationPrior(double x, double y, double md, double minx, double maxx, double miny, double maxy)
        : pos_(x, y), md_(md), area_(minx, miny, maxx - minx, maxy - miny) {}
    cv::Point_<double> position() const { return pos_; }
    cv::Rect_<double> area() const { return area_; }
    double max_distance() const { return md_; }
};
#include <sstream>
#include "_ref/ref_stream_formatter.inc"   // class Stream_Formatter (LocoMouse_class.hpp:379-406)
#include "MyMat.hpp"                      // the reference's own MyMat / MATSPARSE (MyMat/MyMat.hpp, compiled from MyMat.cpp)
#include "_ref/ref_match_to_range.inc"    // template matchToRange (LocoMouse_class.hpp:364-374)
class LocoMouse {
public:
    LocoMouse_Parameters_Stub LM_PARAMS;
    // tail stage (LocoMouse_class.hpp:188-236: same member names)
    LocoMouse_Model_Stub M;
    cv::Mat I_BOTTOM_MOUSE_PAD, I_SIDE_MOUSE_PAD, TAIL_MASK;
    // pass 1 (LocoMouse_TM_DE): imadjust_default + the members computeMouseBox_DE reads
    void imadjust_default(const cv::Mat &Iin, cv::Mat &Iout);
    unsigned int N_COLS = 0;
    cv::Rect BB_SIDE_VIEW;
    // readFrame / correctImage (same member names)
    cv::VideoCapture V;
    cv::Mat BKG, CALIBRATION;
    bool IMAGE_FLIP = false;
    int CURRENT_FRAME = -1;
    void readFrame(cv::Mat &I);
    void correctImage(cv::Mat &Iin, cv::Mat &Iout);
    // bottom / side candidate detection (same member names)
    cv::Mat I_BOTTOM_MOUSE, I_SIDE_MOUSE;
    cv::Rect BB_BOTTOM_TAIL, BB_UNPAD_MOUSE_BOTTOM, BB_UNPAD_MOUSE_SIDE;
    std::vector<std::vector<Candidate> > CANDIDATES_BOTTOM_PAW, CANDIDATES_BOTTOM_SNOUT, CANDIDATES_SIDE_PAW, CANDIDATES_SIDE_SNOUT;
    void detectBottomCandidates();
    void detectSideCandidates();
    std::vector<Candidate> detectPointCandidatesBottom(cv::Mat &I_VIEW_PAD, cv::Rect &BB_UNPAD, LocoMouse_Feature &M_FEAT, cv::Mat &I_bb_bottom_mask);
    std::vector<Candidate> detectPointCandidatesSide(cv::Mat &I_VIEW_PAD, cv::Rect &BB_UNPAD, LocoMouse_Feature &M_FEAT, cv::Mat &I_bb_bottom_mask);
    cv::Rect BB_BOTTOM_TAIL_PAD, BB_SIDE_TAIL_PAD, BB_UNPAD_TAIL_BOTTOM, BB_UNPAD_TAIL_SIDE;
    std::vector<cv::Mat> TRACKS_TAIL;
    void detectTail();
    cv::Mat detectLineCandidates(const LocoMouse_Feature &F, const unsigned int N_line_points, cv::Mat &TAIL_MASK);
    void selectLargestRegion(const cv::Mat &Iin, cv::Mat &Iout);
    void exportDebugVariablesUnused();
    MyMat unaryCostBox(std::vector<Candidate> &p_candidates, cv::Rect &BB, std::vector<LocoMouse_LocationPrior> location_prior);
    MATSPARSE pairwisePotential(std::vector<Candidate> &Ci, std::vector<Candidate> &Cip1, cv::Point_<double> &grid_mapping,
                                double grid_spacing, std::vector<cv::Point_<double> > &ONGi, cv::Size ONG_size,
                                double max_displacement_bottom, double alpha_vel_bottom, double pairwise_occluded_cost);
    // pass 1 of the base class (same member names)
    cv::Mat I_PAD, I_PREV_PAD;
    void largestBWAreaObject(cv::Mat &Iin, cv::Mat &Iout);
    void computeMouseBox(cv::Mat &I_median, cv::Mat &I, cv::Mat &I_side_view, cv::Mat &I_bottom_view, double &bb_x, double &bb_y_bottom,
                         double &bb_y_side, double &bb_width, double &bb_height_bottom, double &bb_height_side, const LocoMouse_Parameters &LM_PARAMS);
    void computeMouseBoxSize(std::vector<double> &bb_w, std::vector<double> &bb_hb, std::vector<double> &bb_ht, cv::Rect &BB_side, cv::Rect &BB_bottom);
    void storePreviousImage();
    MATSPARSE pairwisePotential_SideView(const std::vector<uint> &Zi, const std::vector<uint> &Zip1, double grid_mapping, double grid_spacing,
                                         const std::vector<uint> &ONGi, const unsigned int Nong, const double max_displacement_bottom,
                                         const double alpha_vel_bottom, const double pairwise_occluded_cost);
    std::ofstream DEBUG_TEXT;
    void imadjust(const cv::Mat &Iin, cv::Mat &Iout, double low_in, double high_in, double low_out, double high_out);
    std::vector<P22D> matchingWithVelocityConstraint(std::vector<Candidate> &Candidates_b, std::vector<Candidate> &Candidates_t,
                                                     const cv::Mat &Ibbb, const cv::Mat &Itbb, const cv::Mat &Ibbb_prev,
                                                     const cv::Mat &Itbbb_prev, const cv::Point_<int> padding_pre_bottom,
                                                     const cv::Point_<int> padding_pre_side, bool vel_check, LocoMouse_Feature &F, double T,
                                                     bool debug);
    cv::Mat xDist(const std::vector<Candidate> &P1, const std::vector<Candidate> &P2);
    std::vector<P22D> matchViews(const cv::Mat &boolD, const cv::Mat &D_side_weight, const std::vector<Candidate> &C_b,
                                 const std::vector<Candidate> &C_t, bool vel_check, LocoMouse_Feature &F, const cv::Mat &Ibbb,
                                 const cv::Mat &Itbb, const cv::Mat &Ibbb_prev, const cv::Mat &Itbb_prev,
                                 const cv::Point_<int> padding_pre_bottom, const cv::Point_<int> padding_pre_side, bool debug);
    bool checkVelCriterion(const cv::Mat &I, const cv::Mat &I_prev, const cv::Rect &im_box, int box_area, double alpha, double T);
};

#include "_ref/ref_nms_body.inc"      // vecmovingaverage, nmsMax, peakClustering   (LocoMouse_class.cpp:1559-1905)
#include "_ref/ref_hpp_body.inc"      // template firstLastOverT                     (LocoMouse_class.hpp:411-442)
#include "_ref/ref_imadjust_body.inc" // LocoMouse::imadjust                         (LocoMouse_class.cpp:3204-3242)
class LocoMouse_TM_DE : public LocoMouse {  // LocoMouse_TM_DE.hpp:27-29 defaults
public:
    int MIN_PIXEL_COUNT = 10;
    double WIDTH_MARGIN = 1.1;
    double SIDE_THRESHOLD = 255 * 0.05;
    void computeMouseBox_DE(cv::Mat &I_SIDE, double &bb_x);
};
class LocoMouse_TM : public LocoMouse {  // members computeMouseBox_DD / bwAreaOpen read (LocoMouse_TM.hpp:22-40)
public:
    unsigned int BOTTOM_THRESHOLD = 0, SIDE_THRESHOLD = 0, MIN_PIXEL_COUNT = 1;
    unsigned int ZERO_COL_PRE = 0, ZERO_COL_POST = 0, ZERO_ROW_PRE = 0, ZERO_ROW_POST = 0;
    cv::Mat DISK_FILTER;
    cv::Mat I;   // LocoMouse::I, the calibrated frame LocoMouse_TM::readFrame() fills (LocoMouse_class.hpp:170)
    cv::Mat bwAreaOpen(cv::Mat &Iin);
    void computeMouseBox_DD(cv::Mat &I_SIDE, double &bb_x);
    void imfill(const cv::Mat &Iin, cv::Mat &Iout);
    void readFrame();
};
#include "_ref/ref_imadjust_default_body.inc"  // LocoMouse::imadjust_default (LocoMouse_class.cpp:3244-3311)
#include "_ref/ref_box_de_body.inc"            // LocoMouse_TM_DE::computeMouseBox_DE (LocoMouse_TM_DE.cpp:56-113)
#include "_ref/ref_read_body.inc"     // readFrame(cv::Mat&), correctImage (LocoMouse_class.cpp:1273-1406)
#include "_ref/ref_detect_body.inc"   // detectBottomCandidates, detectSideCandidates, detectPointCandidates* (LocoMouse_class.cpp:771-870)
#include "_ref/ref_tail_body.inc"     // detectTail, detectLineCandidates, selectLargestRegion (LocoMouse_class.cpp:2541-2767)
#include "_ref/ref_cost_body.inc"     // unaryCostBox, pairwisePotential (LocoMouse_class.cpp:1909-2070)
#include "_ref/ref_pair_body.inc"     // matchingWithVelocityConstraint, xDist, matchViews, checkVelCriterion (1023-1267)
#include "_ref/ref_side_cost_body.inc" // pairwisePotential_SideView (LocoMouse_class.cpp:2073-2150)
#include "_ref/ref_box_base_body.inc"  // largestBWAreaObject, computeMouseBox (LocoMouse_class.cpp:921-997)
#include "_ref/ref_box_size_body.inc"  // computeMouseBoxSize, storePreviousImage, medianvec, stdvec (LocoMouse_class.cpp:1481-1556)
using namespace cv;   // LocoMouse_TM.cpp is written against `using namespace cv` (Locomouse.hpp:12)
using namespace std;
#include "_ref/ref_box_tm_body.inc"    // LocoMouse_TM::bwAreaOpen, computeMouseBox_DD, readFrame, imfill (LocoMouse_TM.cpp:158-269)

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

// matchingWithVelocityConstraint on padded crops (I_*_MOUSE_PAD of the current and the previous frame, row-major u8) with the
// candidates in unpadded-crop coordinates, exactly the arguments matchBottomSideCandidates passes (class.cpp:999-1021).
// Output: per bottom candidate number_of_candidates() and the stored (y, score) pairs, concatenated.
int ref_match_views(const ref_cand *cb, int nb, const ref_cand *cs, int ns, int vel_check, int tw_b, int th_b, int tw_s, int th_s,
                    double T, const unsigned char *Ibbb, const unsigned char *Ibbb_prev, int wb, int hb, const unsigned char *Itbb,
                    const unsigned char *Itbb_prev, int ws, int hs, int pad_bx, int pad_by, int pad_sx, int pad_sy, int *match_n,
                    int *match_y, double *match_s, int cap) {
    std::vector<Candidate> B, S;
    for (int i = 0; i < nb; ++i) B.push_back(Candidate(cb[i].x, cb[i].y, cb[i].s));
    for (int i = 0; i < ns; ++i) S.push_back(Candidate(cs[i].x, cs[i].y, cs[i].s));
    cv::Mat Ib(hb, wb, CV_8U, (void *)Ibbb, (size_t)wb), Ibp(hb, wb, CV_8U, (void *)Ibbb_prev, (size_t)wb);
    cv::Mat It(hs, ws, CV_8U, (void *)Itbb, (size_t)ws), Itp(hs, ws, CV_8U, (void *)Itbb_prev, (size_t)ws);
    LocoMouse_Feature F(tw_b, th_b, tw_s, th_s);
    LocoMouse L;
    std::vector<P22D> P;
    try {
        P = L.matchingWithVelocityConstraint(B, S, Ib, It, Ibp, Itp, cv::Point_<int>(pad_bx, pad_by), cv::Point_<int>(pad_sx, pad_sy),
                                             vel_check != 0, F, T, false);
    } catch (const std::exception &) {
        return -1;  // the reference throws (CV_Assert / ROI): e.g. ovlp == 0 makes the weights NaN
    }
    int total = 0;
    for (size_t i = 0; i < P.size(); ++i) {
        const int n = P[i].number_of_candidates();
        match_n[i] = n;
        for (int k = 0; k < n; ++k, ++total)
            if (total < cap) {
                match_y[total] = P[i].y_side_coord((uint)k);
                match_s[total] = P[i].score_side((uint)k);
            }
    }
    return (int)P.size() * 100000 + total;
}

// detectTail on injected tail score maps (row-major f32, hb x tw and hs x tw: what the two filter2D calls of
// detectLineCandidates return on the unpadded tail boxes).  cc: connectedComponentsWithStats of the REAL OpenCV (a ctypes
// callback into cv2).  Outputs: TRACKS_TAIL.back() (3 x n_points int32) and TAIL_MASK (hb x tw, 0 / 255).
static const float *g_maps[2];
static int g_map_i = 0;
static void next_map(float *dst, int rows, int cols, int /*id*/) {
    const float *src = g_maps[g_map_i++ & 1];
    for (int i = 0; i < rows * cols; ++i) dst[i] = src[i];
}
int ref_detect_tail(const float *score_b, const float *score_s, int hb, int hs, int tw, int conn, int n_points, cv::shim_cc_fn cc,
                    int *tracks, unsigned char *tail_mask) {
    try {
        LocoMouse L;
        L.LM_PARAMS.N_tail_points = (unsigned int)n_points;
        L.LM_PARAMS.conn_comp_connectivity = conn;
        L.I_BOTTOM_MOUSE_PAD = cv::Mat(hb, tw, CV_8U);
        L.I_SIDE_MOUSE_PAD = cv::Mat(hs, tw, CV_8U);
        L.BB_BOTTOM_TAIL_PAD = L.BB_UNPAD_TAIL_BOTTOM = cv::Rect(0, 0, tw, hb);
        L.BB_SIDE_TAIL_PAD = L.BB_UNPAD_TAIL_SIDE = cv::Rect(0, 0, tw, hs);
        g_maps[0] = score_b;
        g_maps[1] = score_s;
        g_map_i = 0;
        cv::shim_cc_callback() = cc;
        cv::shim_filter_callback() = next_map;
        L.detectTail();
        const cv::Mat &T = L.TRACKS_TAIL.back();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < n_points; ++c) tracks[r * n_points + c] = T.ptr<int>(r)[c];
        for (int r = 0; r < hb; ++r)
            for (int c = 0; c < tw; ++c) tail_mask[r * tw + c] = L.TAIL_MASK.ptr<unsigned char>(r)[c];
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// readFrame(I) on an injected raw frame: background subtraction, normalisation and mirror run in the real OpenCV (callback),
// correctImage (the calibration gather) is the reference's own loop.  out: n_rows x n_cols.
int ref_read_frame(const unsigned char *frame, const unsigned char *bkg, int vr, int vc, const int *calib, int n_rows, int n_cols, int flip,
                   cv::shim_u8_op_fn op, unsigned char *out) {
    try {
        LocoMouse L;
        L.V.next = cv::Mat(vr, vc, CV_8U, (void *)frame, (size_t)vc);
        L.BKG = cv::Mat(vr, vc, CV_8U, (void *)bkg, (size_t)vc);
        L.CALIBRATION = cv::Mat(n_rows, n_cols, CV_32S, (void *)calib, (size_t)n_cols * 4);
        L.IMAGE_FLIP = flip != 0;
        cv::shim_u8_op_callback() = op;
        cv::Mat I(n_rows, n_cols, CV_8U, (void *)out, (size_t)n_cols);
        L.readFrame(I);
        return L.CURRENT_FRAME;  // 0 after the first frame
    } catch (const std::exception &) {
        return -1;
    }
}

// computeMouseBox_DE on a calibrated side view (u8, side_h x n_cols, modified in place as the reference does): returns bb_x.
// The scaled 8-bit conversion inside imadjust_default runs in the real OpenCV (callback).  adjusted (optional): the side view
// after imadjust_default and the zeroed bands.
int ref_mouse_box_de(unsigned char *side, int side_h, int n_cols, double threshold, int min_count, double margin, cv::shim_scale_fn scale,
                     double *bb_x) {
    try {
        LocoMouse_TM_DE L;
        L.N_COLS = (unsigned int)n_cols;
        L.BB_SIDE_VIEW = cv::Rect(0, 0, n_cols, side_h);
        L.SIDE_THRESHOLD = threshold;
        L.MIN_PIXEL_COUNT = min_count;
        L.WIDTH_MARGIN = margin;
        cv::shim_scale_callback() = scale;
        cv::Mat I(side_h, n_cols, CV_8U, (void *)side, (size_t)n_cols);
        L.computeMouseBox_DE(I, *bb_x);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// LocoMouse_TM::computeMouseBox_DD on a calibrated side view (u8, side_h x n_cols, modified in place as the reference does):
// returns bb_x and (optional) row_sums[n_cols].  zero[4] = ZERO_COL_PRE, ZERO_COL_POST, ZERO_ROW_PRE, ZERO_ROW_POST; disk: the DISK_FILTER matrix (float,
// dk x dk).  connectedComponentsWithStats, filter2D, floodFill and the scaled 8-bit conversions run in the real OpenCV.
int ref_mouse_box_dd(unsigned char *side, int side_h, int n_cols, int threshold, int min_pixel_count, int min_pixel_visible, int conn,
                     const int *zero, const float *disk, int dk, cv::shim_scale_fn scale, cv::shim_cc_fn cc, cv::shim_filter_u8_fn filt,
                     cv::shim_flood_fn flood, double *bb_x, int *row_sums) {
    try {
        LocoMouse_TM L;
        L.N_COLS = (unsigned int)n_cols;
        L.BB_SIDE_VIEW = cv::Rect(0, 0, n_cols, side_h);
        L.SIDE_THRESHOLD = (unsigned int)threshold;
        L.MIN_PIXEL_COUNT = (unsigned int)min_pixel_count;
        L.LM_PARAMS.min_pixel_visible = min_pixel_visible;
        L.LM_PARAMS.conn_comp_connectivity = conn;
        L.ZERO_COL_PRE = zero[0];
        L.ZERO_COL_POST = zero[1];
        L.ZERO_ROW_PRE = zero[2];
        L.ZERO_ROW_POST = zero[3];
        L.DISK_FILTER = cv::Mat(dk, dk, CV_32F, (void *)disk, (size_t)dk * sizeof(float));
        cv::shim_scale_callback() = scale;
        cv::shim_cc_callback() = cc;
        cv::shim_filter_u8_callback() = filt;
        cv::shim_flood_callback() = flood;
        cv::Mat I(side_h, n_cols, CV_8U, (void *)side, (size_t)n_cols);
        L.computeMouseBox_DD(I, *bb_x);
        if (row_sums) {  // Row_side, the CV_32S column sums firstLastOverT was handed (LocoMouse_TM.cpp:223)
            const cv::Mat &R = cv::shim_last_reduce_i32();
            for (int c = 0; c < n_cols && c < R.cols; ++c) row_sums[c] = R.ptr<int>(0)[c];
        }
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_mouse_box_dd: %s\n", e.what());
        return -1;
    }
}

// detectBottomCandidates + detectSideCandidates on one frame.  crops: the unpadded bottom / side crops (u8, hb x w, hs x w);
// tail_mask: TAIL_MASK (hb x tail_w, 0 / 255); maps[4]: the filter2D outputs over the PADDED crops ((hb + pad_b) x (w + pad_b_x) ...)
// for bottom paw, bottom snout, side paw, side snout, with the unpadded windows at unpad_b / unpad_s = {x, y};
// tsz: template sizes {paw_b w, h, paw_s w, h, snout_b w, h, snout_s w, h}.  out: 4 lists of up to cap candidates, counts n[4].
static const float *g_det_maps[4];
static void det_map(float *dst, int rows, int cols, int id) {
    const float *src = g_det_maps[id & 3];
    for (int i = 0; i < rows * cols; ++i) dst[i] = src[i];
}
int ref_detect_candidates(const unsigned char *crop_b, const unsigned char *crop_s, int hb, int hs, int w, const unsigned char *tail_mask,
                          int tail_w, const float *const *maps, const int *pad_b, const int *pad_s, const int *unpad_b, const int *unpad_s,
                          const int *tsz, ref_cand *out, int cap, int *n) {
    try {
        LocoMouse L;
        L.M.paw = LocoMouse_Feature(tsz[0], tsz[1], tsz[2], tsz[3], 0);
        L.M.snout = LocoMouse_Feature(tsz[4], tsz[5], tsz[6], tsz[7], 1);
        L.I_BOTTOM_MOUSE = cv::Mat(hb, w, CV_8U, (void *)crop_b, (size_t)w);
        L.I_SIDE_MOUSE = cv::Mat(hs, w, CV_8U, (void *)crop_s, (size_t)w);
        L.TAIL_MASK = cv::Mat(hb, tail_w, CV_8U, (void *)tail_mask, (size_t)tail_w);
        L.BB_BOTTOM_TAIL = cv::Rect(0, 0, tail_w, hb);
        L.I_BOTTOM_MOUSE_PAD = cv::Mat(pad_b[1], pad_b[0], CV_8U);   // only its size is read (by the injected filter2D)
        L.I_SIDE_MOUSE_PAD = cv::Mat(pad_s[1], pad_s[0], CV_8U);
        L.BB_UNPAD_MOUSE_BOTTOM = cv::Rect(unpad_b[0], unpad_b[1], w, hb);
        L.BB_UNPAD_MOUSE_SIDE = cv::Rect(unpad_s[0], unpad_s[1], w, hs);
        for (int k = 0; k < 4; ++k) g_det_maps[k] = maps[k];
        cv::shim_filter_callback() = det_map;
        L.detectBottomCandidates();
        L.detectSideCandidates();
        const std::vector<Candidate> *lists[4] = {&L.CANDIDATES_BOTTOM_PAW.back(), &L.CANDIDATES_BOTTOM_SNOUT.back(), &L.CANDIDATES_SIDE_PAW.back(),
                                                  &L.CANDIDATES_SIDE_SNOUT.back()};
        for (int k = 0; k < 4; ++k) {
            n[k] = (int)lists[k]->size();
            for (int i = 0; i < n[k] && i < cap; ++i) {
                out[k * cap + i].x = (*lists[k])[i].point().x;
                out[k * cap + i].y = (*lists[k])[i].point().y;
                out[k * cap + i].s = (*lists[k])[i].score();
            }
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

// unaryCostBox on a candidate list: out = the returned MyMat's column-major values (n x n_priors).
// priors: n_priors x 7 doubles {x, y, max_distance, min_x, max_x, min_y, max_y} (the reference constructor's arguments).
void ref_unary_cost_box(const ref_cand *c, int n, int bb_w, int bb_h, const double *priors, int n_priors, double *out) {
    std::vector<Candidate> C;
    for (int i = 0; i < n; ++i) C.push_back(Candidate(c[i].x, c[i].y, c[i].s));
    std::vector<LocoMouse_LocationPrior> P;
    for (int j = 0; j < n_priors; ++j) {
        const double *q = priors + 7 * j;
        P.push_back(LocoMouse_LocationPrior(q[0], q[1], q[2], q[3], q[4], q[5], q[6]));
    }
    cv::Rect BB(0, 0, bb_w, bb_h);
    LocoMouse L;
    MyMat M = L.unaryCostBox(C, BB, P);
    for (int k = 0; k < M.Numel(); ++k) out[k] = M.getValues()[k];
}

// pairwisePotential -> the MATSPARSE it returns: jc[n_cols + 1], ir / pr [nnz]; dims = {n_rows, n_cols, nnz}.
int ref_pairwise_potential(const ref_cand *ci, int ni, const ref_cand *cip1, int nip1, double grid_x, double grid_y, double spacing,
                           int ong_w, int ong_h, double max_disp, double alpha_vel, double occluded_cost, int *jc, int *ir, double *pr,
                           int cap, int *dims) {
    std::vector<Candidate> A, B;
    for (int i = 0; i < ni; ++i) A.push_back(Candidate(ci[i].x, ci[i].y, ci[i].s));
    for (int i = 0; i < nip1; ++i) B.push_back(Candidate(cip1[i].x, cip1[i].y, cip1[i].s));
    cv::Point_<double> gm(grid_x, grid_y);
    std::vector<cv::Point_<double> > ONG((size_t)ong_w * ong_h);   // only its size is read (LocoMouse_class.cpp:1961)
    LocoMouse L;
    MATSPARSE S = L.pairwisePotential(A, B, gm, spacing, ONG, cv::Size(ong_w, ong_h), max_disp, alpha_vel, occluded_cost);
    dims[0] = S.Nrows();
    dims[1] = S.Ncols();
    dims[2] = S.nz();
    if (S.nz() > cap) return 1;
    for (int c = 0; c <= S.Ncols(); ++c) jc[c] = S.getJc()[c];
    for (int k = 0; k < S.nz(); ++k) {
        ir[k] = S.getIr()[k];
        pr[k] = S.getPr()[k];
    }
    return 0;
}

// pairwisePotential_SideView -> the MATSPARSE it returns (side-view tracker transitions of bestSideViewMatch)
int ref_pairwise_potential_side(const unsigned int *zi, int ni, const unsigned int *zip1, int nip1, double grid_mapping, double spacing, int nong,
                                double max_disp, double alpha_vel, double occluded_cost, int *jc, int *ir, double *pr, int cap, int *dims) {
    std::vector<uint> A(zi, zi + ni), B(zip1, zip1 + nip1), ONG((size_t)nong);
    LocoMouse L;
    MATSPARSE S = L.pairwisePotential_SideView(A, B, grid_mapping, spacing, ONG, (unsigned int)nong, max_disp, alpha_vel, occluded_cost);
    dims[0] = S.Nrows();
    dims[1] = S.Ncols();
    dims[2] = S.nz();
    if (S.nz() > cap) return 1;
    for (int c = 0; c <= S.Ncols(); ++c) jc[c] = S.getJc()[c];
    for (int k = 0; k < S.nz(); ++k) {
        ir[k] = S.getIr()[k];
        pr[k] = S.getPr()[k];
    }
    return 0;
}

// computeMouseBox for a SEQUENCE of frames, set up as computeBoundingBox does (LocoMouse_class.cpp:579-631): one padded
// I_median kept across the frames, readFrame's output (here: the injected calibrated images) written into its centre, the
// two views bound to it; bb_y_bottom made absolute.  medianBlur and connectedComponentsWithStats run in the real OpenCV.
int ref_compute_mouse_box(const unsigned char *images, int n, int n_rows, int n_cols, const int *side, const int *bottom, int median_size,
                          int min_pixel_visible, int conn, cv::shim_median_fn med, cv::shim_cc_fn cc, double *box) {
    try {
        LocoMouse L;
        L.LM_PARAMS.median_filter_size = median_size;
        L.LM_PARAMS.min_pixel_visible = min_pixel_visible;
        L.LM_PARAMS.conn_comp_connectivity = conn;
        cv::shim_median_callback() = med;
        cv::shim_cc_callback() = cc;
        const int off = median_size / 2, pad = off * 2;
        cv::Mat I_median = cv::Mat::zeros(n_rows + pad, n_cols + pad, CV_8UC1);
        cv::Mat I_center = I_median(cv::Rect(off, off, n_cols, n_rows));
        for (int f = 0; f < n; ++f) {
            cv::Mat I_bottom_view = I_center(cv::Rect(bottom[0], bottom[1], bottom[2], bottom[3]));
            cv::Mat I_side_view = I_center(cv::Rect(side[0], side[1], side[2], side[3]));
            for (int r = 0; r < n_rows; ++r) std::memcpy(I_center.ptr<unsigned char>(r), images + ((size_t)f * n_rows + r) * n_cols, (size_t)n_cols);
            double *o = box + (size_t)f * 6;
            L.computeMouseBox(I_median, I_center, I_side_view, I_bottom_view, o[0], o[1], o[2], o[3], o[4], o[5], L.LM_PARAMS);
            o[1] += bottom[1];
        }
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
// computeMouseBoxSize: size = {BB_bottom.width, BB_bottom.height, BB_side.height}; the three series are sorted in place
int ref_mouse_box_size(double *w, double *hb, double *hs, int n, int *size) {
    std::vector<double> a(w, w + n), b(hb, hb + n), c(hs, hs + n);
    cv::Rect side, bottom;
    LocoMouse L;
    L.computeMouseBoxSize(a, b, c, side, bottom);
    size[0] = bottom.width;
    size[1] = bottom.height;
    size[2] = side.height;
    std::copy(a.begin(), a.end(), w);
    std::copy(b.begin(), b.end(), hb);
    std::copy(c.begin(), c.end(), hs);
    return side.width == bottom.width ? 0 : 1;
}

// The elementwise primitives of the shim that the pairing code calls, exposed so that tests/test_oracle_vs_cv2.py can pin
// each of them against the real OpenCV (cv2): D <= ovlp, normalize(MINMAX), reduce(SUM) both ways, convertTo + the
// alpha * A + beta expression, and checkVelCriterion's subtract / threshold / sum chain.
void ref_shim_primitives(const int *D, int rows, int cols, int ovlp, unsigned char *bool_raw, unsigned char *bool_norm,
                         float *colsum, float *rowsum, double *weight, const unsigned char *a, const unsigned char *b, int r2,
                         int c2, double thr, double *count) {
    cv::Mat Dm(rows, cols, CV_32SC1, (void *)D, (size_t)cols * 4);
    cv::Mat boolD = Dm <= ovlp;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) bool_raw[r * cols + c] = boolD.ptr<uchar>(r)[c];
    cv::normalize(boolD, boolD, 0, 1, cv::NORM_MINMAX, -1);
    cv::Mat cs, rs, W;
    cv::reduce(boolD, cs, 0, cv::CV_REDUCE_SUM, CV_32FC1);
    cv::reduce(boolD, rs, 1, cv::CV_REDUCE_SUM, CV_32FC1);
    Dm.convertTo(W, CV_64FC1);
    W = 1 - (W / (double)ovlp);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            bool_norm[r * cols + c] = boolD.ptr<uchar>(r)[c];
            weight[r * cols + c] = W.ptr<double>(r)[c];
        }
    for (int c = 0; c < cols; ++c) colsum[c] = cs.ptr<float>(0)[c];
    for (int r = 0; r < rows; ++r) rowsum[r] = rs.ptr<float>(0)[r];
    cv::Mat A(r2, c2, CV_8U, (void *)a, (size_t)c2), B(r2, c2, CV_8U, (void *)b, (size_t)c2), S;
    cv::subtract(A, B, S, cv::noArray(), CV_8UC1);
    cv::threshold(S, S, thr, 1, cv::THRESH_BINARY);
    *count = cv::sum(S)(0);
}

// Second batch of shim primitives (used by the tail / candidate / readFrame code), exposed for tests/test_oracle_vs_cv2.py:
// out8[0] = convertTo_u8(threshold(F, 0, 1, BINARY)), out8[1] = threshold(A, 25.5, 255, BINARY_INV), out8[2] = repeat(reduce(A, 0, MAX), rows, 1),
// out8[3] = (F > 0) & A, out8[4] = compare(labels == value), out8[5] = subtract(A, B), out8[6] = A.setTo(7, B) (in place on a copy);
// outf = F.setTo(0, B); mom = moments(A, binary = true) {m00, m10, m01}.
void ref_shim_primitives2(const unsigned char *a, const unsigned char *b, const float *f, const unsigned short *labels, int value, int rows,
                          int cols, unsigned char *out8, float *outf, double *mom) {
    cv::Mat A(rows, cols, CV_8U, (void *)a, (size_t)cols), B(rows, cols, CV_8U, (void *)b, (size_t)cols);
    cv::Mat F(rows, cols, CV_32F, (void *)f, (size_t)cols * 4), Lb(rows, cols, CV_16U, (void *)labels, (size_t)cols * 2);
    cv::Mat t, r[7];
    cv::threshold(F, t, 0, 1, cv::THRESH_BINARY);
    t.convertTo(r[0], CV_8UC1);
    cv::threshold(A, r[1], 25.5, 255, cv::THRESH_BINARY_INV);
    cv::Mat mx;
    cv::reduce(A, mx, 0, cv::CV_REDUCE_MAX);
    r[2] = cv::repeat(mx, rows, 1);
    r[3] = (F > 0) & A;
    cv::compare(Lb, value, r[4], cv::CMP_EQ);
    cv::subtract(A, B, r[5]);
    A.convertTo(r[6], CV_8U);   // a copy
    r[6].setTo(7, B);
    cv::Mat Fc;
    F.convertTo(Fc, CV_32F);
    Fc.setTo(0, B);
    for (int k = 0; k < 7; ++k)
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) out8[((size_t)k * rows + y) * cols + x] = r[k].ptr<unsigned char>(y)[x];
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) outf[(size_t)y * cols + x] = Fc.ptr<float>(y)[x];
    cv::Moments M = cv::moments(A, true);
    mom[0] = M.m00;
    mom[1] = M.m10;
    mom[2] = M.m01;
}

// calcHist (256 unit bins) + cv::sum of the float histogram, and colRange / rowRange / setTo views (pass-1 code), for the cv2 check
void ref_shim_primitives3(const unsigned char *a, int rows, int cols, float *hist, double *total, unsigned char *banded) {
    cv::Mat A(rows, cols, CV_8U, (void *)a, (size_t)cols), H;
    const int nb = 256;
    float range[] = {0, 256};
    const float *hr = {range};
    cv::calcHist(&A, 1, 0, cv::Mat(), H, 1, &nb, &hr);
    for (int i = 0; i < 256; ++i) hist[i] = H.ptr<float>(0)[i];
    *total = cv::sum(H)(0);
    cv::Mat B;
    A.copyTo(B);
    B.colRange(0, cols / 4).setTo(0);
    B.rowRange(rows / 2, rows).setTo(0);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) banded[r * cols + c] = B.ptr<unsigned char>(r)[c];
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
