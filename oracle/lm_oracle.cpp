/*
 * lm_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY, see lm_oracle.h).
 *
 * Literal CPU restatement of the per-frame detection path of careylab/LocoMouse_cpp.  No code is
 * copied from the reference: each function re-derives the arithmetic of the cited lines
 * (paths relative to the reference root; "class.cpp" = LocoMouse_Core/LocoMouse_class.cpp) and of the
 * OpenCV primitives those lines call (OpenCV is an un-vendored, unpinned dependency of the
 * reference — CMakeLists.txt:3 — whose documented semantics are restated here and pinned against
 * cv2 4.13 in tests/test_oracle_vs_cv2.py).
 *
 * Pin status (the reference ships no tests or golden vectors): readFrame / correctImage, imadjust, the tail
 * stage, the per-frame candidate glue, nmsMax, peakClustering, the pairing stage, the cost builders, vecmovingaverage,
 * firstLastOverT and the Candidate / P22D / MyMat / MATSPARSE value types are pinned against the REFERENCE'S OWN SOURCE
 * LINES, compiled from /root/reference by `make -C oracle ref` (oracle/ref_glue.cpp; OpenCV algorithms they call run in
 * the real OpenCV through callbacks, filter2D outputs are injected) -- tests/test_oracle_vs_reference.py,
 * tests/test_cost_builders.py and the committed vectors under tests/golden/ they produced.  The correlation itself
 * (cv::filter2D) is pinned against cv2.filter2D directly (tests/test_oracle_vs_cv2.py).
 *
 * Build: g++ -O3 -std=c++17 -ffp-contract=off -fPIC -shared  (see oracle/Makefile).
 * -ffp-contract=off matters: the mul+add correlation mode must round twice.
 */
#include "lm_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define LMO_CLONES __attribute__((target_clones("default", "fma", "avx512f")))
#else
#define LMO_CLONES
#endif

using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) {
    return std::chrono::duration<double>(b - a).count();
}

// ---------------------------------------------------------------------------------------------
// Geometry — LocoMouse_Model ctor class.cpp:3157-3161, the move-assignment quirk class.cpp:3164-3179
// (spost_b becomes spre_b because the model is move-assigned at class.cpp:323),
// initializeFeatureLoop class.cpp:672-697.
// ---------------------------------------------------------------------------------------------
struct Geom {
    int spre_b_w, spre_b_h, spost_b_w, spost_b_h;
    int spre_s_w, spre_s_h, spost_s_w, spost_s_h;
    int pad_pre_cols, pad_pre_rows, pad_post_cols, pad_post_rows;
};

int ceil_half(int v) { return (int)std::ceil((double)v / 2.0); }

Geom make_geom(const lm_config &c, const lm_template t[2][3]) {
    Geom g;
    const lm_template &pb = t[LM_BOTTOM][LM_PAW], &sb = t[LM_BOTTOM][LM_SNOUT];
    const lm_template &ps = t[LM_SIDE][LM_PAW], &ss = t[LM_SIDE][LM_SNOUT];
    int mbw = std::max(pb.cols, sb.cols) - 1, mbh = std::max(pb.rows, sb.rows) - 1;
    int msw = std::max(ps.cols, ss.cols) - 1, msh = std::max(ps.rows, ss.rows) - 1;
    g.spre_b_w = ceil_half(mbw);
    g.spre_b_h = ceil_half(mbh);
    g.spre_s_w = ceil_half(msw);
    g.spre_s_h = ceil_half(msh);
    g.spost_s_w = msw / 2;
    g.spost_s_h = msh / 2;
    // Q11: after `M = LocoMouse_Model(MODEL_FILE)` spost_b == spre_b
    g.spost_b_w = g.spre_b_w;
    g.spost_b_h = g.spre_b_h;
    g.pad_pre_rows = std::max(c.bb_h_side, std::max(g.spre_s_h, g.spre_b_h));
    g.pad_post_rows = std::max(g.spost_b_h, g.spost_s_h);
    g.pad_pre_cols = std::max(c.bb_w, std::max(g.spre_s_w, g.spre_b_w));
    g.pad_post_cols = std::max(g.spost_b_w, g.spost_s_w);
    return g;
}

// cropBoundingBox class.cpp:1422-1423,1457-1458 + cv::Mat ROI assertion (0<=x, x+w<=cols ...)
bool roi_ok(const lm_config &c, const Geom &g, uint32_t bbx, uint32_t bbys, uint32_t bbyb) {
    const int64_t canvas_cols = (int64_t)g.pad_pre_cols + c.n_cols + g.pad_post_cols;
    const int64_t canvas_rows = (int64_t)g.pad_pre_rows + c.n_rows + g.pad_post_rows;
    {
        int64_t W = g.spre_b_w + c.bb_w + g.spost_b_w, H = g.spre_b_h + c.bb_h_bottom + g.spost_b_h;
        int64_t x = (int32_t)(bbx + (uint32_t)g.pad_pre_cols - (uint32_t)(W - g.spost_b_w) + 1u);
        int64_t y = (int32_t)(bbyb + (uint32_t)g.pad_pre_rows - (uint32_t)(H - g.spost_b_h) + 1u);
        if (x < 0 || y < 0 || x + W > canvas_cols || y + H > canvas_rows) return false;
    }
    {
        int64_t W = g.spre_s_w + c.bb_w + g.spost_s_w, H = g.spre_s_h + c.bb_h_side + g.spost_s_h;
        int64_t x = (int32_t)(bbx + (uint32_t)g.pad_pre_cols - (uint32_t)(W - g.spost_s_w) + 1u);
        int64_t y = (int32_t)(bbys + (uint32_t)g.pad_pre_rows - (uint32_t)(H - g.spost_s_h) + 1u);
        if (x < 0 || y < 0 || x + W > canvas_cols || y + H > canvas_rows) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// imadjust LUT — LocoMouse::imadjust class.cpp:3204-3242 (round() = half away from zero)
// ---------------------------------------------------------------------------------------------
void imadjust_lut(double low_in, double high_in, double low_out, double high_out, uint8_t lut[256]) {
    low_in *= 255;
    high_in *= 255;
    low_out *= 255;
    high_out *= 255;
    double range_in = high_in - low_in, range_out = high_out - low_out, range_div = range_out / range_in;
    for (int i = 0; i < 256; ++i) {
        double temp;
        if (i <= low_in)
            temp = 0;
        else if (i >= high_in)
            temp = range_out;
        else
            temp = (i - low_in) * range_div;
        lut[i] = (uint8_t)std::round(temp + low_out);
    }
}

inline uint8_t sat_u8_rint(float v) {  // cv::saturate_cast<uchar>(float): cvRound (half-even) + clip
    long r = lrintf(v);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// ---------------------------------------------------------------------------------------------
// readFrame — class.cpp:1273-1333 : subtract (1304, saturating u8), normalize MINMAX to [0,255]
// (1310; cv::normalize + convertTo semantics: scale/shift in double, applied in float with one
// fused rounding, then round-half-even + saturate; verified against cv2 4.13), correctImage gather
// (1337-1406: Iout[x] = Iin[CALIBRATION[x]]), flip(I,I,1) (1323-1324); LocoMouse_TM::readFrame
// (LocoMouse_TM.cpp:243-249) adds imadjust(I,I,0,0.6,0,1).
// ---------------------------------------------------------------------------------------------
void preprocess(const lm_config &c, const uint8_t *bkg, const int32_t *calib, const uint8_t *frame,
                uint8_t *I, int32_t *minmax) {
    const int64_t nraw = (int64_t)c.vid_rows * c.vid_cols;
    std::vector<uint8_t> d(nraw);
    int smin = 255, smax = 0;
    for (int64_t p = 0; p < nraw; ++p) {
        int v = (int)frame[p] - (int)bkg[p];
        v = v < 0 ? 0 : v;
        d[p] = (uint8_t)v;
        smin = std::min(smin, v);
        smax = std::max(smax, v);
    }
    if (minmax) {
        minmax[0] = smin;
        minmax[1] = smax;
    }
    const double dmin = 0.0, dmax = 255.0;
    double scale = (dmax - dmin) * (((double)smax - (double)smin) > DBL_EPSILON ? 1.0 / ((double)smax - (double)smin) : 0.0);
    double shift = dmin - (double)smin * scale;
    const float a = (float)scale, b = (float)shift;
    uint8_t nlut[256];
    for (int v = 0; v < 256; ++v) nlut[v] = sat_u8_rint(__builtin_fmaf((float)v, a, b));
    uint8_t alut[256];
    if (c.imadjust)
        imadjust_lut(0, 0.6, 0, 1, alut);
    else
        for (int v = 0; v < 256; ++v) alut[v] = (uint8_t)v;
    for (int r = 0; r < c.n_rows; ++r) {
        const int32_t *pc = calib + (int64_t)r * c.n_cols;
        uint8_t *po = I + (int64_t)r * c.n_cols;
        for (int x = 0; x < c.n_cols; ++x) {
            int xs = c.flip ? (c.n_cols - 1 - x) : x;
            po[x] = alut[nlut[d[pc[xs]]]];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Template correlation — cv::filter2D(src ROI of I_PAD, CV_32F, K, Point(-1,-1), -rho,
// BORDER_CONSTANT) at class.cpp:845, 860, 2575-2576.  Semantics restated: correlation
// (no kernel flip), anchor (cols/2, rows/2), source ROI not isolated -> out-of-ROI taps read the
// surrounding canvas, which is the calibrated image zero-extended; float accumulator initialised to
// float(-rho); taps visited row-major.  fma_mode 0: two roundings per tap (bit-exact with OpenCV's
// direct filter engine, checked with cv2 for kernels it does not send to the DFT path);
// fma_mode 1: one rounding per tap (the mode the CUDA kernels are benchmarked in).
// ---------------------------------------------------------------------------------------------
// XB outputs of one row are kept in registers while the taps are visited row-major; every output
// sees exactly the same operation sequence as in a scalar loop, so blocking does not change results.
constexpr int XB = 32;

template <bool FMA>
__attribute__((always_inline)) static inline void correlate_rows(const float *patch, int pw, const float *K, int kh, int kw, float init, int w,
                                  int h, float *scores) {
    for (int Y = 0; Y < h; ++Y) {
        for (int X0 = 0; X0 < w; X0 += XB) {
            float acc[XB];
            for (int x = 0; x < XB; ++x) acc[x] = init;
            for (int j = 0; j < kh; ++j) {
                const float *row = patch + (int64_t)(Y + j) * pw + X0;
                const float *kr = K + j * kw;
                for (int i = 0; i < kw; ++i) {
                    const float wv = kr[i];
                    const float *src = row + i;
                    if (FMA) {
                        for (int x = 0; x < XB; ++x) acc[x] = __builtin_fmaf(wv, src[x], acc[x]);
                    } else {
                        for (int x = 0; x < XB; ++x) {
                            float prod = wv * src[x];
                            acc[x] = acc[x] + prod;
                        }
                    }
                }
            }
            const int nx = std::min(XB, w - X0);
            for (int x = 0; x < nx; ++x) scores[(int64_t)Y * w + X0 + x] = acc[x];
        }
    }
}

LMO_CLONES
void correlate_patch(const float *patch, int pw, const float *K, int kh, int kw, float init, int w, int h,
                     int fma_mode, float *scores) {
    if (fma_mode)
        correlate_rows<true>(patch, pw, K, kh, kw, init, w, h, scores);
    else
        correlate_rows<false>(patch, pw, K, kh, kw, init, w, h, scores);
}

void correlate(const uint8_t *I, int n_rows, int n_cols, const lm_template &t, int x0, int y0, int w, int h,
               int fma_mode, float *scores) {
    const int kh = t.rows, kw = t.cols, ay = kh / 2, ax = kw / 2;
    const int pw_real = w + kw - 1, ph = h + kh - 1;
    const int pw = ((w + XB - 1) / XB) * XB + kw - 1;  // zero columns so full XB blocks can be read
    std::vector<float> patch((size_t)pw * ph, 0.f);
    for (int r = 0; r < ph; ++r) {
        int yy = y0 + r - ay;
        if (yy < 0 || yy >= n_rows) continue;
        for (int q = 0; q < pw_real; ++q) {
            int xx = x0 + q - ax;
            if (xx >= 0 && xx < n_cols) patch[(size_t)r * pw + q] = (float)I[(int64_t)yy * n_cols + xx];
        }
    }
    const float init = (float)(-t.rho);
    correlate_patch(patch.data(), pw, t.w, kh, kw, init, w, h, fma_mode, scores);
}

// ---------------------------------------------------------------------------------------------
// Positive detections + ordering — class.cpp:1638-1658 and 1776-1800: row-major scan for score>0,
// then std::sort(compareCandidate) = score descending (Candidates.cpp:33-36).  std::sort is
// unstable for ties; the total order fixed here (and on the GPU) is
// (score desc, row-major pixel index asc) == stable sort of the row-major scan.
// ---------------------------------------------------------------------------------------------
struct Det {
    int x, y;
    double s;
};

std::vector<Det> collect_sorted(const float *scores, int rows, int cols) {
    std::vector<Det> d;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            float v = scores[(int64_t)r * cols + c];
            if (v > 0) d.push_back({c, r, (double)v});
        }
    std::stable_sort(d.begin(), d.end(), [](const Det &a, const Det &b) { return a.s > b.s; });
    return d;
}

// cv::Rect intersection area of two (w x h) boxes anchored at the detections
inline int inter_area(const Det &a, const Det &b, int w, int h) {
    int iw = std::min(a.x, b.x) + w - std::max(a.x, b.x);
    int ih = std::min(a.y, b.y) + h - std::max(a.y, b.y);
    if (iw <= 0 || ih <= 0) return 0;
    return iw * ih;
}

// cv::saturate_cast<int>(double) == cvRound == round-half-to-even (Point_<double> -> Point_<int>)
inline int round_half_even(double v) { return (int)std::nearbyint(v); }

// ---------------------------------------------------------------------------------------------
// nmsMax — class.cpp:1610-1747.  Chain suppression: the outer loop visits EVERY detection in rank
// order, discarded ones included (no `continue`); a later, not yet discarded detection j
// overlapping i by  inter/(2wh-inter) > 0.5  is discarded and inherits maxima_index[i].  Weighted
// mean over all detections in rank order in double (1731-1739); point = Point_<double>/sum converted
// to Point_<int> with round-half-even (1743); score = score of the root.
// ---------------------------------------------------------------------------------------------
std::vector<lm_cand> nms_max(const float *scores, int rows, int cols, int bw, int bh) {
    std::vector<lm_cand> out;
    std::vector<Det> det = collect_sorted(scores, rows, cols);
    const size_t N = det.size();
    if (!N) return out;
    const double area2 = 2.0 * (double)(bw * bh);
    std::vector<uint32_t> maxima_index(N);
    std::vector<uint32_t> cand_slot(N, 0);  // maxima_index_mapping
    std::vector<uint32_t> candidate_index;
    std::vector<char> discard(N, 0);
    for (size_t i = 0; i < N; ++i) {
        if (!discard[i]) {
            cand_slot[i] = (uint32_t)candidate_index.size();
            candidate_index.push_back((uint32_t)i);
            maxima_index[i] = (uint32_t)i;
        }
        for (size_t j = i + 1; j < N; ++j) {
            if (discard[j]) continue;
            int ia = inter_area(det[i], det[j], bw, bh);
            if (ia == 0) continue;
            double criterion = ia / (area2 - ia);
            if (criterion > 0.5) {
                discard[j] = 1;
                maxima_index[j] = maxima_index[i];
            }
        }
    }
    const size_t NC = candidate_index.size();
    std::vector<double> wx(NC, 0.0), wy(NC, 0.0), ss(NC, 0.0);
    for (size_t i = 0; i < N; ++i) {
        uint32_t slot = cand_slot[maxima_index[i]];
        wx[slot] += (double)det[i].x * det[i].s;
        wy[slot] += (double)det[i].y * det[i].s;
        ss[slot] += det[i].s;
    }
    out.resize(NC);
    for (size_t k = 0; k < NC; ++k) {
        out[k].x = round_half_even(wx[k] / ss[k]);
        out[k].y = round_half_even(wy[k] / ss[k]);
        out[k].s = det[candidate_index[k]].s;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// peakClustering — class.cpp:1749-1905 (called with cluster_method 3, overlap 0.5 at 868).
// Greedy: already clustered detections are skipped (1817); cluster point = round() (half away
// from zero, 1883) of the double weighted mean accumulated in cluster (= rank) order; score = the
// maximum's score; singletons are copied (1887).
// ---------------------------------------------------------------------------------------------
std::vector<lm_cand> peak_clustering(const float *scores, int rows, int cols, int bw, int bh) {
    std::vector<lm_cand> out;
    std::vector<Det> det = collect_sorted(scores, rows, cols);
    const size_t N = det.size();
    if (!N) return out;
    const double area2 = 2.0 * (double)bh * (double)bw;
    std::vector<char> kp(N, 0);
    std::vector<size_t> cluster;
    for (size_t i = 0; i < N; ++i) {
        if (kp[i]) continue;
        cluster.clear();
        cluster.push_back(i);
        for (size_t j = i + 1; j < N; ++j) {
            if (kp[j]) continue;
            int ia = inter_area(det[i], det[j], bw, bh);
            if (ia == 0) continue;
            double criterion = ia / (area2 - ia);
            if (criterion > 0.5) {
                kp[j] = 1;
                cluster.push_back(j);
            }
        }
        lm_cand c;
        if (cluster.size() > 1) {
            double px = 0, py = 0, sum = 0;
            for (size_t k : cluster) {
                px += (double)det[k].x * det[k].s;
                py += (double)det[k].y * det[k].s;
                sum += det[k].s;
            }
            c.x = (int)std::round(px / sum);
            c.y = (int)std::round(py / sum);
            c.s = det[i].s;
        } else {
            c.x = det[i].x;
            c.y = det[i].y;
            c.s = det[i].s;
        }
        out.push_back(c);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// selectLargestRegion — class.cpp:2744-2767 : cv::connectedComponentsWithStats(conn) then the
// label with the largest area, strict '>' so ties keep the lowest label (2751-2756); no foreground
// -> zeros (2762-2764).  OpenCV numbers labels by the scan position at which a component is first
// met: pixel raster order for its 4-connectivity scan (SAUF), 2x2-BLOCK raster order for its
// 8-connectivity scan (BBDT / Spaghetti); restated here as the tie-break key and pinned against
// cv2 in tests/test_oracle_vs_cv2.py.
// ---------------------------------------------------------------------------------------------
void largest_region(const uint8_t *bin, int rows, int cols, int conn, uint8_t *out) {
    const int64_t n = (int64_t)rows * cols;
    std::vector<int32_t> label(n, -1);
    std::vector<int32_t> stack;
    int32_t best_label = -1;
    int64_t best_area = 0, best_key = 0;
    const int bcols = (cols + 1) / 2;
    int32_t next = 0;
    for (int64_t p0 = 0; p0 < n; ++p0) {
        if (!bin[p0] || label[p0] >= 0) continue;
        int32_t lab = next++;
        int64_t area = 0, key = INT64_MAX;
        stack.clear();
        stack.push_back((int32_t)p0);
        label[p0] = lab;
        while (!stack.empty()) {
            int32_t p = stack.back();
            stack.pop_back();
            ++area;
            int r = p / cols, c = p % cols;
            int64_t k = (conn == 8) ? (int64_t)(r >> 1) * bcols + (c >> 1) : (int64_t)p;
            key = std::min(key, k);
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    if (!dr && !dc) continue;
                    if (conn != 8 && dr && dc) continue;
                    int rr = r + dr, cc = c + dc;
                    if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
                    int64_t q = (int64_t)rr * cols + cc;
                    if (bin[q] && label[q] < 0) {
                        label[q] = lab;
                        stack.push_back((int32_t)q);
                    }
                }
        }
        if (best_label < 0 || area > best_area || (area == best_area && key < best_key)) {
            best_label = lab;
            best_area = area;
            best_key = key;
        }
    }
    for (int64_t p = 0; p < n; ++p) out[p] = (best_label >= 0 && label[p] == best_label) ? 255 : 0;
}

// ---------------------------------------------------------------------------------------------
// detectLineCandidates after the two filter2D calls — class.cpp:2593-2742:
// bottom binary (>0) -> largest region = TAIL_MASK; its column-max masks the side binary; side
// largest region; bottom extent [first,last) split in n_points segments (first `remainder` ones one
// pixel wider); per segment centroid by cv::moments(binary) with (int) truncation; side z only when
// x > 0 (2729).
// ---------------------------------------------------------------------------------------------
void tail_from_binary(const uint8_t *bin_b, int rows_b, const uint8_t *bin_s, int rows_s, int cols, int conn,
                      int n_points, int32_t *tracks, uint8_t *tail_mask) {
    for (int i = 0; i < 3 * n_points; ++i) tracks[i] = -1;
    largest_region(bin_b, rows_b, cols, conn, tail_mask);
    std::vector<uint8_t> colmax(cols, 0);
    for (int r = 0; r < rows_b; ++r)
        for (int c = 0; c < cols; ++c) colmax[c] = std::max(colmax[c], tail_mask[(int64_t)r * cols + c]);
    std::vector<uint8_t> side_in((size_t)rows_s * cols), side_mask((size_t)rows_s * cols);
    for (int r = 0; r < rows_s; ++r)
        for (int c = 0; c < cols; ++c)
            side_in[(size_t)r * cols + c] = (bin_s[(size_t)r * cols + c] ? 255 : 0) & colmax[c];
    largest_region(side_in.data(), rows_s, cols, conn, side_mask.data());

    int first = -1;
    for (int i = 0; i < cols; ++i)
        if (colmax[i] > 0) {
            first = i;
            break;
        }
    if (first < 0) return;
    int last = first;
    for (int i = cols - 1; i > first; --i)
        if (colmax[i] > 0) {
            last = i;
            break;
        }
    const int tail_width = last - first;
    const int remainder = tail_width % n_points;
    const int regular = (tail_width - remainder) / n_points;
    int32_t *tx = tracks, *ty = tracks + n_points, *tz = tracks + 2 * n_points;
    int segx = first;
    for (int i = 0; i < n_points; ++i) {
        int segw = regular + (i < remainder ? 1 : 0);
        double m00 = 0, m10 = 0, m01 = 0;
        for (int r = 0; r < rows_b; ++r)
            for (int c = 0; c < segw; ++c)
                if (tail_mask[(int64_t)r * cols + segx + c]) {
                    m00 += 1;
                    m10 += c;
                    m01 += r;
                }
        if (m00 > 0) {
            tx[i] = (int)(m10 / m00) + segx;
            ty[i] = (int)(m01 / m00);
        }
        segx += segw;
    }
    for (int i = 0; i < n_points; ++i) {
        if (tx[i] > 0) {
            double m00 = 0, m01 = 0;
            for (int r = 0; r < rows_s; ++r)
                if (side_mask[(size_t)r * cols + tx[i]]) {
                    m00 += 1;
                    m01 += r;
                }
            if (m00 > 0) tz[i] = (int)(m01 / m00);
        }
    }
}

// zero-extended read of a calibrated image (the padded canvas I_PAD is zeros outside the image,
// class.cpp:684-689)
inline int px_at(const uint8_t *I, int n_rows, int n_cols, int x, int y) {
    if (x < 0 || y < 0 || x >= n_cols || y >= n_rows) return 0;
    return I[(int64_t)y * n_cols + x];
}

// checkVelCriterion — class.cpp:1256-1267: count of saturating (cur - prev) > 25 over the window,
// compared with (full template area) * alpha.
bool check_vel(const uint8_t *I, const uint8_t *Ip, int n_rows, int n_cols, int wx, int wy, int ww, int wh,
               int box_area, double alpha) {
    int cnt = 0;
    for (int r = 0; r < wh; ++r)
        for (int c = 0; c < ww; ++c) {
            int d = px_at(I, n_rows, n_cols, wx + c, wy + r) - px_at(Ip, n_rows, n_cols, wx + c, wy + r);
            if (d < 0) d = 0;
            if (d > 25) ++cnt;
        }
    return (double)cnt >= ((double)box_area) * alpha;
}

struct MatchBox {
    int tlx, tly, w, h;
};
// LocoMouse_Feature ctor class.cpp:2954-2969: half-size window, round() half away from zero
MatchBox match_box(int tw, int th) {
    MatchBox m;
    m.w = (int)std::round((double)tw / 2);
    m.h = (int)std::round((double)th / 2);
    m.tlx = -(m.w / 2);
    m.tly = -(m.h / 2);
    return m;
}

// ---------------------------------------------------------------------------------------------
// matchingWithVelocityConstraint + xDist + matchViews — class.cpp:1023-1254.
//  ovlp = (int)(w_bottom * (1 - T)) (1047); D = |xb - xs| (1075-1107); boolD = D <= ovlp followed by
//  cv::normalize(boolD, 0, 1, NORM_MINMAX) (1064-1065) which maps an all-equal matrix to all zeros
//  (SURVEY Q7); weight = 1 - D/ovlp evaluated as the cv::MatExpr  D * (-(1/ovlp)) + 1 in double
//  (1069: Mat / double multiplies by the reciprocal); reductions of boolD (1144-1150); per bottom
//  candidate the side candidates are visited in order; `colsum > 1 & vel_check` (1207, Q8).
// ---------------------------------------------------------------------------------------------
// Branch-coverage counters of the pairing stage (tests use them to prove that a parity input exercises the quirks):
// 0 pairings with both lists non-empty, 1 all-equal boolD zeroed by normalize (Q7), 2 velocity comparisons (Q8),
// 3 side candidates rejected by the velocity criterion, 4 accepted after a velocity comparison, 5 windows found moving,
// 6 bottom candidates left without a side match, 7 side matches emitted.
std::atomic<long long> g_cov[8];

int match_views(const lm_cand *cb, int nb, const lm_cand *cs, int ns, int vel_check, int tw_b, int th_b,
                int tw_s, int th_s, double T, const uint8_t *I, const uint8_t *Ip, int n_rows, int n_cols,
                int x0, int y0b, int y0s, int32_t *match_n, int32_t *match_y, double *match_s, int match_cap,
                int32_t *n_match) {
    int total = 0, overflow = 0;
    const int ovlp = (int)(tw_b * (1 - T));
    if (nb == 0) {
        *n_match = 0;
        return 0;
    }
    std::vector<uint8_t> boolD((size_t)nb * std::max(ns, 1), 0);
    std::vector<double> wgt((size_t)nb * std::max(ns, 1), 0.0);
    std::vector<float> colsum(std::max(ns, 1), 0.f), rowsum(nb, 0.f);
    if (ns > 0) {
        int mn = 255, mx = 0;
        const double alpha = -(1.0 / (double)ovlp);
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < ns; ++j) {
                int D = std::abs(cb[i].x - cs[j].x);
                uint8_t b = (D <= ovlp) ? 255 : 0;
                boolD[(size_t)i * ns + j] = b;
                mn = std::min<int>(mn, b);
                mx = std::max<int>(mx, b);
                double prod = (double)D * alpha;
                wgt[(size_t)i * ns + j] = prod + 1.0;
            }
        ++g_cov[0];
        if (mx == mn && mx > 0) ++g_cov[1];
        // normalize(boolD, boolD, 0, 1, NORM_MINMAX): scale = 1/(max-min) or 0 when max == min
        for (size_t k = 0; k < (size_t)nb * ns; ++k) boolD[k] = (mx > mn) ? (boolD[k] ? 1 : 0) : 0;
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < ns; ++j) {
                colsum[j] += boolD[(size_t)i * ns + j];
                rowsum[i] += boolD[(size_t)i * ns + j];
            }
    }
    const MatchBox mb = match_box(tw_b, th_b), ms = match_box(tw_s, th_s);
    std::vector<char> need_t(std::max(ns, 1), 1), mov_t(std::max(ns, 1), 0);
    for (int i = 0; i < nb; ++i) {
        int cnt = 0;
        if (ns > 0 && rowsum[i] != 0) {
            bool moving_b = false, need_b = true;
            for (int j = 0; j < ns; ++j) {
                if (boolD[(size_t)i * ns + j] < 1) continue;
                bool match = true;
                if ((colsum[j] > 1) & (vel_check != 0)) {
                    if (need_b) {
                        moving_b = check_vel(I, Ip, n_rows, n_cols, x0 + cb[i].x + mb.tlx, y0b + cb[i].y + mb.tly,
                                             mb.w, mb.h, tw_b * th_b, 0.02);
                        need_b = false;
                        if (moving_b) ++g_cov[5];
                    }
                    if (need_t[j]) {
                        mov_t[j] = check_vel(I, Ip, n_rows, n_cols, x0 + cs[j].x + ms.tlx, y0s + cs[j].y + ms.tly,
                                             ms.w, ms.h, tw_s * th_s, 0.05);
                        need_t[j] = 0;
                        if (mov_t[j]) ++g_cov[5];
                    }
                    match = (moving_b == (bool)mov_t[j]);
                    ++g_cov[2];
                    ++g_cov[match ? 4 : 3];
                }
                if (match) {
                    if (total < match_cap) {
                        match_y[total] = cs[j].y;
                        match_s[total] = cs[j].s * wgt[(size_t)i * ns + j];
                    } else
                        overflow = 1;
                    ++total;
                    ++cnt;
                }
            }
        }
        match_n[i] = cnt;
        if (!cnt) ++g_cov[6];
        g_cov[7] += cnt;
    }
    *n_match = total;
    return overflow;
}

// ---------------------------------------------------------------------------------------------
// One frame of the hot loop, reference main.cpp:54-82 (host cost builders excluded).
// ---------------------------------------------------------------------------------------------
struct Scratch {
    std::vector<uint8_t> I, Iprev;
    std::vector<float> sc;
    std::vector<uint8_t> bin_b, bin_s, tail_mask, mask;
};

void detect_frame(const lm_config &c, const lm_template t[2][3], const Geom &g, const uint8_t *bkg,
                  const int32_t *calib, const uint8_t *frame, const uint8_t *prev, bool vel_check,
                  uint32_t bbx, uint32_t bbys, uint32_t bbyb, int64_t f, lm_results *out, Scratch &S,
                  double *st) {
    (void)g;
    const int cap = out->cand_cap, mcap = out->match_cap, ntp = out->n_tail_points;
    auto t0 = clk::now();
    S.I.resize((size_t)c.n_rows * c.n_cols);
    preprocess(c, bkg, calib, frame, S.I.data(), nullptr);
    if (vel_check) {
        S.Iprev.resize(S.I.size());
        preprocess(c, bkg, calib, prev, S.Iprev.data(), nullptr);
    }
    auto t1 = clk::now();
    const int W = c.bb_w, Hb = c.bb_h_bottom, Hs = c.bb_h_side, TW = c.tail_w;
    // unpadded crop origins in image coordinates (class.cpp:1422-1423, 696-697)
    const int x0 = (int)bbx - W + 1, y0b = (int)bbyb - Hb + 1, y0s = (int)bbys - Hs + 1;
    const uint8_t *I = S.I.data();
    double t_corr = 0, t_tail = 0, t_nms = 0, t_pair = 0;

    // ---- detectTail (class.cpp:2541-2742)
    auto a = clk::now();
    S.sc.resize((size_t)std::max(Hb, Hs) * W);
    S.bin_b.assign((size_t)Hb * TW, 0);
    S.bin_s.assign((size_t)Hs * TW, 0);
    correlate(I, c.n_rows, c.n_cols, t[LM_BOTTOM][LM_TAIL], x0, y0b, TW, Hb, c.fma_mode, S.sc.data());
    for (size_t k = 0; k < (size_t)Hb * TW; ++k) S.bin_b[k] = S.sc[k] > 0 ? 1 : 0;
    correlate(I, c.n_rows, c.n_cols, t[LM_SIDE][LM_TAIL], x0, y0s, TW, Hs, c.fma_mode, S.sc.data());
    for (size_t k = 0; k < (size_t)Hs * TW; ++k) S.bin_s[k] = S.sc[k] > 0 ? 1 : 0;
    auto b = clk::now();
    t_corr += secs(a, b);
    S.tail_mask.resize((size_t)Hb * TW);
    tail_from_binary(S.bin_b.data(), Hb, S.bin_s.data(), Hs, TW, c.conn, ntp, out->tail + f * 3 * ntp,
                     S.tail_mask.data());
    a = clk::now();
    t_tail += secs(b, a);

    // ---- detectBottomCandidates (class.cpp:771-807, 841-854): mask = (px <= 25) | TAIL_MASK
    S.mask.assign((size_t)Hb * W, 0);
    for (int r = 0; r < Hb; ++r)
        for (int x = 0; x < W; ++x) {
            int px = px_at(I, c.n_rows, c.n_cols, x0 + x, y0b + r);
            uint8_t m = px <= 25 ? 255 : 0;  // threshold(.., 25.5, 255, THRESH_BINARY_INV)
            if (x < TW && S.tail_mask[(size_t)r * TW + x]) m = 255;
            S.mask[(size_t)r * W + x] = m;
        }
    std::vector<lm_cand> cb[2], cs[2];
    for (int k = 0; k < 2; ++k) {
        a = clk::now();
        correlate(I, c.n_rows, c.n_cols, t[LM_BOTTOM][k], x0, y0b, W, Hb, c.fma_mode, S.sc.data());
        for (size_t q = 0; q < (size_t)Hb * W; ++q)
            if (S.mask[q]) S.sc[q] = 0.f;
        b = clk::now();
        t_corr += secs(a, b);
        cb[k] = nms_max(S.sc.data(), Hb, W, t[LM_BOTTOM][k].cols, t[LM_BOTTOM][k].rows);
        a = clk::now();
        t_nms += secs(b, a);
    }
    // ---- detectSideCandidates (class.cpp:809-870): mask = px <= 25; skipped when the bottom list
    // of that feature is empty (820, 828)
    S.mask.assign((size_t)Hs * W, 0);
    for (int r = 0; r < Hs; ++r)
        for (int x = 0; x < W; ++x)
            S.mask[(size_t)r * W + x] = px_at(I, c.n_rows, c.n_cols, x0 + x, y0s + r) <= 25 ? 255 : 0;
    for (int k = 0; k < 2; ++k) {
        if (cb[k].empty()) continue;
        a = clk::now();
        correlate(I, c.n_rows, c.n_cols, t[LM_SIDE][k], x0, y0s, W, Hs, c.fma_mode, S.sc.data());
        for (size_t q = 0; q < (size_t)Hs * W; ++q)
            if (S.mask[q]) S.sc[q] = 0.f;
        b = clk::now();
        t_corr += secs(a, b);
        cs[k] = peak_clustering(S.sc.data(), Hs, W, t[LM_SIDE][k].cols, t[LM_SIDE][k].rows);
        a = clk::now();
        t_nms += secs(b, a);
    }
    // ---- matchBottomSideCandidates (class.cpp:999-1021)
    a = clk::now();
    uint32_t flags = 0;
    for (int k = 0; k < 2; ++k) {
        int nb = (int)cb[k].size(), ns = (int)cs[k].size();
        if (nb > cap || ns > cap) flags |= LM_FLAG_CAND_OVERFLOW;
        int nbc = std::min(nb, cap), nsc = std::min(ns, cap);
        out->n_bottom[f * 2 + k] = nbc;
        out->n_side[f * 2 + k] = nsc;
        lm_cand *ob = out->bottom + (f * 2 + k) * cap, *os = out->side + (f * 2 + k) * cap;
        for (int i = 0; i < cap; ++i) {
            ob[i] = (i < nbc) ? cb[k][i] : lm_cand{-1, -1, -1.0};
            os[i] = (i < nsc) ? cs[k][i] : lm_cand{-1, -1, -1.0};
        }
        int32_t *mn = out->match_n + (f * 2 + k) * cap;
        int32_t *my = out->match_y + (f * 2 + k) * mcap;
        double *msv = out->match_s + (f * 2 + k) * mcap;
        for (int i = 0; i < cap; ++i) mn[i] = 0;
        for (int i = 0; i < mcap; ++i) {
            my[i] = -1;
            msv[i] = -1.0;
        }
        int32_t nm = 0;
        int ov = match_views(cb[k].data(), nbc, cs[k].data(), nsc, vel_check ? 1 : 0, t[LM_BOTTOM][k].cols,
                             t[LM_BOTTOM][k].rows, t[LM_SIDE][k].cols, t[LM_SIDE][k].rows, c.min_overlap, I,
                             vel_check ? S.Iprev.data() : nullptr, c.n_rows, c.n_cols, x0, y0b, y0s, mn, my, msv,
                             mcap, &nm);
        if (ov) flags |= LM_FLAG_MATCH_OVERFLOW;
    }
    out->flags[f] = flags;
    b = clk::now();
    t_pair += secs(a, b);
    if (st) {
        st[0] += secs(t0, t1);
        st[1] += t_corr;
        st[2] += t_tail;
        st[3] += t_nms;
        st[4] += t_pair;
        st[5] += secs(t0, b);
    }
}


// ---------------------------------------------------------------------------------------------
// Pass 1, LocoMouse_TM_DE: LocoMouse_TM_DE.cpp:8-113.
// ---------------------------------------------------------------------------------------------
// imadjust_default (class.cpp:3244-3311): histogram -> first bin whose normalised cumulative count is > 0.01
// (imin) / >= 0.99 (imax), all in float as in the reference; ranges = index / 255 (float); then the in-place
// cv::MatExpr  Iout = (Iout - r0) / (r1 - r0)  which OpenCV evaluates as ONE convertTo(CV_8U, alpha, beta) with
// alpha = 1 / (double)(r1 - r0) and beta = -(double)r0 * alpha (MatOp_AddEx::multiply), i.e. per pixel
// saturate_cast<uchar>(fmaf(src, (float)alpha, (float)beta)) -- unless alpha == 1, where MatOp_AddEx::assign takes
// the cv::add(src, -0) branch and the image is unchanged.  (Pinned against cv2.calcHist / cv2.convertScaleAbs.)
void imadjust_default_lut(const uint32_t *hist, uint8_t *lut, int32_t *imin_imax) {
    const float min_tol = 0.01f, max_tol = 0.99f;
    float sum_histf = 0.f;
    {
        double acc = 0.0;  // cv::sum accumulates in double, the result is narrowed to float (class.cpp:3260-3261)
        for (int i = 0; i < 256; ++i) acc += (double)(float)hist[i];
        sum_histf = (float)acc;
    }
    float cumsum_step = 0.f;
    int indices[2] = {0, 0}, imin = 0, imax = 0;
    bool check_min = true, check_max = true;
    for (int i = 0; i < 256; ++i) {
        cumsum_step += (float)hist[i];
        const float cn = cumsum_step / sum_histf;
        if ((cn > min_tol) & check_min) {
            indices[0] = i;
            check_min = false;
            imin = i;
        }
        if ((cn >= max_tol) & check_max) {
            indices[1] = i;
            check_max = false;
            imax = i;
        }
        if (!(check_min || check_max)) break;
    }
    if (imin == imax) indices[1] = 256;
    const float r0 = (float)indices[0] / 255.f, r1 = (float)indices[1] / 255.f;
    const double s = (double)(r1 - r0);
    const double alpha = 1.0 / s, beta = -(double)r0 * alpha;
    if (imin_imax) {
        imin_imax[0] = indices[0];
        imin_imax[1] = indices[1];
    }
    if (std::fabs(alpha) == 1.0) {
        for (int v = 0; v < 256; ++v) lut[v] = (uint8_t)v;
        return;
    }
    const float a = (float)alpha, b = (float)beta;
    for (int v = 0; v < 256; ++v) lut[v] = sat_u8_rint(__builtin_fmaf((float)v, a, b));
}

// firstLastOverT<int> (class.hpp:411-442): the first qualifying index goes to slot 0, every later one to slot 1;
// with exactly one qualifying column slot 1 keeps its initial 0; none -> (-1, -1).
void first_last_over_t(const float *p, uint32_t L, int th, int32_t *first_last) {
    bool has_first = false;
    first_last[0] = 0;
    first_last[1] = 0;
    int index = 0;
    for (uint32_t i = 0; i < L; ++i)
        if (p[i] >= (float)th) {
            first_last[index] = (int32_t)i;
            if (!has_first) {
                index = 1;
                has_first = true;
            }
        }
    if (!has_first) first_last[0] = first_last[1] = -1;
}

// (uint32_t)double as x86-64 compilers evaluate it for the values that occur here (the conversion of a negative
// double is undefined behaviour in C++; cvttsd2si to 64 bits, low 32 bits kept).
inline uint32_t u32_from_double(double v) { return (uint32_t)(int64_t)v; }

// vecmovingaverage (class.cpp:1559-1608)
void vecmovingaverage(const double *v, int64_t n, int window, uint32_t *out) {
    if ((int64_t)window >= n) {
        for (int64_t i = 0; i < n; ++i) out[i] = u32_from_double(v[i]);
        return;
    }
    double current_sum = 0;
    const int half = window / 2;
    for (int i = 0; i < half; ++i) out[i] = u32_from_double(v[i]);
    for (int i = 0; i < window; ++i) current_sum += v[i];
    out[half] = u32_from_double(std::floor(current_sum / window));
    for (int64_t i = 0; i < n - window; ++i) {
        current_sum = current_sum - v[i] + v[i + window];
        out[half + 1 + i] = u32_from_double(std::floor(current_sum / window));
    }
    for (int64_t i = n - half - 1; i < n; ++i) out[i] = u32_from_double(v[i]);
}

// computeMouseBox_DE (LocoMouse_TM_DE.cpp:56-113) on the calibrated image of the BASE readFrame
void bounding_box_tm_de_frame(const lm_config &c, const uint8_t *bkg, const int32_t *calib, const uint8_t *frame,
                              const lm_bb_de_params &p, std::vector<uint8_t> &I, double *bb_x, int32_t *lims) {
    lm_config base = c;
    base.imadjust = 0;  // LocoMouse::readFrame(I), not LocoMouse_TM::readFrame (LocoMouse_TM_DE.cpp:36)
    preprocess(base, bkg, calib, frame, I.data(), nullptr);
    uint32_t hist[256] = {0};
    for (int r = 0; r < p.side_h; ++r) {
        const uint8_t *row = I.data() + (int64_t)(p.side_y + r) * c.n_cols + p.side_x;
        for (int x = 0; x < p.side_w; ++x) ++hist[row[x]];
    }
    uint8_t lut[256];
    imadjust_default_lut(hist, lut, nullptr);
    std::vector<float> colsum(p.side_w, 0.f);
    // colRange(0, pre) / colRange(post, N_COLS) / rowRange(0, pre) / rowRange(post, rows) set to zero (68-71),
    // threshold(> SIDE_THRESHOLD -> 1) (75), reduce(SUM over rows, CV_32F) (91)
    const int c0 = std::max(0, p.zero_col_pre), c1 = std::min(p.side_w, p.zero_col_post);
    const int r0 = std::max(0, p.zero_row_pre), r1 = std::min(p.side_h, p.zero_row_post);
    for (int r = r0; r < r1; ++r) {
        const uint8_t *row = I.data() + (int64_t)(p.side_y + r) * c.n_cols + p.side_x;
        for (int x = c0; x < c1; ++x)
            if ((double)lut[row[x]] > p.threshold) colsum[x] += 1.f;
    }
    int32_t fl[2];
    first_last_over_t(colsum.data(), (uint32_t)p.side_w, p.min_count, fl);
    if (lims) {
        lims[0] = fl[0];
        lims[1] = fl[1];
    }
    *bb_x = std::min((double)(p.side_w - 1), (double)fl[1] * p.width_margin);
}

// ---------------------------------------------------------------------------------------------
// Pass 1 of the base class — computeMouseBox (class.cpp:921-997) on the image the base readFrame produced.
//  * medianBlur(I_median, I_median, k) runs on the image padded by k/2 zeros (class.cpp:584-601); the padding stays zero
//    from frame to frame (a padding pixel's window holds at most k*(k-1)/2 image pixels, fewer than half), so the result
//    inside the image is the median over a zero-extended k x k window.
//  * threshold(I, I, 2.55, 1, THRESH_BINARY): 1 where the median is >= 3.
//  * largestBWAreaObject on each view: largest component, ties -> lowest OpenCV label (as selectLargestRegion), 0 / 255.
//  * reduce(SUM, CV_32S) along both axes, firstLastOverT<int>.  firstLastOverT reads through `const float *`
//    (class.hpp:417): with sums_as_float the int32 bit patterns are compared as floats, exactly what the reference does.
// ---------------------------------------------------------------------------------------------
void first_last_i32(const int32_t *sums, uint32_t L, int th, bool as_float, int32_t *fl) {
    std::vector<float> v(L);
    for (uint32_t i = 0; i < L; ++i) {
        if (as_float)
            std::memcpy(&v[i], &sums[i], 4);
        else
            v[i] = (float)sums[i];
    }
    first_last_over_t(v.data(), L, th, fl);
}

void mouse_box_base(const uint8_t *I, int n_rows, int n_cols, int conn, const lm_bb_base_params &p, double *box, int32_t *lims) {
    const int k = p.median_filter_size, h = k / 2, need = (k * k + 1) / 2;
    // median >= 3  <=>  at least (k*k+1)/2 of the k*k window values (zeros outside the image) are >= 3
    std::vector<uint8_t> bin((size_t)n_rows * n_cols);
    std::vector<int32_t> integral((size_t)(n_rows + 1) * (n_cols + 1), 0);
    for (int r = 0; r < n_rows; ++r)
        for (int c = 0; c < n_cols; ++c)
            integral[(size_t)(r + 1) * (n_cols + 1) + c + 1] = (I[(size_t)r * n_cols + c] >= 3) + integral[(size_t)r * (n_cols + 1) + c + 1] +
                                                                integral[(size_t)(r + 1) * (n_cols + 1) + c] - integral[(size_t)r * (n_cols + 1) + c];
    for (int r = 0; r < n_rows; ++r) {
        const int r0 = std::max(0, r - h), r1 = std::min(n_rows, r + h + 1);
        for (int c = 0; c < n_cols; ++c) {
            const int c0 = std::max(0, c - h), c1 = std::min(n_cols, c + h + 1);
            const int cnt = integral[(size_t)r1 * (n_cols + 1) + c1] - integral[(size_t)r0 * (n_cols + 1) + c1] -
                            integral[(size_t)r1 * (n_cols + 1) + c0] + integral[(size_t)r0 * (n_cols + 1) + c0];
            bin[(size_t)r * n_cols + c] = cnt >= need ? 1 : 0;
        }
    }
    int32_t fl[4][2];
    const int vx[2] = {p.side_x, p.bottom_x}, vy[2] = {p.side_y, p.bottom_y}, vw[2] = {p.side_w, p.bottom_w}, vh[2] = {p.side_h, p.bottom_h};
    for (int v = 0; v < 2; ++v) {
        std::vector<uint8_t> view((size_t)vw[v] * vh[v]), big(view.size());
        for (int r = 0; r < vh[v]; ++r)
            for (int c = 0; c < vw[v]; ++c) view[(size_t)r * vw[v] + c] = bin[(size_t)(vy[v] + r) * n_cols + vx[v] + c];
        largest_region(view.data(), vh[v], vw[v], conn, big.data());
        std::vector<int32_t> row(vw[v], 0), col(vh[v], 0);
        for (int r = 0; r < vh[v]; ++r)
            for (int c = 0; c < vw[v]; ++c) {
                row[c] += big[(size_t)r * vw[v] + c];
                col[r] += big[(size_t)r * vw[v] + c];
            }
        // firstLastOverT(Row_*, I.cols, ...) scans N_COLS entries of a vector that has view-width entries: identical when the
        // view spans the image width (the reference's calibration files); a narrower view is scanned over its own width
        first_last_i32(row.data(), (uint32_t)std::min(n_cols, vw[v]), p.min_pixel_visible, p.sums_as_float != 0, fl[v]);
        first_last_i32(col.data(), (uint32_t)vh[v], p.min_pixel_visible, p.sums_as_float != 0, fl[2 + v]);
    }
    const int32_t *rs = fl[0], *rb = fl[1], *cs = fl[2], *cb = fl[3];
    box[0] = rb[1] > rs[1] ? (double)rb[1] : (double)rs[1];
    box[1] = (double)cb[1] + (double)p.bottom_y;  // class.cpp:628: made absolute by the caller
    box[2] = (double)cs[1];
    const unsigned int wt = (unsigned int)(rs[1] - rs[0]), wb = (unsigned int)(rb[1] - rb[0]);
    box[3] = wt > wb ? (double)wt : (double)wb;
    box[4] = (double)(cb[1] - cb[0]);
    box[5] = (double)(cs[1] - cs[0]);
    if (lims) std::memcpy(lims, fl, sizeof fl);
}

// ---------------------------------------------------------------------------------------------
// Pass 1 of LocoMouse_TM — computeMouseBox_DD / bwAreaOpen / imfill (LocoMouse_TM.cpp:158-269) on the image the base
// readFrame produced.  Stage by stage (each stage is compared with what the reference's own compiled lines hand to
// OpenCV, tests/test_oracle_pass1_tm.py):
//  * imadjust_default on the side view (a 256-entry table, as for LocoMouse_TM_DE), four bands zeroed (203-206);
//  * threshold(I, ., SIDE_THRESHOLD, 1, THRESH_BINARY): 1 where the value is > the threshold (210);
//  * bwAreaOpen (158-187): components (4- / 8-connected) with area >= MIN_PIXEL_COUNT kept (255 / 255 -> 1); all zero when
//    there is no foreground label;
//  * filter2D(., CV_8UC1, DISK_FILTER, (-1,-1), 0, BORDER_REPLICATE) (216): anchor = size / 2, clamped coordinates, float
//    accumulation over the kernel in row-major order, saturate_cast<uchar> = round half to even (OpenCV's direct path);
//  * imfill (252-269): floodFill(copy, (0,0), 255) with the default zero tolerances and 4-connectivity reaches the pixels
//    connected to (0,0) through pixels of the same value; out = in | ~filled: a reached pixel keeps its value, every other
//    pixel becomes 255;
//  * reduce(SUM over rows, CV_32S) (221), firstLastOverT<int>(Row_side, N_COLS, ., min_pixel_visible) with the float read of
//    the integer sums (class.hpp:417) when sums_as_float; bb_x = last.
// ---------------------------------------------------------------------------------------------
bool tm_params_ok(const lm_bb_tm_params *p, int n_rows, int n_cols) {
    return p && p->disk && p->disk_size >= 1 && p->disk_size <= 63 && p->side_x >= 0 && p->side_y >= 0 && p->side_h > 0 &&
           p->side_x == 0 && p->side_w == n_cols &&  // colRange(ZERO_COL_POST, N_COLS) on the side view (TM.cpp:205)
           p->side_y + p->side_h <= n_rows && p->side_threshold >= 0 && p->side_threshold <= 255 && p->min_pixel_count >= 1 &&
           p->min_pixel_visible >= 0 && p->zero_col_pre >= 0 && p->zero_col_pre <= p->side_w && p->zero_col_post >= 0 &&
           p->zero_col_post <= n_cols && p->zero_row_pre >= 0 && p->zero_row_pre <= p->side_h && p->zero_row_post >= 0 &&
           p->zero_row_post <= p->side_h;
}

// filter2D(src, dst, CV_8UC1, kernel, (-1,-1), 0, BORDER_REPLICATE) on 8-bit data, OpenCV's direct path
void filter2d_u8_replicate(const uint8_t *src, int H, int W, const float *kern, int K, uint8_t *dst) {
    const int an = K / 2;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            float s = 0.f;
            for (int j = 0; j < K; ++j) {
                const int rr = std::min(H - 1, std::max(0, r + j - an));
                for (int i = 0; i < K; ++i) {
                    const int cc = std::min(W - 1, std::max(0, c + i - an));
                    s += kern[j * K + i] * (float)src[(size_t)rr * W + cc];   // -ffp-contract=off: multiply, then add
                }
            }
            dst[(size_t)r * W + c] = sat_u8_rint(s);
        }
}

struct TmStages {  // optional taps (tests)
    uint8_t *adjusted = nullptr, *binary = nullptr, *opened = nullptr, *filtered = nullptr;
    int32_t *row_sums = nullptr;
};

void mouse_box_tm(const uint8_t *I, int n_cols, int conn, const lm_bb_tm_params &p, double *bb_x, int32_t *lims, const TmStages &T) {
    const int W = p.side_w, H = p.side_h;
    const size_t N = (size_t)W * H;
    std::vector<uint8_t> a(N), b(N), o(N), f(N);
    uint32_t hist[256] = {0};
    for (int r = 0; r < H; ++r) {
        const uint8_t *row = I + (int64_t)(p.side_y + r) * n_cols + p.side_x;
        for (int x = 0; x < W; ++x) ++hist[row[x]];
    }
    uint8_t lut[256];
    imadjust_default_lut(hist, lut, nullptr);
    for (int r = 0; r < H; ++r) {
        const uint8_t *row = I + (int64_t)(p.side_y + r) * n_cols + p.side_x;
        const bool rz = r < p.zero_row_pre || r >= p.zero_row_post;
        for (int x = 0; x < W; ++x) {
            const bool z = rz || x < p.zero_col_pre || x >= p.zero_col_post;
            a[(size_t)r * W + x] = z ? 0 : lut[row[x]];
            b[(size_t)r * W + x] = a[(size_t)r * W + x] > p.side_threshold ? 1 : 0;
        }
    }
    // bwAreaOpen
    {
        std::vector<int32_t> label(N, -1), stack;
        std::vector<int64_t> area;
        for (size_t p0 = 0; p0 < N; ++p0) {
            if (!b[p0] || label[p0] >= 0) continue;
            const int32_t lab = (int32_t)area.size();
            area.push_back(0);
            stack.assign(1, (int32_t)p0);
            label[p0] = lab;
            while (!stack.empty()) {
                const int32_t q = stack.back();
                stack.pop_back();
                ++area[lab];
                const int r = q / W, c = q % W;
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        if ((!dr && !dc) || (conn != 8 && dr && dc)) continue;
                        const int rr = r + dr, cc = c + dc;
                        if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                        const size_t t = (size_t)rr * W + cc;
                        if (b[t] && label[t] < 0) {
                            label[t] = lab;
                            stack.push_back((int32_t)t);
                        }
                    }
            }
        }
        for (size_t q = 0; q < N; ++q) o[q] = (label[q] >= 0 && (uint64_t)area[label[q]] >= (uint64_t)(uint32_t)p.min_pixel_count) ? 1 : 0;
    }
    filter2d_u8_replicate(o.data(), H, W, p.disk, p.disk_size, f.data());
    // imfill + column sums
    std::vector<int32_t> sums(W, 0);
    {
        std::vector<uint8_t> reached(N, 0);
        std::vector<int32_t> stack(1, 0);
        const uint8_t seed = f[0];
        reached[0] = 1;
        while (!stack.empty()) {
            const int32_t q = stack.back();
            stack.pop_back();
            const int r = q / W, c = q % W;
            const int nr[4] = {r - 1, r + 1, r, r}, nc[4] = {c, c, c - 1, c + 1};
            for (int d = 0; d < 4; ++d) {
                if (nr[d] < 0 || nr[d] >= H || nc[d] < 0 || nc[d] >= W) continue;
                const size_t t = (size_t)nr[d] * W + nc[d];
                if (!reached[t] && f[t] == seed) {
                    reached[t] = 1;
                    stack.push_back((int32_t)t);
                }
            }
        }
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) sums[c] += reached[(size_t)r * W + c] ? (int32_t)seed : 255;
    }
    int32_t fl[2];
    // firstLastOverT(Row_side, N_COLS, ...): N_COLS == side_w (checked)
    first_last_i32(sums.data(), (uint32_t)W, p.min_pixel_visible, p.sums_as_float != 0, fl);
    if (lims) {
        lims[0] = fl[0];
        lims[1] = fl[1];
    }
    *bb_x = (double)fl[1];
    if (T.adjusted) std::memcpy(T.adjusted, a.data(), N);
    if (T.binary) std::memcpy(T.binary, b.data(), N);
    if (T.opened) std::memcpy(T.opened, o.data(), N);
    if (T.filtered) std::memcpy(T.filtered, f.data(), N);
    if (T.row_sums) std::memcpy(T.row_sums, sums.data(), (size_t)W * 4);
}

// medianvec / stdvec / computeMouseBoxSize (class.cpp:1481-1556).  medianvec sorts its argument; for an odd count it
// returns the element BELOW the middle (v[N/2 - 1]), as written there.  stdvec runs on the sorted data.
double medianvec(std::vector<double> &v) {
    const int N = (int)v.size();
    if (N == 1) return v[0];
    std::sort(v.begin(), v.end());
    const int half = N / 2;
    return N % 2 == 0 ? (v[half - 1] + v[half]) / 2 : v[half - 1];
}
double stdvec(const std::vector<double> &v) {
    const int N = (int)v.size();
    if (N == 1) return 0.0;
    double sum = 0.0;
    for (double x : v) sum += x;               // std::accumulate, left to right
    const double mean = sum / N;
    double sq = 0.0;
    for (double x : v) sq += (x - mean) * (x - mean);  // std::inner_product of the differences with themselves
    return std::sqrt(sq / (N - 1));
}
void mouse_box_size(double *w, double *hb, double *hs, int64_t n, int32_t size[3]) {
    double *series[3] = {w, hb, hs};
    for (int q = 0; q < 3; ++q) {
        std::vector<double> v(series[q], series[q] + n);
        const double med = medianvec(v);
        const double sd = stdvec(v);
        const uint32_t m3 = u32_from_double(med + 3 * sd);
        // `uint32_t final = (m3 < v[N-1]) ? m3 : v[N-1]`: the comparison and the conditional's value are double
        const double last = v[(size_t)n - 1];
        size[q] = (int32_t)u32_from_double((double)m3 < last ? (double)m3 : last);
        std::copy(v.begin(), v.end(), series[q]);
    }
}

int validate(const lm_config *c, const lm_template t[2][3]) {
    if (!c || !t) return LM_ERR_INVALID;
    if (c->vid_rows <= 0 || c->vid_cols <= 0 || c->n_rows <= 0 || c->n_cols <= 0) return LM_ERR_INVALID;
    if (c->bb_w <= 0 || c->bb_h_bottom <= 0 || c->bb_h_side <= 0) return LM_ERR_INVALID;
    if (c->tail_w < 0 || c->tail_w > c->bb_w) return LM_ERR_INVALID;
    if (c->conn != 4 && c->conn != 8) return LM_ERR_INVALID;
    if (c->n_tail_points <= 0) return LM_ERR_INVALID;
    for (int v = 0; v < 2; ++v)
        for (int k = 0; k < 3; ++k)
            if (!t[v][k].w || t[v][k].rows <= 0 || t[v][k].cols <= 0) return LM_ERR_INVALID;
    return LM_OK;
}

}  // namespace

extern "C" {

int lmo_bounding_box_tm_de(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames, int64_t n,
                           const lm_bb_de_params *p, double *bb_x, int32_t *lims) {
    if (!cfg || !bkg || !calib || !frames || !p || !bb_x || n < 0) return LM_ERR_INVALID;
    if (p->side_x < 0 || p->side_y < 0 || p->side_w <= 0 || p->side_h <= 0 || p->side_x + p->side_w > cfg->n_cols ||
        p->side_y + p->side_h > cfg->n_rows)
        return LM_ERR_INVALID;
    std::vector<uint8_t> I((size_t)cfg->n_rows * cfg->n_cols);
    const int64_t fsz = (int64_t)cfg->vid_rows * cfg->vid_cols;
    for (int64_t f = 0; f < n; ++f)
        bounding_box_tm_de_frame(*cfg, bkg, calib, frames + f * fsz, *p, I, bb_x + f, lims ? lims + 2 * f : nullptr);
    return LM_OK;
}
static bool base_params_ok(const lm_bb_base_params *p, int n_rows, int n_cols) {
    return p && p->median_filter_size >= 1 && (p->median_filter_size & 1) && p->median_filter_size <= 255 && p->min_pixel_visible >= 0 &&
           p->side_x >= 0 && p->side_y >= 0 && p->side_w > 0 && p->side_h > 0 && p->side_x + p->side_w <= n_cols && p->side_y + p->side_h <= n_rows &&
           p->bottom_x >= 0 && p->bottom_y >= 0 && p->bottom_w > 0 && p->bottom_h > 0 && p->bottom_x + p->bottom_w <= n_cols &&
           p->bottom_y + p->bottom_h <= n_rows;
}
int lmo_mouse_box_base(const uint8_t *I, int32_t n_rows, int32_t n_cols, int32_t conn, const lm_bb_base_params *p, double *box, int32_t *lims) {
    if (!I || !box || !base_params_ok(p, n_rows, n_cols) || (conn != 4 && conn != 8)) return LM_ERR_INVALID;
    mouse_box_base(I, n_rows, n_cols, conn, *p, box, lims);
    return LM_OK;
}
int lmo_bounding_box_base(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames, int64_t n,
                          const lm_bb_base_params *p, double *box, int32_t *lims) {
    if (!cfg || !bkg || !calib || !frames || !box || n < 0 || !base_params_ok(p, cfg->n_rows, cfg->n_cols)) return LM_ERR_INVALID;
    lm_config base = *cfg;
    base.imadjust = 0;  // LocoMouse::readFrame(I_center), class.cpp:622
    std::vector<uint8_t> I((size_t)cfg->n_rows * cfg->n_cols);
    const int64_t fsz = (int64_t)cfg->vid_rows * cfg->vid_cols;
    for (int64_t f = 0; f < n; ++f) {
        preprocess(base, bkg, calib, frames + f * fsz, I.data(), nullptr);
        mouse_box_base(I.data(), cfg->n_rows, cfg->n_cols, cfg->conn, *p, box + f * 6, lims ? lims + f * 8 : nullptr);
    }
    return LM_OK;
}
void lmo_filter2d_u8(const uint8_t *src, int32_t rows, int32_t cols, const float *kernel, int32_t k, uint8_t *dst) {
    filter2d_u8_replicate(src, rows, cols, kernel, k, dst);
}
int lmo_mouse_box_tm(const uint8_t *I, int32_t n_rows, int32_t n_cols, int32_t conn, const lm_bb_tm_params *p, double *bb_x, int32_t *lims,
                     uint8_t *adjusted, uint8_t *binary, uint8_t *opened, uint8_t *filtered, int32_t *row_sums) {
    if (!I || !bb_x || !tm_params_ok(p, n_rows, n_cols) || (conn != 4 && conn != 8)) return LM_ERR_INVALID;
    TmStages T;
    T.adjusted = adjusted;
    T.binary = binary;
    T.opened = opened;
    T.filtered = filtered;
    T.row_sums = row_sums;
    mouse_box_tm(I, n_cols, conn, *p, bb_x, lims, T);
    return LM_OK;
}
int lmo_bounding_box_tm(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames, int64_t n,
                        const lm_bb_tm_params *p, double *bb_x, int32_t *lims) {
    if (!cfg || !bkg || !calib || !frames || !bb_x || n < 0 || !tm_params_ok(p, cfg->n_rows, cfg->n_cols)) return LM_ERR_INVALID;
    lm_config base = *cfg;
    base.imadjust = 0;  // LocoMouse::readFrame(I), LocoMouse_TM.cpp:138
    std::vector<uint8_t> I((size_t)cfg->n_rows * cfg->n_cols);
    const int64_t fsz = (int64_t)cfg->vid_rows * cfg->vid_cols;
    for (int64_t f = 0; f < n; ++f) {
        preprocess(base, bkg, calib, frames + f * fsz, I.data(), nullptr);
        mouse_box_tm(I.data(), cfg->n_cols, cfg->conn, *p, bb_x + f, lims ? lims + 2 * f : nullptr, TmStages());
    }
    return LM_OK;
}
void lmo_mouse_box_size(double *w, double *hb, double *hs, int64_t n, int32_t size[3]) { mouse_box_size(w, hb, hs, n, size); }
void lmo_imadjust_default_lut(const uint32_t *hist, uint8_t *lut, int32_t *imin_imax) { imadjust_default_lut(hist, lut, imin_imax); }
void lmo_first_last_over_t(const float *values, uint32_t L, int32_t th, int32_t *first_last) { first_last_over_t(values, L, th, first_last); }
void lmo_vecmovingaverage(const double *v, int64_t n, int32_t window, uint32_t *out) { vecmovingaverage(v, n, window, out); }
void lmo_coverage(int64_t out[8], int reset) {
    for (int i = 0; i < 8; ++i) {
        out[i] = g_cov[i].load();
        if (reset) g_cov[i] = 0;
    }
}

int lmo_geometry(const lm_config *cfg, const lm_template t[2][3], int32_t pads[8], int32_t canvas[4]) {
    if (validate(cfg, t)) return LM_ERR_INVALID;
    Geom g = make_geom(*cfg, t);
    int32_t p[8] = {g.spre_b_w, g.spre_b_h, g.spost_b_w, g.spost_b_h, g.spre_s_w, g.spre_s_h, g.spost_s_w, g.spost_s_h};
    int32_t cv[4] = {g.pad_pre_cols, g.pad_pre_rows, g.pad_post_cols, g.pad_post_rows};
    if (pads) memcpy(pads, p, sizeof p);
    if (canvas) memcpy(canvas, cv, sizeof cv);
    return LM_OK;
}

int lmo_check_roi(const lm_config *cfg, const lm_template t[2][3], uint32_t bb_x, uint32_t bb_y_side,
                  uint32_t bb_y_bottom) {
    if (validate(cfg, t)) return LM_ERR_INVALID;
    Geom g = make_geom(*cfg, t);
    return roi_ok(*cfg, g, bb_x, bb_y_side, bb_y_bottom) ? LM_OK : LM_ERR_ROI;
}

int lmo_preprocess(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frame,
                   uint8_t *I, int32_t *minmax) {
    preprocess(*cfg, bkg, calib, frame, I, minmax);
    return LM_OK;
}

void lmo_imadjust_lut(double low_in, double high_in, double low_out, double high_out, uint8_t lut[256]) {
    imadjust_lut(low_in, high_in, low_out, high_out, lut);
}

void lmo_correlate(const uint8_t *I, int32_t n_rows, int32_t n_cols, const lm_template *t, int32_t x0,
                   int32_t y0, int32_t w, int32_t h, int32_t fma_mode, float *scores) {
    correlate(I, n_rows, n_cols, *t, x0, y0, w, h, fma_mode, scores);
}

int lmo_nms_max(const float *scores, int32_t rows, int32_t cols, int32_t box_w, int32_t box_h, lm_cand *out,
                int32_t cap) {
    std::vector<lm_cand> c = nms_max(scores, rows, cols, box_w, box_h);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) out[i] = c[i];
    return (int)c.size();
}

int lmo_peak_clustering(const float *scores, int32_t rows, int32_t cols, int32_t box_w, int32_t box_h,
                        lm_cand *out, int32_t cap) {
    std::vector<lm_cand> c = peak_clustering(scores, rows, cols, box_w, box_h);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) out[i] = c[i];
    return (int)c.size();
}

void lmo_largest_region(const uint8_t *bin, int32_t rows, int32_t cols, int32_t conn, uint8_t *out) {
    largest_region(bin, rows, cols, conn, out);
}

void lmo_tail_from_binary(const uint8_t *bin_bottom, int32_t rows_b, const uint8_t *bin_side, int32_t rows_s,
                          int32_t cols, int32_t conn, int32_t n_points, int32_t *tracks, uint8_t *tail_mask) {
    tail_from_binary(bin_bottom, rows_b, bin_side, rows_s, cols, conn, n_points, tracks, tail_mask);
}

// ---------------------------------------------------------------------------------------------
// Cost builders of the host tracker.
// unaryCostBox — class.cpp:1909-1952: candidate normalised by the box size; inside the prior's area (cv::Point_::inside:
// x <= px < x + w, same for y); val = sqrt(d.d) * (1 / sqrt(2)); stored (1 - val) * score when val <= max_distance.
// ---------------------------------------------------------------------------------------------
void lmo_unary_cost_box(const lm_cand *c, int32_t n, int32_t bb_w, int32_t bb_h, const lm_location_prior *priors,
                        int32_t n_priors, double *out) {
    for (int64_t k = 0; k < (int64_t)n * n_priors; ++k) out[k] = 0.0;  // MyMat(n, m) zero-fills (MyMat.cpp:50-63)
    const double norm_fact = 1 / std::sqrt(2.0);
    for (int i = 0; i < n; ++i) {
        const double cx = (double)c[i].x / (double)bb_w, cy = (double)c[i].y / (double)bb_h;
        for (int j = 0; j < n_priors; ++j) {
            const lm_location_prior &P = priors[j];
            if (P.area_x <= cx && cx < P.area_x + P.area_w && P.area_y <= cy && cy < P.area_y + P.area_h) {
                const double dx = cx - P.pos_x, dy = cy - P.pos_y;
                const double val = std::sqrt(dx * dx + dy * dy) * norm_fact;
                if (val <= P.max_distance) out[(int64_t)j * n + i] = (1 - val) * c[i].s;  // column-major put (MyMat.cpp:65-69)
            }
        }
    }
}

// pairwisePotential — class.cpp:1954-2070, literally on a dense D, then MATSPARSE(const MyMat*) — MyMat.cpp:141-178
// (column by column, rows ascending, entries equal to 0 dropped).  Quirk kept: the "ONG -> X(i+1)" entries are written
// inside the i == 0 iteration of the loop over frame i's candidates (2012-2024), so they are missing when Ci is empty.
int lmo_pairwise_potential(const lm_cand *ci, int32_t ni, const lm_cand *cip1, int32_t nip1, const lm_pairwise_params *p,
                           int32_t *jc, int32_t *ir, double *pr, int64_t cap, int32_t dims[3]) {
    const int nong = p->ong_w * p->ong_h;
    const int nrows = nip1 + nong, ncols = ni + nong;
    std::vector<double> D((size_t)nrows * ncols, 0.0);
    auto put = [&](int r, int c, double v) { D[(size_t)c * nrows + r] = v; };
    auto clampi = [](int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); };  // matchToRange (class.hpp:364-374)
    const double occ = p->occluded_cost * p->alpha_vel;
    const int xa = p->ong_w - 1, ya = p->ong_h - 1;
    for (int i = 0; i < ni; ++i) {
        const int32_t xc = (int32_t)std::round((p->grid_x - (double)ci[i].x) / p->grid_spacing);
        const int32_t yc = (int32_t)std::round((p->grid_y - (double)ci[i].y) / p->grid_spacing);
        put(nip1 + (clampi(yc, 0, ya) * p->ong_w + clampi(xc, 0, xa)), i, occ);
        for (int j = 0; j < nip1; ++j) {
            if (i == 0) {
                const int32_t x2 = (int32_t)std::round((p->grid_x - (double)cip1[j].x) / p->grid_spacing);
                const int32_t y2 = (int32_t)std::round((p->grid_y - (double)cip1[j].y) / p->grid_spacing);
                put(j, ni + (clampi(y2, 0, ya) * p->ong_w + clampi(x2, 0, xa)), occ);
            }
            const double dx = (double)cip1[j].x - (double)ci[i].x, dy = (double)cip1[j].y - (double)ci[i].y;
            const double dist = std::sqrt(dx * dx + dy * dy);
            if (dist < p->max_displacement) {
                double inv = 1 - (dist / p->max_displacement);
                inv = inv * p->alpha_vel;
                put(j, i, inv);
            }
        }
    }
    for (int g = 0; g < nong; ++g) put(nip1 + g, ni + g, occ);
    int64_t nz = 0;
    jc[0] = 0;
    for (int c = 0; c < ncols; ++c) {
        for (int r = 0; r < nrows; ++r) {
            const double v = D[(size_t)c * nrows + r];
            if (v != 0) {
                if (nz < cap) {
                    ir[nz] = r;
                    pr[nz] = v;
                }
                ++nz;
            }
        }
        jc[c + 1] = (int32_t)nz;
    }
    dims[0] = nrows;
    dims[1] = ncols;
    dims[2] = (int32_t)nz;
    return nz > cap ? 1 : 0;
}

int lmo_match_views(const lm_cand *cb, int32_t nb, const lm_cand *cs, int32_t ns, int32_t vel_check,
                    int32_t tw_b, int32_t th_b, int32_t tw_s, int32_t th_s, double T, const uint8_t *I,
                    const uint8_t *Iprev, int32_t n_rows, int32_t n_cols, int32_t x0, int32_t y0b, int32_t y0s,
                    int32_t *match_n, int32_t *match_y, double *match_s, int32_t match_cap, int32_t *n_match) {
    return match_views(cb, nb, cs, ns, vel_check, tw_b, th_b, tw_s, th_s, T, I, Iprev, n_rows, n_cols, x0, y0b,
                       y0s, match_n, match_y, match_s, match_cap, n_match);
}

int lmo_detect(const lm_config *cfg, const lm_template t[2][3], const uint8_t *bkg, const int32_t *calib,
               const uint8_t *frames, const uint8_t *prev_frame, int64_t n, int64_t first_frame_index,
               const uint32_t *bb_x, const uint32_t *bb_y_side, const uint32_t *bb_y_bottom, lm_results *out,
               int n_threads, double *stage_seconds) {
    if (int e = validate(cfg, t)) return e;
    if (!bkg || !calib || !frames || !out || n < 0) return LM_ERR_INVALID;
    if (first_frame_index > 0 && !prev_frame && n > 0) return LM_ERR_INVALID;
    const Geom g = make_geom(*cfg, t);
    for (int64_t f = 0; f < n; ++f)
        if (!roi_ok(*cfg, g, bb_x[f], bb_y_side[f], bb_y_bottom[f])) return LM_ERR_ROI;
    const int64_t fsz = (int64_t)cfg->vid_rows * cfg->vid_cols;
    n_threads = std::max(1, std::min<int>(n_threads, (int)std::max<int64_t>(n, 1)));
    std::vector<std::vector<double>> st(n_threads, std::vector<double>(6, 0.0));
    std::atomic<int64_t> next(0);
    auto worker = [&](int tid) {
        Scratch S;
        for (;;) {
            int64_t f = next.fetch_add(1);
            if (f >= n) break;
            const uint8_t *fr = frames + f * fsz;
            const uint8_t *pv = f > 0 ? frames + (f - 1) * fsz : prev_frame;
            bool vel = (first_frame_index + f) > 0;
            detect_frame(*cfg, t, g, bkg, calib, fr, pv, vel, bb_x[f], bb_y_side[f], bb_y_bottom[f], f, out, S,
                         st[tid].data());
        }
    };
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i);
        for (auto &x : th) x.join();
    }
    if (stage_seconds)
        for (int k = 0; k < 6; ++k) {
            stage_seconds[k] = 0;
            for (int i = 0; i < n_threads; ++i) stage_seconds[k] += st[i][k];
        }
    int rc = LM_OK;
    for (int64_t f = 0; f < n; ++f)
        if (out->flags[f]) rc = LM_ERR_OVERFLOW;
    return rc;
}

}  // extern "C"
