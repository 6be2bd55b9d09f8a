/*
 * lm_oracle.h — CPU oracle for the LocoMouse per-frame detection path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference algorithm
 * (careylab/LocoMouse_cpp, file:line cited on every function in lm_oracle.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (locomouse_cpp_b200/) never does.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures, and cannot be built
 * here (it needs the OpenCV C++ SDK, which is absent).  The oracle is therefore pinned against
 * the only executable piece of the reference stack available in this image — OpenCV 4.13 through
 * python `cv2` — primitive by primitive (tests/test_oracle_vs_cv2.py): filter2D, normalize,
 * subtract, threshold, LUT, flip, connectedComponentsWithStats, moments; plus hand-derived
 * known-answer tests for the reference's own loops (nmsMax, peakClustering, matchViews).
 * nmsMax, peakClustering, vecmovingaverage, firstLastOverT, LocoMouse::imadjust and the Candidate / P22D
 * classes use OpenCV for value types only (imadjust additionally cv::LUT), so the
 * reference's OWN source lines for them are compiled from /root/reference against a value-type shim
 * (oracle/ref_shim, `make -C oracle ref` -> oracle/_ref/libref_nms.so) and the oracle is checked
 * against that code bit for bit (tests/test_oracle_vs_reference.py, golden vectors in
 * tests/golden/reference_nms.npz).  The remaining loops (matchViews, tail segmentation, readFrame)
 * call OpenCV algorithms and have no independent executable check: "parity unpinned" in the sense of
 * SURVEY.md §8c for those, pinned primitive by primitive against cv2.
 */
#ifndef LM_ORACLE_H
#define LM_ORACLE_H

#include "../include/locomouse_b200.h" /* shared POD types only: lm_config, lm_template, lm_results */

#ifdef __cplusplus
extern "C" {
#endif

/* Full path for n consecutive frames (same argument meaning as lm_detect_batch, host memory only).
 * n_threads >= 1 farms frames over std::thread (the reference itself is single threaded).
 * stage_seconds[6] (optional) accumulates CPU seconds: preprocess, correlation, tail, nms bottom+side,
 * pairing, total — summed over threads. */
int lmo_detect(const lm_config *cfg, const lm_template t[2][3], const uint8_t *bkg,
               const int32_t *calib, const uint8_t *frames, const uint8_t *prev_frame, int64_t n,
               int64_t first_frame_index, const uint32_t *bb_x, const uint32_t *bb_y_side,
               const uint32_t *bb_y_bottom, lm_results *out, int n_threads, double *stage_seconds);

/* geometry of initializeFeatureLoop (class.cpp:655-721): same layout as lm_get_geometry */
int lmo_geometry(const lm_config *cfg, const lm_template t[2][3], int32_t pads[8], int32_t canvas[4]);
/* 0 if the frame's boxes are valid ROIs of the padded canvas, else -3 (LM_ERR_ROI) */
int lmo_check_roi(const lm_config *cfg, const lm_template t[2][3], uint32_t bb_x, uint32_t bb_y_side,
                  uint32_t bb_y_bottom);

/* ---- stage-level entry points, for pinning against cv2 and for known-answer tests ---------- */
/* readFrame (+imadjust if cfg->imadjust): raw frame -> calibrated image I [n_rows][n_cols];
 * minmax[2] (optional) returns the min/max of sat(F-BKG). */
int lmo_preprocess(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frame,
                   uint8_t *I, int32_t *minmax);
void lmo_imadjust_lut(double low_in, double high_in, double low_out, double high_out, uint8_t lut[256]);
/* filter2D(ROI of zero-extended image, CV_32F, K, anchor(-1,-1), delta=-rho, BORDER_CONSTANT):
 * scores for the w x h window whose top-left is (x0,y0) in image coordinates. */
void lmo_correlate(const uint8_t *I, int32_t n_rows, int32_t n_cols, const lm_template *t, int32_t x0,
                   int32_t y0, int32_t w, int32_t h, int32_t fma_mode, float *scores);
/* nmsMax / peakClustering on a float score map; returns candidate count (uncapped), writes <= cap */
int lmo_nms_max(const float *scores, int32_t rows, int32_t cols, int32_t box_w, int32_t box_h,
                lm_cand *out, int32_t cap);
int lmo_peak_clustering(const float *scores, int32_t rows, int32_t cols, int32_t box_w, int32_t box_h,
                        lm_cand *out, int32_t cap);
/* selectLargestRegion: binary u8 (nonzero = foreground) -> u8 0/255 */
void lmo_largest_region(const uint8_t *bin, int32_t rows, int32_t cols, int32_t conn, uint8_t *out);
/* detectLineCandidates after binarisation: tail tracks + masks from the two >0 maps */
void lmo_tail_from_binary(const uint8_t *bin_bottom, int32_t rows_b, const uint8_t *bin_side,
                          int32_t rows_s, int32_t cols, int32_t conn, int32_t n_points, int32_t *tracks,
                          uint8_t *tail_mask);
/* matchingWithVelocityConstraint on given candidate lists. I / Iprev are calibrated images
 * (pre-processed, [n_rows][n_cols]); (x0,y0b,y0s) the unpadded crop origins in image coordinates.
 * Returns 0; fills match_n[nb], match_y/match_s (concatenated), *n_match total. */
int lmo_match_views(const lm_cand *cb, int32_t nb, const lm_cand *cs, int32_t ns, int32_t vel_check,
                    int32_t tw_b, int32_t th_b, int32_t tw_s, int32_t th_s, double T, const uint8_t *I,
                    const uint8_t *Iprev, int32_t n_rows, int32_t n_cols, int32_t x0, int32_t y0b,
                    int32_t y0s, int32_t *match_n, int32_t *match_y, double *match_s, int32_t match_cap,
                    int32_t *n_match);

/* ---- pass 1, LocoMouse_TM_DE -------------------------------------------------------- */
/* computeMouseBox_DE per frame (LocoMouse_TM_DE.cpp:56-113 after the base readFrame): bb_x[n], lims[n][2] (optional) */
int lmo_bounding_box_tm_de(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames,
                           int64_t n, const lm_bb_de_params *p, double *bb_x, int32_t *lims);
/* imadjust_default's mapping for a given 256-bin histogram (LocoMouse_class.cpp:3244-3311): lut[256], {imin, imax} */
void lmo_imadjust_default_lut(const uint32_t *hist, uint8_t *lut, int32_t *imin_imax);
/* firstLastOverT<int> (LocoMouse_class.hpp:411-442) on float column sums */
void lmo_first_last_over_t(const float *values, uint32_t L, int32_t th, int32_t *first_last);
/* ---- pass 1, base class: computeMouseBox (LocoMouse_class.cpp:921-997) after the base readFrame ---- */
int lmo_bounding_box_base(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames, int64_t n,
                          const lm_bb_base_params *p, double *box, int32_t *lims);
/* the same from an already pre-processed image I [n_rows][n_cols] (what the reference-compiled checker is fed) */
int lmo_mouse_box_base(const uint8_t *I, int32_t n_rows, int32_t n_cols, int32_t conn, const lm_bb_base_params *p, double *box, int32_t *lims);
/* computeMouseBoxSize + medianvec + stdvec (LocoMouse_class.cpp:1481-1556); sorts the inputs like the reference */
/* LocoMouse_TM pass 1 (LocoMouse_TM.cpp:115-269); lmo_mouse_box_tm works on one calibrated image and can hand back every
 * intermediate image (any of the five output pointers may be NULL) */
void lmo_filter2d_u8(const uint8_t *src, int32_t rows, int32_t cols, const float *kernel, int32_t k, uint8_t *dst);
int lmo_mouse_box_tm(const uint8_t *I, int32_t n_rows, int32_t n_cols, int32_t conn, const lm_bb_tm_params *p, double *bb_x, int32_t *lims,
                     uint8_t *adjusted, uint8_t *binary, uint8_t *opened, uint8_t *filtered, int32_t *row_sums);
int lmo_bounding_box_tm(const lm_config *cfg, const uint8_t *bkg, const int32_t *calib, const uint8_t *frames, int64_t n,
                        const lm_bb_tm_params *p, double *bb_x, int32_t *lims);
void lmo_mouse_box_size(double *w, double *hb, double *hs, int64_t n, int32_t size[3]);
/* ---- cost builders ----------------------------------------------------------------- */
/* unaryCostBox (LocoMouse_class.cpp:1909-1952): out = MyMat(n x n_priors), column-major */
void lmo_unary_cost_box(const lm_cand *c, int32_t n, int32_t bb_w, int32_t bb_h, const lm_location_prior *priors,
                        int32_t n_priors, double *out);
/* pairwisePotential (1954-2070) followed by MATSPARSE(const MyMat*) (MyMat.cpp:141-178).  jc has ni + Nong + 1 entries,
 * ir / pr up to cap; dims = {n_rows, n_cols, nnz}.  Returns 1 when cap is too small (nnz still reported). */
int lmo_pairwise_potential(const lm_cand *ci, int32_t ni, const lm_cand *cip1, int32_t nip1, const lm_pairwise_params *p,
                           int32_t *jc, int32_t *ir, double *pr, int64_t cap, int32_t dims[3]);
/* branch-coverage counters of the pairing stage since the last reset (see g_cov in lm_oracle.cpp) */
void lmo_coverage(int64_t out[8], int reset);
/* vecmovingaverage (LocoMouse_class.cpp:1559-1608) */
void lmo_vecmovingaverage(const double *v, int64_t n, int32_t window, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
