// LocoMouse_class.hpp — host-side C++ mirror of the reference's tracking-problem classes for the
// per-frame detection path (LocoMouse_Core/LocoMouse_class.hpp:170-350, LocoMouse_TM.hpp:45-47,
// LocoMouse_TM_DE.hpp:39-43).  Same class names, same public method names, same call sequence as
// the reference's main.cpp:43-91, same result members (CANDIDATES_*, CANDIDATES_MATCHED_VIEWS_*,
// TRACKS_TAIL, BB_*), same error convention (std::invalid_argument / std::runtime_error).
//
// What differs is WHERE the work happens: the per-frame methods do no pixel work.  The first
// readFrame() of a chunk sends the chunk's raw frames through lm_detect_batch (include/locomouse_b200.h)
// and the per-frame methods then move that frame's records from the batched result buffers into the
// reference's per-frame vectors, so a driver written against the reference's interface runs
// unmodified.  There is no CPU implementation behind these methods.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/locomouse_b200.h"
#include "Candidates.hpp"
#include "LocoMouse_ParseInputs.hpp"
#include "MyMat.hpp"

// LocoMouse_Parameters (class.hpp:48-100): the subset that reaches the detection path, plus the
// inputs that replace the out-of-scope pass 1: per-frame box positions from a file.
// LocoMouse_LocationPrior (class.hpp:33-45, ctor class.cpp:3196-3202)
class LocoMouse_LocationPrior {
    double X = 0, Y = 0, MAX_DISTANCE = 1, MIN_X = 0, MAX_X = 1, MIN_Y = 0, MAX_Y = 1;

public:
    LocoMouse_LocationPrior() = default;
    LocoMouse_LocationPrior(double x, double y, double md, double minx, double maxx, double miny, double maxy);
    double max_distance() const { return MAX_DISTANCE; }
    lm_location_prior to_c() const { return lm_location_prior{X, Y, MAX_DISTANCE, MIN_X, MIN_Y, MAX_X - MIN_X, MAX_Y - MIN_Y}; }
};

// Page-locked host array (lm_host_alloc): frames upload at the full PCIe rate and result arrays receive their device ->
// host copies directly.  Falls back to nothing: lm_host_alloc failing is an error (std::runtime_error).
template <typename T>
class PinnedArray {
    T *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;  // shrinking keeps the allocation (page-locking memory is slow)

public:
    PinnedArray() = default;
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    ~PinnedArray() { lm_host_free(p_); }
    void resize(size_t n) {
        if (n <= cap_) {
            n_ = n;
            return;
        }
        lm_host_free(p_);
        p_ = nullptr;
        n_ = cap_ = 0;
        void *q = nullptr;
        if (lm_host_alloc(&q, n * sizeof(T)) != LM_OK) throw std::runtime_error("lm_host_alloc failed (page-locked host memory)");
        p_ = static_cast<T *>(q);
        n_ = cap_ = n;
    }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
};

class LocoMouse_Parameters {
public:
    // host-tracker cost builders (class.hpp:59-68, 88; class.cpp:132-145): used by computeUnaryCostsBottom / computePairwiseCostsBottom
    int max_displacement_bottom = 15;
    int occlusion_grid_spacing_pixels_bottom = 20;
    double occlusion_grid_max_width = 0.75;
    double alpha_vel_bottom = 1E-1;
    double pairwise_occluded_cost = 1E-2;
    // side-view tracker (class.hpp:60-61, 67): bestSideViewMatch / pairwisePotential_SideView
    int max_displacement_side = 15;
    int occlusion_grid_spacing_pixels_side = 20;
    double alpha_vel_side = 100;
    int tracker_threads = 0;          // host threads for the independent match2nd problems (0: one per hardware thread)
    std::vector<LocoMouse_LocationPrior> PRIOR_PAW, PRIOR_SNOUT;  // config key location_prior (5 x 7); empty: cost builders are skipped
    int conn_comp_connectivity = 8;
    int median_filter_size = 11;      // class.hpp:54-55: pass 1 of the base class
    int min_pixel_visible = 1;
    int pass1_integer_sums = 0;       // 0: firstLastOverT reads the CV_32S sums as floats, as the reference does; 1: as integers
    double side_bottom_min_overlap = 0.7;
    double tail_sub_bounding_box = 0.6;
    int use_provided_bb = 0;
    cv::Rect BB_USER_SIDE, BB_USER_BOTTOM;
    static constexpr unsigned int N_paws = 4, N_snout = 1, N_tail_points = 15;
    // TM / TM_DE (LocoMouse_TM.cpp:57-112)
    int bb_width = 400, bb_height_side = 150;
    // LocoMouse_TM_Parameters (LocoMouse_TM.cpp:44-113); -1: key absent from the configuration file
    int bw_threshold_bottom = -1, bw_threshold_side = -1, min_pixel_count = -1;
    int zero_col_pre = -1, zero_col_post = -1, zero_row_pre = -1, zero_row_post = -1;
    std::string disk_filter_file;   // default: diskfilter.yml beside the executable (LocoMouse_TM.cpp:6)
    int moving_average_window = 5;
    // B200 path
    std::string bounding_box_file;  // pass-1 output (BB_X_POS, BB_Y_SIDE_POS, BB_Y_BOTTOM_POS), see lm_files.hpp
    int device = 0;
    int batch_frames = 4096;        // frames per lm_detect_batch call
    int fma_mode = 1;
    int cand_cap = 64, det_cap = 8192, match_cap = 256;

    LocoMouse_Parameters() = default;
    explicit LocoMouse_Parameters(const std::string &config_file_name);  // throws std::invalid_argument
};

// LocoMouse_Feature / LocoMouse_Model (class.hpp:110-167): templates, biases, sizes, match boxes.
class LocoMouse_Feature {
    std::vector<float> W_B, W_S;
    cv::Size SIZE_B, SIZE_S;
    double RHO_B = 0, RHO_S = 0;
    cv::Rect MATCH_BOX_B, MATCH_BOX_S;

public:
    LocoMouse_Feature() = default;
    LocoMouse_Feature(std::vector<float> w_b, cv::Size size_b, double rho_b, std::vector<float> w_s, cv::Size size_s,
                      double rho_s);
    const std::vector<float> &w_b() const { return W_B; }
    const std::vector<float> &w_s() const { return W_S; }
    double rho_b() const { return RHO_B; }
    double rho_s() const { return RHO_S; }
    cv::Size size_bottom() const { return SIZE_B; }
    cv::Size size_side() const { return SIZE_S; }
    cv::Rect match_box_bottom() const { return MATCH_BOX_B; }  // class.cpp:2954-2969
    cv::Rect match_box_side() const { return MATCH_BOX_S; }
};

class LocoMouse_Model {
public:
    LocoMouse_Feature paw, snout, tail;
    LocoMouse_Model() = default;
    explicit LocoMouse_Model(const std::string &model_file_name);  // throws std::runtime_error
};

namespace cvyaml {
class Writer;
}
using cvyaml_writer = cvyaml::Writer;

class LocoMouse {
protected:
    LocoMouse_Parameters LM_PARAMS;
    std::string LM_CALL, CONFIG_FILE, VIDEO_FILE, BKG_FILE, MODEL_FILE, CALIBRATION_FILE, FLIP_CHAR, OUTPUT_PATH;
    std::string output_file;

    PinnedArray<uint8_t> VIDEO;       // raw 8-bit frames (channel 0), N_FRAMES x vid_rows x vid_cols, page-locked
    int VID_ROWS = 0, VID_COLS = 0;
    std::vector<uint8_t> BKG;
    std::vector<int32_t> CALIBRATION;  // ind_warp_mapping, N_ROWS x N_COLS
    bool IMAGE_FLIP = false;

    cv::Rect BB_SIDE_VIEW, BB_BOTTOM_VIEW;
    cv::Rect BB_BOTTOM_MOUSE, BB_SIDE_MOUSE;
    std::vector<unsigned int> BB_X_POS, BB_Y_SIDE_POS, BB_Y_BOTTOM_POS;

    unsigned int N_FRAMES = 0, N_ROWS = 0, N_COLS = 0;
    int CURRENT_FRAME = -1, METHOD = 0;

    LocoMouse_Model M;

    std::vector<std::vector<Candidate>> CANDIDATES_BOTTOM_PAW, CANDIDATES_BOTTOM_SNOUT;
    std::vector<std::vector<Candidate>> CANDIDATES_SIDE_PAW, CANDIDATES_SIDE_SNOUT;
    std::vector<std::vector<P22D>> CANDIDATES_MATCHED_VIEWS_PAW, CANDIDATES_MATCHED_VIEWS_SNOUT;
    std::vector<std::vector<int32_t>> TRACKS_TAIL;  // per frame 3 x N_tail_points (x, y, z), -1 = missing
    std::vector<MyMat> UNARY_BOTTOM_PAW, UNARY_BOTTOM_SNOUT;            // class.hpp: same names
    std::vector<MATSPARSE> PAIRWISE_BOTTOM_PAW, PAIRWISE_BOTTOM_SNOUT;
    std::string costs_file, tracks_file;
    // occlusion grids (class.hpp:238-246; filled by initializeFeatureLoop, class.cpp:726-759) and the tracker's results
    std::vector<cv::Point_<double>> ONG;
    cv::Size ONG_size;
    cv::Point_<double> ONG_BR_corner;
    std::vector<unsigned int> ONG_SIDE;
    unsigned int ONG_SIDE_LOWEST_POINT = 0;
    cv::Mat TRACK_INDEX_PAW_BOTTOM, TRACK_INDEX_SNOUT_BOTTOM, TRACK_INDEX_PAW_SIDE, TRACK_INDEX_SNOUT_SIDE;  // points x frames labels
    std::vector<cv::Mat> EXPORTED;    // the matrices exportResults wrote, in file order (tests)

    // ---- device side ------------------------------------------------------------------------------
    lm_ctx *CTX = nullptr;
    struct Batch;                      // result buffers of the chunk that contains CURRENT_FRAME
    std::unique_ptr<Batch> BATCH, SPARE;   // SPARE: the chunk before BATCH, whose page-locked arrays the next chunk reuses
    bool LOOP_READY = false;

    void initializePaths(const LocoMouse_ParseInputs &INPUT);
    void loadVideo();
    void loadBackground();
    void loadCalibration();
    void loadFlip();
    void validateImageVideoSize();
    void check(int rc) const;          // lm_status -> exception
    void runChunk(unsigned int first_frame);
    void configureDevice();            // lm_configure + background + calibration for the current box sizes
    const Batch &batchFor(int frame) const;
    virtual bool usesImadjust() const { return false; }  // LocoMouse_TM::readFrame applies imadjust(0, 0.6)
    // class.cpp:2221-2346, 2385-2482
    cv::Mat bestSideViewMatch(const cv::Mat &T, const std::vector<std::vector<P22D>> &candidates_bottom_side_matched,
                              const std::vector<unsigned int> &ONG_side, unsigned int lowest_point, unsigned int N_features);
    void exportPointTracks(cvyaml_writer &out, const cv::Mat &T_bottom, const cv::Mat &T_side,
                           const std::vector<std::vector<P22D>> &candidates_bottom_side_matched, const std::string &feature_name, unsigned int N_features);
    void exportTracks();
    void exportLineTracks(cvyaml_writer &out, const std::vector<std::vector<int32_t>> &Tracks, const std::string &track_name, int N_line_points);

public:
    explicit LocoMouse(LocoMouse_ParseInputs INPUTS);
    virtual ~LocoMouse();
    LocoMouse(const LocoMouse &) = delete;
    LocoMouse &operator=(const LocoMouse &) = delete;

    virtual void readFrame();
    virtual void getBoundingBox();
    virtual void computeBoundingBox();
    void initializeFeatureLoop();
    void cropBoundingBox();
    void detectTail();
    void detectBottomCandidates();
    void computeUnaryCostsBottom();
    void computePairwiseCostsBottom();
    void detectSideCandidates();
    void matchBottomSideCandidates();
    void storePreviousImage();
    void computeBottomTracks();
    void computeSideTracks();
    void exportResults();

    unsigned int N_frames() const { return N_FRAMES; }

    // read access for the host stages that follow (cost builders, match2nd) and for tests
    const std::vector<std::vector<Candidate>> &candidatesBottomPaw() const { return CANDIDATES_BOTTOM_PAW; }
    const std::vector<std::vector<Candidate>> &candidatesBottomSnout() const { return CANDIDATES_BOTTOM_SNOUT; }
    const std::vector<std::vector<Candidate>> &candidatesSidePaw() const { return CANDIDATES_SIDE_PAW; }
    const std::vector<std::vector<Candidate>> &candidatesSideSnout() const { return CANDIDATES_SIDE_SNOUT; }
    const std::vector<std::vector<P22D>> &candidatesMatchedViewsPaw() const { return CANDIDATES_MATCHED_VIEWS_PAW; }
    const std::vector<std::vector<P22D>> &candidatesMatchedViewsSnout() const { return CANDIDATES_MATCHED_VIEWS_SNOUT; }
    const std::vector<std::vector<int32_t>> &tracksTail() const { return TRACKS_TAIL; }
    const std::vector<MyMat> &unaryBottomPaw() const { return UNARY_BOTTOM_PAW; }
    const std::vector<MyMat> &unaryBottomSnout() const { return UNARY_BOTTOM_SNOUT; }
    const std::vector<MATSPARSE> &pairwiseBottomPaw() const { return PAIRWISE_BOTTOM_PAW; }
    const std::vector<MATSPARSE> &pairwiseBottomSnout() const { return PAIRWISE_BOTTOM_SNOUT; }
    const cv::Mat &trackIndexPawBottom() const { return TRACK_INDEX_PAW_BOTTOM; }
    const cv::Mat &trackIndexSnoutBottom() const { return TRACK_INDEX_SNOUT_BOTTOM; }
    const cv::Mat &trackIndexPawSide() const { return TRACK_INDEX_PAW_SIDE; }
    const cv::Mat &trackIndexSnoutSide() const { return TRACK_INDEX_SNOUT_SIDE; }
    // LocoMouse_class.cpp:2073-2150 (see lm_track::side_view_transitions)
    static MATSPARSE pairwisePotential_SideView(const std::vector<unsigned int> &Zi, const std::vector<unsigned int> &Zip1, double grid_mapping,
                                                double grid_spacing, const std::vector<unsigned int> &ONGi, unsigned int Nong,
                                                double max_displacement_bottom, double alpha_vel_bottom, double pairwise_occluded_cost);
};

// LocoMouse_TM (LocoMouse_TM.hpp:45-47): imadjust in readFrame; box = bb_width x bb_height_side (side,
// anchored at row 164) and bb_width x bottom-view height (LocoMouse_TM.cpp:142-155).
class LocoMouse_TM : public LocoMouse {
protected:
    bool usesImadjust() const override { return true; }

public:
    explicit LocoMouse_TM(LocoMouse_ParseInputs INPUTS);
    void computeBoundingBox() override;

protected:
    std::vector<float> DISK_FILTER;   // diskfilter.yml "H" (LocoMouse_TM.cpp:4-14), row-major, DISK_SIZE x DISK_SIZE
    int DISK_SIZE = 0;
    std::string REF_PATH;
};

// LocoMouse_TM_DE (LocoMouse_TM_DE.hpp:39-43): same readFrame; 400-wide boxes over the full view heights
// (LocoMouse_TM_DE.cpp:38-51).
class LocoMouse_TM_DE : public LocoMouse_TM {
public:
    explicit LocoMouse_TM_DE(LocoMouse_ParseInputs INPUTS);
    void computeBoundingBox() override;
};

// LocoMouse_Methods.cpp:3-26
std::unique_ptr<LocoMouse> LocoMouse_Initialize(LocoMouse_ParseInputs INPUT);
