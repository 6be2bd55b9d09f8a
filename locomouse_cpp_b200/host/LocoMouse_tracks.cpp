// LocoMouse_tracks.cpp — the stages of the reference's driver that follow the per-frame loop (main.cpp:85-91):
// computeBottomTracks, computeSideTracks / bestSideViewMatch / pairwisePotential_SideView (LocoMouse_class.cpp:2073-2346) and
// exportResults' track writers exportPointTracks / exportLineTracks (2385-2482) into output_<stem>.yml.
// north_star keeps the sequential track assignment on the host; what differs from the reference is only that the
// tracker (match2nd.cpp here) is re-entrant, so the independent problems -- the paw permutations and the snout; the
// side-view tracks of every feature -- are solved on several host threads at once.
#include <iostream>
#include <stdexcept>

#include "LocoMouse_class.hpp"
#include "cv_yaml.hpp"
#include "match2nd.hpp"

namespace {
// LocoMouse_Parameters::PAW_PERMUTATIONS is a 4 x 24 matrix filled row by row with the 24 permutations, and
// computeBottomTracks reads ROW i_perm for i_perm < N_paws (class.hpp:91-92, class.cpp:2170-2171): the four orders it
// really tries are the first four entries of rows 0..3 of that layout.
const int PAW_ORDERS[4][4] = {{3, 2, 1, 0}, {2, 3, 1, 0}, {1, 2, 3, 0}, {0, 2, 1, 3}};
}  // namespace

MATSPARSE LocoMouse::pairwisePotential_SideView(const std::vector<unsigned int> &Zi, const std::vector<unsigned int> &Zip1, double grid_mapping,
                                                double grid_spacing, const std::vector<unsigned int> &ONGi, unsigned int Nong,
                                                double max_displacement_bottom, double alpha_vel_bottom, double pairwise_occluded_cost) {
    (void)ONGi;  // the reference passes the grid but only uses its size
    return lm_track::side_view_transitions(Zi, Zip1, grid_mapping, grid_spacing, Nong, max_displacement_bottom, alpha_vel_bottom, pairwise_occluded_cost);
}

void LocoMouse::computeBottomTracks() {
    if (LM_PARAMS.PRIOR_PAW.empty() || LM_PARAMS.PRIOR_SNOUT.empty()) {
        // The reference refuses to start without a location_prior (class.cpp:132-145); this mirror also serves callers
        // that only want the per-frame detections (candidate files), so it says so and leaves the tracks empty.
        std::cout << "location_prior is not set: the tracker's cost matrices were not built, tracks are not computed." << std::endl;
        return;
    }
    if (UNARY_BOTTOM_PAW.size() != N_FRAMES || UNARY_BOTTOM_SNOUT.size() != N_FRAMES || PAIRWISE_BOTTOM_PAW.size() + 1 != N_FRAMES)
        throw std::runtime_error(
            "computeBottomTracks(): the unary / pairwise costs of every frame are needed (set location_prior in the configuration and run "
            "computeUnaryCostsBottom / computePairwiseCostsBottom for every frame).");
    const int nong = (int)ONG.size();
    std::vector<lm_track::Job> jobs(LM_PARAMS.N_paws + 1);
    for (unsigned int i_perm = 0; i_perm < LM_PARAMS.N_paws; ++i_perm) {
        lm_track::Job &j = jobs[i_perm];
        j.unary = &UNARY_BOTTOM_PAW;
        j.pairwise = &PAIRWISE_BOTTOM_PAW;
        j.Nong = nong;
        j.frames = N_FRAMES;
        j.points = LM_PARAMS.N_paws;
        j.permutation = PAW_ORDERS[i_perm];
    }
    const int snout_order = 0;
    {
        lm_track::Job &j = jobs[LM_PARAMS.N_paws];
        j.unary = &UNARY_BOTTOM_SNOUT;
        j.pairwise = &PAIRWISE_BOTTOM_SNOUT;
        j.Nong = nong;
        j.frames = N_FRAMES;
        j.points = LM_PARAMS.N_snout;
        j.permutation = &snout_order;
    }
    lm_track::match2nd_concurrent(jobs, (unsigned int)std::max(0, LM_PARAMS.tracker_threads));
    // the best-scoring order wins, the first one on ties (class.cpp:2165-2182)
    double current_cost = -1;
    unsigned int current_perm = 0;
    for (unsigned int i_perm = 0; i_perm < LM_PARAMS.N_paws; ++i_perm) {
        const double c_i = computeCostTrack(jobs[i_perm].result, UNARY_BOTTOM_PAW, PAIRWISE_BOTTOM_PAW, PAW_ORDERS[i_perm]);
        if (c_i > current_cost) {
            current_perm = i_perm;
            current_cost = c_i;
        }
    }
    const cv::Mat &best = jobs[current_perm].result;
    TRACK_INDEX_PAW_BOTTOM = cv::Mat::zeros(best.rows, best.cols, CV_32SC1);
    for (int r = 0; r < 4 && r < best.rows; ++r)  // undo the order: track r of the solver is paw PAW_ORDERS[.][r]
        std::copy(best.ptr<int>(r), best.ptr<int>(r) + best.cols, TRACK_INDEX_PAW_BOTTOM.ptr<int>(PAW_ORDERS[current_perm][r]));
    TRACK_INDEX_SNOUT_BOTTOM = jobs[LM_PARAMS.N_paws].result;
}

void LocoMouse::computeSideTracks() {
    if (LM_PARAMS.PRIOR_PAW.empty() || LM_PARAMS.PRIOR_SNOUT.empty()) return;  // see computeBottomTracks
    if (TRACK_INDEX_PAW_BOTTOM.empty() || TRACK_INDEX_SNOUT_BOTTOM.empty()) throw std::runtime_error("computeSideTracks(): computeBottomTracks() must run first.");
    TRACK_INDEX_PAW_SIDE = bestSideViewMatch(TRACK_INDEX_PAW_BOTTOM, CANDIDATES_MATCHED_VIEWS_PAW, ONG_SIDE, ONG_SIDE_LOWEST_POINT, LM_PARAMS.N_paws);
    TRACK_INDEX_SNOUT_SIDE = bestSideViewMatch(TRACK_INDEX_SNOUT_BOTTOM, CANDIDATES_MATCHED_VIEWS_SNOUT, ONG_SIDE, ONG_SIDE_LOWEST_POINT, LM_PARAMS.N_snout);
}

// class.cpp:2221-2346: per feature, the side candidates of the bottom candidate the bottom track chose in every frame form
// a one-point tracking problem along the image rows (with a 1-D occlusion grid).
cv::Mat LocoMouse::bestSideViewMatch(const cv::Mat &T, const std::vector<std::vector<P22D>> &matched, const std::vector<unsigned int> &ONG_side,
                                     unsigned int lowest_point, unsigned int N_features) {
    const unsigned int nong_side = (unsigned int)ONG_side.size();
    cv::Mat T_side = cv::Mat::zeros((int)N_features, (int)N_FRAMES, CV_32SC1);
    std::vector<std::vector<MyMat>> unary(N_features);
    std::vector<std::vector<MATSPARSE>> pairwise(N_features);
    std::vector<lm_track::Job> jobs(N_features);
    const int order = 0;
    for (unsigned int k = 0; k < N_features; ++k) {
        const int *p_T = T.ptr<int>((int)k);
        std::vector<unsigned int> Z_prev;
        unary[k].reserve(N_FRAMES);
        pairwise[k].reserve(N_FRAMES > 0 ? N_FRAMES - 1 : 0);
        for (unsigned int f = 0; f < N_FRAMES; ++f) {
            std::vector<unsigned int> Z;
            MyMat frame_potentials(0, 1);
            if (p_T[f] >= 0 && (size_t)p_T[f] < matched[f].size()) {  // the reference compares as unsigned: -1 is "no solution"
                const P22D &c = matched[f][(size_t)p_T[f]];
                const int n = c.number_of_candidates();
                frame_potentials = MyMat((unsigned int)n, 1);
                for (int i = 0; i < n; ++i) {
                    frame_potentials.put((unsigned int)i, 0, c.score_side((uint)i));
                    Z.push_back((unsigned int)c.y_side_coord((uint)i));
                }
            }
            unary[k].push_back(std::move(frame_potentials));
            if (f > 0)
                pairwise[k].push_back(pairwisePotential_SideView(Z_prev, Z, (double)lowest_point, (double)LM_PARAMS.occlusion_grid_spacing_pixels_side,
                                                                 ONG_side, nong_side, LM_PARAMS.max_displacement_side, LM_PARAMS.alpha_vel_side,
                                                                 LM_PARAMS.pairwise_occluded_cost));
            Z_prev = Z;
        }
        lm_track::Job &j = jobs[k];
        j.unary = &unary[k];
        j.pairwise = &pairwise[k];
        j.Nong = (int)nong_side;
        j.frames = N_FRAMES;
        j.points = 1;
        j.permutation = &order;
    }
    lm_track::match2nd_concurrent(jobs, (unsigned int)std::max(0, LM_PARAMS.tracker_threads));
    for (unsigned int k = 0; k < N_features; ++k)
        if (!jobs[k].result.empty()) std::copy(jobs[k].result.ptr<int>(0), jobs[k].result.ptr<int>(0) + N_FRAMES, T_side.ptr<int>((int)k));
    return T_side;
}

// class.cpp:2385-2452: per feature an N_FRAMES x 3 matrix (x, y in the bottom view, z = row in the side view) in image
// coordinates, -1 where the track is occluded / has no side match.
void LocoMouse::exportPointTracks(cvyaml_writer &out, const cv::Mat &T_bottom, const cv::Mat &T_side, const std::vector<std::vector<P22D>> &matched,
                                  const std::string &feature_name, unsigned int N_features) {
    for (unsigned int k = 0; k < N_features; ++k) {
        const int *p_T = T_bottom.ptr<int>((int)k), *p_T_side = T_side.ptr<int>((int)k);
        cv::Mat M((int)N_FRAMES, 3, CV_32SC1, -1);
        for (unsigned int f = 0; f < N_FRAMES; ++f) {
            if (p_T[f] < 0 || (size_t)p_T[f] >= matched[f].size()) continue;
            const P22D &c = matched[f][(size_t)p_T[f]];
            int *p_M = M.ptr<int>((int)f);
            p_M[0] = (int)(BB_X_POS[f] - (unsigned int)BB_BOTTOM_MOUSE.width + 1u + (unsigned int)c.x_coord());
            p_M[1] = (int)(BB_Y_BOTTOM_POS[f] - (unsigned int)BB_BOTTOM_MOUSE.height + 1u + (unsigned int)c.y_bottom_coord());
            // the reference tests `p_T_side < number_of_candidates()` as signed and would index with -1; a side label of -1
            // (no satisfiable side track) leaves z = -1 here
            if (p_T_side[f] >= 0 && p_T_side[f] < c.number_of_candidates())
                p_M[2] = (int)(BB_Y_SIDE_POS[f] - (unsigned int)BB_SIDE_MOUSE.height + 1u + (unsigned int)c.y_side_coord((uint)p_T_side[f]));
        }
        out.write(feature_name + std::to_string(k), M.rows, M.cols, M.ptr<int>(0));
        EXPORTED.push_back(M);
    }
}

// class.cpp:2454-2482: 3 x (N_line_points * N_FRAMES), rows x / y / z, frame after frame
void LocoMouse::exportLineTracks(cvyaml_writer &out, const std::vector<std::vector<int32_t>> &Tracks, const std::string &track_name, int N_line_points) {
    const int cols = N_line_points * (int)N_FRAMES;
    cv::Mat L(3, cols, CV_32SC1, -1);
    int *tx = L.ptr<int>(0), *ty = L.ptr<int>(1), *tz = L.ptr<int>(2);
    for (unsigned int f = 0; f < N_FRAMES; ++f) {
        const int32_t *t = Tracks[f].data();  // 3 x N_line_points (x, y, z)
        for (int k = 0; k < N_line_points; ++k, ++tx, ++ty, ++tz) {
            if (t[k] >= 0) *tx = (int)(BB_X_POS[f] - (unsigned int)BB_BOTTOM_MOUSE.width + 1u + (unsigned int)t[k]);
            if (t[N_line_points + k] >= 0) *ty = (int)(BB_Y_BOTTOM_POS[f] - (unsigned int)BB_BOTTOM_MOUSE.height + 1u + (unsigned int)t[N_line_points + k]);
            if (t[2 * N_line_points + k] >= 0) *tz = (int)(BB_Y_SIDE_POS[f] - (unsigned int)BB_SIDE_MOUSE.height + 1u + (unsigned int)t[2 * N_line_points + k]);
        }
    }
    out.write(track_name, L.rows, L.cols, L.ptr<int>(0));
    EXPORTED.push_back(L);
}

// exportResults (class.cpp:2348-2383): paw_tracks0..3, snout_tracks0, tracks_tail
void LocoMouse::exportTracks() {
    EXPORTED.clear();
    if (TRACK_INDEX_PAW_SIDE.empty() || TRACK_INDEX_SNOUT_SIDE.empty() || TRACKS_TAIL.size() != N_FRAMES) return;
    cvyaml::Writer out(tracks_file);
    exportPointTracks(out, TRACK_INDEX_PAW_BOTTOM, TRACK_INDEX_PAW_SIDE, CANDIDATES_MATCHED_VIEWS_PAW, "paw_tracks", LM_PARAMS.N_paws);
    exportPointTracks(out, TRACK_INDEX_SNOUT_BOTTOM, TRACK_INDEX_SNOUT_SIDE, CANDIDATES_MATCHED_VIEWS_SNOUT, "snout_tracks", LM_PARAMS.N_snout);
    exportLineTracks(out, TRACKS_TAIL, "tracks_tail", (int)LM_PARAMS.N_tail_points);
    if (!out.good()) throw std::runtime_error("Could not write " + tracks_file);
}
