// match2nd.hpp — the host tracker behind the reference's match2nd() / computeCostTrack() interface
// (match2nd/match2nd.h:566-568, match2nd/match2nd.cpp:11-191): multi-target max-product message passing over the sparse
// candidate trellis the cost builders emit (MyMat unary matrices, MATSPARSE transitions), track by track, with exclusion
// messages between tracks.  north_star keeps this sequential stage on the host; this is a fresh implementation with the
// reference's exact arithmetic (every sum in the reference's association order, so labels agree bit for bit -- pinned
// against the reference's own match2nd.cpp / match2nd.h compiled from /root/reference, tests/test_match2nd.py) and
// WITHOUT the reference's process-global state (match2nd.cpp:4-8: locations, occ, occ_score, BAM), so several trackers
// can run at once: see lm_track::Solver and match2nd_concurrent().
#pragma once
#include <vector>

#include "MyMat.hpp"
#include "cv_shim.hpp"

// Label matrix `points x frames` (CV_32SC1): label < unary_costs[f].Nrows() -> that candidate of frame f,
// label >= Nrows() -> occlusion-grid node (label - Nrows()), -1 -> no satisfiable labelling (match2nd.h:341-345).
cv::Mat match2nd(const std::vector<MyMat> &unary_costs, const std::vector<MATSPARSE> &pairwise_costs, int Nong, double occlusion_point_cost,
                 double bam_tie, unsigned int frames, unsigned int points, const int *permutation);

// Sum of the unary and pairwise terms along the four paw tracks of M (match2nd.cpp:162-191).
double computeCostTrack(const cv::Mat &M, const std::vector<MyMat> &unary_costs, const std::vector<MATSPARSE> &pairwise_costs, const int *permutation);

namespace lm_track {

// LocoMouse::pairwisePotential_SideView (LocoMouse_class.cpp:2073-2150): transitions of the side-view tracker between the
// side candidates (image rows Zi, Zip1) of one feature in consecutive frames, with a 1-D occlusion grid of Nong nodes
// below `grid_mapping` spaced `grid_spacing` apart.  (Nip1 + Nong) x (Ni + Nong), entries equal to 0 not stored.
MATSPARSE side_view_transitions(const std::vector<unsigned int> &Zi, const std::vector<unsigned int> &Zip1, double grid_mapping,
                                double grid_spacing, unsigned int Nong, double max_displacement, double alpha_vel, double pairwise_occluded_cost);

struct Job {  // one match2nd() call
    const std::vector<MyMat> *unary = nullptr;
    const std::vector<MATSPARSE> *pairwise = nullptr;
    int Nong = 0;
    double occlusion_point_cost = 0, bam_tie = 0;
    unsigned int frames = 0, points = 0;
    const int *permutation = nullptr;
    cv::Mat result;
};
// Runs the jobs on up to n_threads host threads (0: one per hardware thread); results as if each had been passed to
// match2nd() in turn.  The reference cannot do this: its tracker keeps its state in globals.
void match2nd_concurrent(std::vector<Job> &jobs, unsigned int n_threads = 0);

}  // namespace lm_track
