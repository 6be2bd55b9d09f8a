// match2nd.cpp — see match2nd.hpp.  Reference: match2nd/match2nd.cpp:11-191 (entry points), match2nd/match2nd.h:27-563
// (class point: margins, assignment; class bundle: track order).
//
// Model.  Frame f has n_f candidates followed by `occ` occlusion-grid nodes: L_f = n_f + occ nodes.  Transition f -> f+1 is
// a CSC matrix with one column per node of frame f; stored entry e of column i = an edge (i -> ir[e]) worth pr[e]
// (MATSPARSE, MyMat.cpp:141-178).  A track is a path through the trellis; its score is the sum of node terms
// (unary cost of a candidate / occlusion cost of a grid node, plus a message shared by all tracks) and edge terms.
// Per track, in the caller's order: forward / backward max-marginals per EDGE, the best and second-best marginal per
// frame, and a penalty (second - best - BAM) on the best node of every frame so that the next track prefers other nodes.
// Then, in reverse order, every track takes back its own penalty, is decoded greedily frame by frame along the backward
// marginals, and forbids its candidates (message = -inf) to the tracks decoded after it.
//
// Every floating-point expression below keeps the reference's operand order (e.g. `m += u + p` is m + (u + p)); the
// labels therefore equal the reference's on every input, including the -inf / NaN cases of empty trellises.
#include "match2nd.hpp"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <thread>

namespace {

constexpr double NEG_INF = -std::numeric_limits<double>::infinity();

struct Trellis {
    int F = 0, occ = 0;
    double occ_score = 0, bam = 0;
    std::vector<int> n;                 // candidates per frame ("locations")
    std::vector<const int *> jc, ir;    // per transition: column starts (L_f + 1), edge heads
    std::vector<const double *> pr;     // per transition: edge values
    std::vector<int> nz;                // per transition: edges
    std::vector<std::vector<double>> message;  // [frame][node], shared by all tracks
    int nodes(int f) const { return n[f] + occ; }
};

struct Track {
    const Trellis *T;
    std::vector<const double *> unary;  // [frame]: this track's column of the frame's unary matrix
    std::vector<std::vector<double>> fore, back;  // [transition][edge]
    std::vector<double> best, second;
    std::vector<int> bestloc, label;

    Track(const Trellis *t) : T(t), unary(t->F), fore(t->F - 1), back(t->F - 1), best(t->F), second(t->F), bestloc(t->F, 0), label(t->F, 0) {
        for (int f = 0; f + 1 < t->F; ++f) {
            fore[f].resize((size_t)t->nz[f]);
            back[f].resize((size_t)t->nz[f]);
        }
    }
    // node term (match2nd.h:31-42): candidates carry their unary cost, grid nodes the occlusion cost
    double node(int f, int b) const { return (b < T->n[f] ? unary[f][b] : T->occ_score) + T->message[f][b]; }

    void clear() {  // match2nd.h:170-180
        for (int f = 0; f < T->F; ++f) label[f] = -2;
        for (int f = 0; f + 1 < T->F; ++f) {
            std::fill(fore[f].begin(), fore[f].end(), NEG_INF);
            std::fill(back[f].begin(), back[f].end(), NEG_INF);
        }
    }
    // forward max-marginal of every edge of transition f (match2nd.h:190-229)
    void forward(int f) {
        const int *jc = T->jc[f];
        std::vector<double> &m = fore[f];
        if (f > 0) {
            const std::vector<double> &p = fore[f - 1];
            const int *head = T->ir[f - 1];
            for (int e = 0; e < T->nz[f - 1]; ++e) {
                const int j = head[e];
                const double c = p[e];
                for (int k = jc[j]; k != jc[j + 1]; ++k) m[k] = std::max(m[k], c);
            }
        } else {
            for (int i = 0; i < T->nodes(0); ++i)
                for (int e = jc[i]; e != jc[i + 1]; ++e) m[e] = node(0, i);
        }
        const int *head = T->ir[f];
        const double *val = T->pr[f];
        for (int e = 0; e < T->nz[f]; ++e) m[e] += node(f + 1, head[e]) + val[e];
    }
    // backward max-marginal of every edge of transition f (match2nd.h:231-267)
    void backward(int f) {
        const int *jc = T->jc[f], *head = T->ir[f];
        const double *val = T->pr[f];
        std::vector<double> &m = back[f];
        if (f < T->F - 2) {
            const std::vector<double> &nx = back[f + 1];
            const int *jn = T->jc[f + 1];
            for (int e = 0; e < T->nz[f]; ++e) {
                const int j = head[e];
                for (int k = jn[j]; k != jn[j + 1]; ++k) m[e] = std::max(m[e], nx[k]);
            }
        } else {
            for (int e = 0; e < T->nz[f]; ++e) m[e] = node(f + 1, head[e]);
        }
        for (int i = 0; i < T->nodes(f); ++i)
            for (int e = jc[i]; e != jc[i + 1]; ++e) m[e] += val[e] + node(f, i);
    }
    // best / second-best path score through frames f and f+1 and where the best sits (match2nd.h:270-325).  bestloc
    // deliberately carries over from the previous call / previous track use, as in the reference.
    void find_best(int f) {
        best[f] = best[f + 1] = second[f] = second[f + 1] = NEG_INF;
        const int *jc = T->jc[f], *head = T->ir[f];
        const double *val = T->pr[f];
        int loc = 0;
        for (int e = 0; e < T->nz[f]; ++e) {
            while (jc[loc + 1] <= e) ++loc;
            const int to = head[e];
            const double t = back[f][e] + fore[f][e] - node(f, loc) - node(f + 1, to) - val[e];
            if (best[f] < t) {
                if (bestloc[f] == loc)
                    best[f] = t;
                else {
                    second[f] = best[f];
                    best[f] = t;
                    bestloc[f] = loc;
                }
                if (bestloc[f + 1] == to)
                    best[f + 1] = t;
                else {
                    second[f + 1] = best[f + 1];
                    best[f + 1] = t;
                    bestloc[f + 1] = to;
                }
            } else {
                if (second[f] < t && bestloc[f] != loc) second[f] = t;
                if (second[f + 1] < t && bestloc[f + 1] != to) second[f + 1] = t;
            }
        }
    }
    void margin(Trellis &W) {  // match2nd.h:393-413
        clear();
        for (int f = 0; f + 1 < T->F; ++f) forward(f);
        for (int f = T->F - 2; f >= 0; --f) backward(f);
        for (int f = 0; f + 1 < T->F; ++f) find_best(f);
        for (int f = 0; f < T->F; ++f)  // update_unary_first (183-189): only candidates are penalised
            if (bestloc[f] < T->n[f]) W.message[f][bestloc[f]] += second[f] - best[f] - T->bam;
    }
    // greedy decoding along the backward marginals (match2nd.h:327-377)
    void decode(int f) {
        if (f == 1) return;
        double top = NEG_INF;
        if (f == 0) {
            const int *jc = T->jc[0], *head = T->ir[0];
            for (int loc = 0; loc < T->nodes(0); ++loc)
                for (int e = jc[loc]; e != jc[loc + 1]; ++e) {
                    const double c = back[0][e];
                    if (top <= c) {
                        top = c;
                        label[0] = loc;
                        label[1] = head[e];
                    }
                }
            if (top == NEG_INF) label[0] = label[1] = -1;
            return;
        }
        // the reference scans transition f-2 for the edge (label[f-2] -> label[f-1]) and, where it finds it, relaxes the
        // edges leaving label[f-1]; with negative labels nothing matches
        const int a = label[f - 2], b = label[f - 1];
        if (a >= 0 && b >= 0 && a < T->nodes(f - 2)) {
            const int *jp = T->jc[f - 2], *hp = T->ir[f - 2];
            const int *jc = T->jc[f - 1], *head = T->ir[f - 1];
            const double *val = T->pr[f - 1];
            for (int e = jp[a]; e != jp[a + 1]; ++e) {
                if (hp[e] != b) continue;
                for (int k = jc[b]; k != jc[b + 1]; ++k) {
                    const double c = val[k] + back[f - 1][k];
                    if (top < c) {
                        top = c;
                        label[f] = head[k];
                    }
                }
            }
        }
        if (top == NEG_INF) label[f] = -1;
    }
    void assign(Trellis &W) {  // match2nd.h:435-460
        clear();
        for (int f = 0; f < T->F; ++f) {  // update_unary_second_pre (379-384): take the own penalty back
            if (T->bam == std::numeric_limits<double>::infinity())
                W.message[f][bestloc[f]] = 0;
            else if (bestloc[f] < T->n[f])
                W.message[f][bestloc[f]] += -second[f] + best[f] + T->bam;
        }
        for (int f = T->F - 2; f >= 0; --f) backward(f);
        for (int f = 0; f < T->F; ++f) decode(f);
        for (int f = 0; f < T->F; ++f)  // update_unary_second (386-391): the chosen candidates are taken
            if (label[f] >= 0 && label[f] < T->n[f]) W.message[f][label[f]] += NEG_INF;
    }
};

cv::Mat solve(const std::vector<MyMat> &unary_costs, const std::vector<MATSPARSE> &pairwise_costs, int Nong, double occlusion_point_cost,
              double bam_tie, unsigned int frames, unsigned int points, const int *permutation) {
    cv::Mat out = cv::Mat::zeros((int)points, (int)frames, CV_32SC1);
    if (frames < 2 || points < 1) {
        std::cout << "There must be at least one point and 2 frames." << std::endl;
        return out;
    }
    if (unary_costs.size() < frames || pairwise_costs.size() + 1 < frames || !permutation) return out;
    Trellis W;
    W.F = (int)frames;
    W.occ = Nong;
    W.occ_score = occlusion_point_cost;
    W.bam = bam_tie;
    W.n.resize(frames);
    for (unsigned int f = 0; f < frames; ++f) {
        if ((unsigned int)unary_costs[f].Ncols() != points) {
            std::cout << "Wrong form of unary potentials." << unary_costs[f].Ncols() << "!=" << points << std::endl;
            return out;
        }
        W.n[f] = unary_costs[f].Nrows();
    }
    W.jc.resize(frames - 1);
    W.ir.resize(frames - 1);
    W.pr.resize(frames - 1);
    W.nz.resize(frames - 1);
    for (unsigned int f = 0; f + 1 < frames; ++f) {
        const MATSPARSE &S = pairwise_costs[f];
        if (S.Nrows() != W.n[f + 1] + W.occ || S.Ncols() != W.n[f] + W.occ) return out;  // match2nd.cpp:84-97
        W.jc[f] = S.getJc();
        W.ir[f] = S.getIr();
        W.pr[f] = S.getPr();
        W.nz[f] = S.nz();
    }
    W.message.resize(frames);
    for (unsigned int f = 0; f < frames; ++f) W.message[f].assign((size_t)W.nodes((int)f), 0.0);
    std::vector<Track> tracks;
    tracks.reserve(points);
    for (unsigned int p = 0; p < points; ++p) {
        tracks.emplace_back(&W);
        for (unsigned int f = 0; f < frames; ++f)  // column permutation[p] of the column-major unary matrix (match2nd.cpp:59-61)
            tracks[p].unary[f] = unary_costs[f].getValues() + (size_t)permutation[p] * (size_t)W.n[f];
    }
    for (unsigned int p = 0; p < points; ++p) tracks[p].margin(W);          // bundle::run, match2nd.h:525-551
    for (int p = (int)points - 1; p >= 0; --p) tracks[(size_t)p].assign(W);
    for (unsigned int p = 0; p < points; ++p) {
        int *row = out.ptr<int>((int)p);
        for (unsigned int f = 0; f < frames; ++f) row[f] = tracks[p].label[f];
    }
    return out;
}

}  // namespace

cv::Mat match2nd(const std::vector<MyMat> &unary_costs, const std::vector<MATSPARSE> &pairwise_costs, int Nong, double occlusion_point_cost,
                 double bam_tie, unsigned int frames, unsigned int points, const int *permutation) {
    return solve(unary_costs, pairwise_costs, Nong, occlusion_point_cost, bam_tie, frames, points, permutation);
}

// match2nd.cpp:162-191.  Two properties of the reference are kept on purpose: the loop runs over exactly four tracks, and
// the pairwise term adds nothing, because the reference's MATSPARSE::get returns 0 before it looks anything up
// (MyMat.cpp:371-374).  One is not: a label of -1 makes the reference read outside the unary matrix (undefined
// behaviour with NDEBUG); here such a frame contributes 0.
double computeCostTrack(const cv::Mat &M, const std::vector<MyMat> &unary_costs, const std::vector<MATSPARSE> &pairwise_costs, const int *permutation) {
    (void)pairwise_costs;
    double c = 0;
    const int n_frames = M.cols;
    for (int t = 0; t < 4 && t < M.rows; ++t) {
        const int *row = M.ptr<int>(t);
        for (int f = 0; f < n_frames; ++f) {
            double un = 0;
            if (row[f] >= 0 && row[f] < unary_costs[(size_t)f].Nrows()) un = unary_costs[(size_t)f].get((unsigned int)row[f], (unsigned int)permutation[t]);
            c += un;
            if (f < n_frames - 1) c += 0.0;
        }
    }
    return c;
}

namespace lm_track {

MATSPARSE side_view_transitions(const std::vector<unsigned int> &Zi, const std::vector<unsigned int> &Zip1, double grid_mapping,
                                double grid_spacing, unsigned int Nong, double max_displacement, double alpha_vel, double pairwise_occluded_cost) {
    const int Ni = (int)Zi.size(), Nip1 = (int)Zip1.size();
    const double occluded = pairwise_occluded_cost * alpha_vel;
    const int last = (int)Nong - 1;
    // matchToRange (LocoMouse_class.hpp:364-374) of round((grid_mapping - z) / grid_spacing): the grid node next to row z
    auto node_of = [&](unsigned int z) {
        const int32_t q = (int32_t)std::round((grid_mapping - (double)z) / grid_spacing);
        return q < 0 ? 0 : (q > last ? last : q);
    };
    MyMat D((unsigned int)(Nip1 + (int)Nong), (unsigned int)(Ni + (int)Nong));
    for (int i = 0; i < Ni; ++i) {
        D.put((unsigned int)(Nip1 + node_of(Zi[(size_t)i])), (unsigned int)i, occluded);         // candidate -> grid
        for (int j = 0; j < Nip1; ++j) {
            const double dist = std::abs((double)Zip1[(size_t)j] - (double)Zi[(size_t)i]);
            if (dist < max_displacement) D.put((unsigned int)j, (unsigned int)i, (1 - (dist / max_displacement)) * alpha_vel);
        }
    }
    for (int j = 0; j < Nip1; ++j) D.put((unsigned int)j, (unsigned int)(Ni + node_of(Zip1[(size_t)j])), occluded);  // grid -> candidate
    for (int i = 0; i < (int)Nong; ++i) D.put((unsigned int)(Nip1 + i), (unsigned int)(Ni + i), occluded);          // grid -> grid
    return MATSPARSE(&D);
}

void match2nd_concurrent(std::vector<Job> &jobs, unsigned int n_threads) {
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = (unsigned int)std::min<size_t>(n_threads, jobs.size());
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t i = next.fetch_add(1); i < jobs.size(); i = next.fetch_add(1)) {
            Job &j = jobs[i];
            j.result = solve(*j.unary, *j.pairwise, j.Nong, j.occlusion_point_cost, j.bam_tie, j.frames, j.points, j.permutation);
        }
    };
    if (n_threads <= 1) {
        work();
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned int t = 0; t < n_threads; ++t) pool.emplace_back(work);
    for (std::thread &t : pool) t.join();
}

}  // namespace lm_track
