// oracle/ref_match2nd_glue.cpp — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own host tracker:
// match2nd() / computeCostTrack() (match2nd/match2nd.cpp, match2nd/match2nd.h) and MyMat / MATSPARSE (MyMat/MyMat.cpp),
// compiled unchanged from where they lie under /root/reference (oracle/Makefile, target _ref/libref_match2nd.so) against the
// value-type shim in ref_shim/.  It pins locomouse_cpp_b200/host/match2nd.cpp (tests/test_match2nd.py).  A library of its
// own because the reference's tracker keeps its state in globals (match2nd.cpp:4-8).
#include <cstdint>
#include <vector>

#include "match2nd.h"  // the reference's header, found through -I/root/reference/match2nd

namespace {
void build(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, int nong, const int32_t *jc,
           const int64_t *jc_off, const int32_t *ir, const double *pr, const int64_t *nz_off, std::vector<MyMat> &U, std::vector<MATSPARSE> &P) {
    for (int f = 0; f < frames; ++f) {
        MyMat M((unsigned int)n_loc[f], (unsigned int)points);
        for (int j = 0; j < points; ++j)
            for (int i = 0; i < n_loc[f]; ++i) M.put((unsigned int)i, (unsigned int)j, unary[unary_off[f] + (int64_t)j * n_loc[f] + i]);
        U.push_back(M);
    }
    for (int f = 0; f + 1 < frames; ++f) {
        const int rows = n_loc[f + 1] + nong, cols = n_loc[f] + nong;
        MyMat D((unsigned int)rows, (unsigned int)cols);  // dense, then the reference's own sparsifier (MyMat.cpp:141-178)
        for (int c = 0; c < cols; ++c)
            for (int e = jc[jc_off[f] + c]; e < jc[jc_off[f] + c + 1]; ++e) D.put((unsigned int)ir[nz_off[f] + e], (unsigned int)c, pr[nz_off[f] + e]);
        MATSPARSE S(&D);
        P.push_back(S);
    }
}
}  // namespace

extern "C" {
int ref_match2nd(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, int nong, const int32_t *jc,
                 const int64_t *jc_off, const int32_t *ir, const double *pr, const int64_t *nz_off, double occ_cost, double bam,
                 const int32_t *permutation, int32_t *labels) {
    std::vector<MyMat> U;
    std::vector<MATSPARSE> P;
    build(frames, points, n_loc, unary, unary_off, nong, jc, jc_off, ir, pr, nz_off, U, P);
    cv::Mat T = match2nd(U, P, nong, occ_cost, bam, (unsigned int)frames, (unsigned int)points, permutation);
    for (int p = 0; p < points; ++p)
        for (int f = 0; f < frames; ++f) labels[(size_t)p * frames + f] = T.ptr<int>(p)[f];
    return 0;
}
double ref_cost_track(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, const int32_t *permutation,
                      const int32_t *labels) {
    std::vector<MyMat> U;
    std::vector<MATSPARSE> P;
    for (int f = 0; f < frames; ++f) {
        MyMat M((unsigned int)n_loc[f], (unsigned int)points);
        for (int j = 0; j < points; ++j)
            for (int i = 0; i < n_loc[f]; ++i) M.put((unsigned int)i, (unsigned int)j, unary[unary_off[f] + (int64_t)j * n_loc[f] + i]);
        U.push_back(M);
    }
    for (int f = 0; f + 1 < frames; ++f) {
        MyMat D(1, 1);
        MATSPARSE S(&D);
        P.push_back(S);
    }
    cv::Mat T(points, frames, CV_32SC1);
    for (int p = 0; p < points; ++p)
        for (int f = 0; f < frames; ++f) T.ptr<int>(p)[f] = labels[(size_t)p * frames + f];
    return computeCostTrack(T, U, P, permutation);
}
}
