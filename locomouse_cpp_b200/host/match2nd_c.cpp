// match2nd_c.cpp — plain-C entry points of the host tracker (match2nd.hpp) for the Python tests: the trellis is passed as
// packed arrays, exactly what lm_unary_costs / lm_pairwise_costs return and what the reference-compiled checker
// (oracle/ref_glue.cpp: ref_match2nd) takes, so both are driven with identical inputs.
#include <cstdint>
#include <vector>

#include "cv_yaml.hpp"
#include "lm_media.hpp"
#include "match2nd.hpp"

namespace {
void build(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, int nong, const int32_t *jc,
           const int64_t *jc_off, const int32_t *ir, const double *pr, const int64_t *nz_off, std::vector<MyMat> &U, std::vector<MATSPARSE> &P) {
    U.reserve((size_t)frames);
    for (int f = 0; f < frames; ++f) {
        MyMat M((unsigned int)n_loc[f], (unsigned int)points);
        for (int64_t i = 0; i < (int64_t)n_loc[f] * points; ++i) M.getValues()[i] = unary[unary_off[f] + i];
        U.push_back(std::move(M));
    }
    for (int f = 0; f + 1 < frames; ++f)
        P.push_back(MATSPARSE(n_loc[f + 1] + nong, n_loc[f] + nong, jc + jc_off[f], ir + nz_off[f], pr + nz_off[f]));
}
}  // namespace

extern "C" {
// unary: frame f's n_loc[f] x points column-major block at unary_off[f]; transition f: jc at jc_off[f] (n_loc[f] + nong + 1
// entries, starting at 0), ir / pr at nz_off[f].  labels: points x frames, row-major.
int lmh_match2nd(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, int nong, const int32_t *jc,
                 const int64_t *jc_off, const int32_t *ir, const double *pr, const int64_t *nz_off, double occ_cost, double bam,
                 const int32_t *permutation, int32_t *labels) {
    std::vector<MyMat> U;
    std::vector<MATSPARSE> P;
    build(frames, points, n_loc, unary, unary_off, nong, jc, jc_off, ir, pr, nz_off, U, P);
    const cv::Mat T = match2nd(U, P, nong, occ_cost, bam, (unsigned int)frames, (unsigned int)points, permutation);
    for (int p = 0; p < points; ++p)
        for (int f = 0; f < frames; ++f) labels[(size_t)p * frames + f] = T.at<int>(p, f);
    return 0;
}
double lmh_cost_track(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, const int32_t *permutation,
                      const int32_t *labels) {
    std::vector<MyMat> U;
    std::vector<MATSPARSE> P;
    for (int f = 0; f < frames; ++f) {
        MyMat M((unsigned int)n_loc[f], (unsigned int)points);
        for (int64_t i = 0; i < (int64_t)n_loc[f] * points; ++i) M.getValues()[i] = unary[unary_off[f] + i];
        U.push_back(std::move(M));
    }
    cv::Mat T(points, frames, CV_32SC1);
    for (int p = 0; p < points; ++p)
        for (int f = 0; f < frames; ++f) T.at<int>(p, f) = labels[(size_t)p * frames + f];
    return computeCostTrack(T, U, P, permutation);
}
// lm_track::side_view_transitions -> jc[ni + nong + 1], ir / pr [nnz]; dims = {rows, cols, nnz}
int lmh_pairwise_potential_side(const uint32_t *zi, int ni, const uint32_t *zip1, int nip1, double grid_mapping, double spacing, int nong,
                                double max_disp, double alpha_vel, double occluded_cost, int32_t *jc, int32_t *ir, double *pr, int cap, int32_t *dims) {
    const std::vector<unsigned int> A(zi, zi + ni), B(zip1, zip1 + nip1);
    const MATSPARSE S = lm_track::side_view_transitions(A, B, grid_mapping, spacing, (unsigned int)nong, max_disp, alpha_vel, occluded_cost);
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
// cvyaml::Writer: `n` int32 matrices (names separated by '\n', rows / cols per matrix, data back to back) into `path`
int lmh_yaml_write(const char *path, const char *names, int n, const int32_t *rows, const int32_t *cols, const int32_t *data) {
    try {
        cvyaml::Writer w(path);
        std::string all(names);
        size_t pos = 0;
        for (int i = 0; i < n; ++i) {
            const size_t e = all.find('\n', pos);
            const std::string name = all.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            pos = e == std::string::npos ? all.size() : e + 1;
            w.write(name, rows[i], cols[i], data);
            data += (size_t)rows[i] * cols[i];
        }
        return w.good() ? 0 : 1;
    } catch (const std::exception &) {
        return 2;
    }
}
// lm_media.hpp readers: dims = {n, rows, cols}; with out == NULL only the dimensions are returned.  0 ok, 1 error (msg filled)
int lmh_read_avi(const char *path, int32_t *dims, uint8_t *out, char *msg, int msg_cap) {
    try {
        const lmmedia::Video V = lmmedia::read_avi(path);
        dims[0] = V.n;
        dims[1] = V.rows;
        dims[2] = V.cols;
        if (out) std::memcpy(out, V.frames.data(), V.frames.size());
        return 0;
    } catch (const std::exception &e) {
        if (msg && msg_cap > 0) std::snprintf(msg, (size_t)msg_cap, "%s", e.what());
        return 1;
    }
}
int lmh_read_png(const char *path, int32_t *dims, uint8_t *out, char *msg, int msg_cap) {
    try {
        const lmmedia::Image I = lmmedia::read_png_gray(path);
        dims[0] = 1;
        dims[1] = I.rows;
        dims[2] = I.cols;
        if (out) std::memcpy(out, I.px.data(), I.px.size());
        return 0;
    } catch (const std::exception &e) {
        if (msg && msg_cap > 0) std::snprintf(msg, (size_t)msg_cap, "%s", e.what());
        return 1;
    }
}
// the same job `copies` times on `threads` host threads; returns 1 when every result equals the serial one
int lmh_match2nd_concurrent_check(int frames, int points, const int32_t *n_loc, const double *unary, const int64_t *unary_off, int nong,
                                  const int32_t *jc, const int64_t *jc_off, const int32_t *ir, const double *pr, const int64_t *nz_off, double occ_cost,
                                  double bam, const int32_t *permutation, int copies, int threads) {
    std::vector<MyMat> U;
    std::vector<MATSPARSE> P;
    build(frames, points, n_loc, unary, unary_off, nong, jc, jc_off, ir, pr, nz_off, U, P);
    const cv::Mat T = match2nd(U, P, nong, occ_cost, bam, (unsigned int)frames, (unsigned int)points, permutation);
    std::vector<lm_track::Job> jobs((size_t)copies);
    for (lm_track::Job &j : jobs) {
        j.unary = &U;
        j.pairwise = &P;
        j.Nong = nong;
        j.occlusion_point_cost = occ_cost;
        j.bam_tie = bam;
        j.frames = (unsigned int)frames;
        j.points = (unsigned int)points;
        j.permutation = permutation;
    }
    lm_track::match2nd_concurrent(jobs, (unsigned int)threads);
    for (const lm_track::Job &j : jobs)
        for (int p = 0; p < points; ++p)
            for (int f = 0; f < frames; ++f)
                if (j.result.at<int>(p, f) != T.at<int>(p, f)) return 0;
    return 1;
}
}
