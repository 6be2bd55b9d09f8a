// k_bbox.cu — pass 1 of LocoMouse_TM_DE on the device: the per-frame part of
// LocoMouse_TM_DE::computeBoundingBox / computeMouseBox_DE (LocoMouse_TM_DE.cpp:8-113).
//
// The reference reads every frame with the base-class readFrame (subtract, normalise, calibration gather, flip;
// class.cpp:1273-1333), stretches the side view with imadjust_default (3244-3311), zeroes four border bands
// (TM_DE.cpp:68-71), thresholds (75), sums columns (91) and takes the first / last column whose sum reaches
// MIN_PIXEL_COUNT (firstLastOverT, class.hpp:411-442).  Every pixel operation after the gather is a 256-entry
// look-up, so no image is materialised:
//   k_minmax + k_lut (k_pre.cu)  per-frame min/max of sat(F - BKG) -> normalisation LUT  (imadjust(0, 0.6) off)
//   k_bb_hist   histogram of the normalised side view, gathered through the calibration map
//   k_bb_pred   one thread replays imadjust_default's float cumulative scan (order matters), all threads then build
//               pred[d] = ( imadjust_default( normalise(d) ) > threshold )
//   k_bb_cols   column sums of pred over the rows / columns that survive the zeroed bands, first / last, bb_x
// The whole-video moving average that follows is sequential and stays on the host (lm_moving_average).
#include "lm_internal.h"

namespace {

struct BBoxDev {
    const uint8_t *frames;
    int64_t frame_bytes;
    const uint8_t *bkg;
    const int32_t *calib;
    int n_cols, flip, B;
    lm_bb_de_params p;
    const uint8_t *lut;   // [B + 1][256] normalisation LUT, slot f + 1
    uint32_t *hist;       // [B][256]
    uint8_t *pred;        // [B][256]
    double *bb_x;         // [B]
    int32_t *lims;        // [B][2]
};

__global__ void __launch_bounds__(256) k_bb_hist(const __grid_constant__ BBoxDev P) {
    // per-warp private histograms + warp-aggregated increments: a dark side view puts most pixels into a handful of
    // bins, and 32 lanes hitting one shared-memory word would serialise
    __shared__ uint32_t sh[8][256];
    __shared__ uint8_t lut[256];
    const int f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    lut[tid] = P.lut[(f + 1) * 256 + tid];
    __syncthreads();
    const uint8_t *F = P.frames + (int64_t)f * P.frame_bytes;
    const int npx = P.p.side_w * P.p.side_h;
    // each lane takes UNR pixels per step, 32 * UNR consecutive pixels per warp; all calibration loads of a step are
    // issued first, then all frame / background loads: the gather chain is latency-bound otherwise
    constexpr int UNR = 8;
    const int stride = gridDim.x * 256 * UNR;
    for (int i0 = (blockIdx.x * 256 + warp * 32) * UNR; i0 < npx; i0 += stride) {
        int idx[UNR], bin[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int i = i0 + q * 32 + lane;
            idx[q] = -1;
            if (i < npx) {
                const int r = i / P.p.side_w, x = P.p.side_x + (i - r * P.p.side_w);
                const int xs = P.flip ? (P.n_cols - 1 - x) : x;
                idx[q] = __ldg(P.calib + (int64_t)(P.p.side_y + r) * P.n_cols + xs);
            }
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            bin[q] = -1;
            if (idx[q] >= 0) {
                const int d = (int)__ldg(F + idx[q]) - (int)__ldg(P.bkg + idx[q]);
                bin[q] = d < 0 ? 0 : d;
            }
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
            const int b = bin[q] >= 0 ? (int)lut[bin[q]] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            if (b >= 0 && lane == __ffs(peers) - 1) sh[warp][b] += __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sh[w][tid];
    if (tot) atomicAdd(&P.hist[f * 256 + tid], tot);
}

__global__ void __launch_bounds__(256) k_bb_pred(const __grid_constant__ BBoxDev P) {
    __shared__ float s_ab[2];
    __shared__ int s_identity;
    __shared__ uint32_t h[256];
    const int f = blockIdx.x, tid = threadIdx.x;
    h[tid] = P.hist[f * 256 + tid];
    __syncthreads();
    if (tid == 0) {
        // imadjust_default, LocoMouse_class.cpp:3258-3296, operation for operation (float scan, double cv::sum)
        double acc = 0.0;
        for (int i = 0; i < 256; ++i) acc = __dadd_rn(acc, (double)(float)h[i]);
        const float sum_histf = (float)acc;
        float cumsum = 0.f;
        int i0 = 0, i1 = 0, imin = 0, imax = 0;
        bool check_min = true, check_max = true;
        for (int i = 0; i < 256; ++i) {
            cumsum = __fadd_rn(cumsum, (float)h[i]);
            const float cn = __fdiv_rn(cumsum, sum_histf);
            if ((cn > 0.01f) & check_min) {
                i0 = i;
                check_min = false;
                imin = i;
            }
            if ((cn >= 0.99f) & check_max) {
                i1 = i;
                check_max = false;
                imax = i;
            }
            if (!(check_min || check_max)) break;
        }
        if (imin == imax) i1 = 256;
        const float r0 = __fdiv_rn((float)i0, 255.f), r1 = __fdiv_rn((float)i1, 255.f);
        const double s = (double)__fsub_rn(r1, r0);
        const double alpha = __ddiv_rn(1.0, s), beta = __dmul_rn(-(double)r0, alpha);
        s_identity = (fabs(alpha) == 1.0);
        s_ab[0] = (float)alpha;
        s_ab[1] = (float)beta;
    }
    __syncthreads();
    // pred[d]: raw difference d -> normalised pixel -> stretched pixel -> threshold
    const int pn = P.lut[(f + 1) * 256 + tid];
    int q = pn;
    if (!s_identity) {
        q = __float2int_rn(__fmaf_rn((float)pn, s_ab[0], s_ab[1]));
        q = min(255, max(0, q));
    }
    P.pred[f * 256 + tid] = ((double)q > P.p.threshold) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_bb_cols(const __grid_constant__ BBoxDev P) {
    __shared__ uint8_t pred[256];
    __shared__ int s_first, s_last, s_cnt;
    const int f = blockIdx.x, tid = threadIdx.x;
    pred[tid] = P.pred[f * 256 + tid];
    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
        s_cnt = 0;
    }
    __syncthreads();
    const uint8_t *F = P.frames + (int64_t)f * P.frame_bytes;
    const int c0 = max(0, P.p.zero_col_pre), c1 = min(P.p.side_w, P.p.zero_col_post);
    const int r0 = max(0, P.p.zero_row_pre), r1 = min(P.p.side_h, P.p.zero_row_post);
    for (int x = tid; x < P.p.side_w; x += 256) {
        int sum = 0;
        if (x >= c0 && x < c1) {
            // rows in groups of 8 with the two dependent gathers (calibration index -> frame / background bytes) issued
            // back to back: the column sum is latency-bound otherwise
            const int xs = P.flip ? (P.n_cols - 1 - (P.p.side_x + x)) : (P.p.side_x + x);
            for (int r = r0; r < r1; r += 8) {
                int idx[8], d[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) idx[q] = (r + q < r1) ? __ldg(P.calib + (int64_t)(P.p.side_y + r + q) * P.n_cols + xs) : -1;
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = idx[q] >= 0 ? (int)__ldg(F + idx[q]) - (int)__ldg(P.bkg + idx[q]) : -1000;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (d[q] > -1000) sum += pred[d[q] < 0 ? 0 : d[q]];
            }
        }
        if (sum >= P.p.min_count) {  // firstLastOverT: p[i] >= th
            atomicMin(&s_first, x);
            atomicMax(&s_last, x);
            atomicAdd(&s_cnt, 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int first = -1, last = -1;
        if (s_cnt > 0) {
            first = s_first;
            last = s_cnt >= 2 ? s_last : 0;  // with one qualifying column the reference leaves slot 1 at its initial 0
        }
        P.lims[f * 2 + 0] = first;
        P.lims[f * 2 + 1] = last;
        P.bb_x[f] = fmin((double)(P.p.side_w - 1), __dmul_rn((double)last, P.p.width_margin));
    }
}

}  // namespace

static BBoxDev bbox_dev(const LmBatch &b, const lm_bb_de_params &p, uint32_t *hist, uint8_t *pred, double *bb_x, int32_t *lims) {
    BBoxDev P{};
    P.frames = b.frames;
    P.frame_bytes = b.frame_bytes;
    P.bkg = b.bkg;
    P.calib = b.calib;
    P.n_cols = b.n_cols;
    P.flip = b.flip;
    P.B = b.B;
    P.p = p;
    P.lut = b.lut;
    P.hist = hist;
    P.pred = pred;
    P.bb_x = bb_x;
    P.lims = lims;
    return P;
}

// The front end LocoMouse_TM_DE and LocoMouse_TM share: per-frame normalisation LUT, side-view histogram, imadjust_default
// and the threshold, folded into pred[f][d] (d = raw difference).  p: the side view rectangle and the threshold are used.
int lm_launch_bbox_pred(const LmBatch &b, const lm_bb_de_params &p, uint32_t *hist, uint8_t *pred, cudaStream_t s) {
    int launches = 0;
    if (cudaMemsetAsync(b.minmax, 0, (size_t)(b.B + 1) * 2 * sizeof(int32_t), s) != cudaSuccess) return -1;
    if (cudaMemsetAsync(hist, 0, (size_t)b.B * 256 * sizeof(uint32_t), s) != cudaSuccess) return -1;
    int nl = lm_launch_minmax(b, s);
    if (nl < 0) return -1;
    launches += nl;
    const BBoxDev P = bbox_dev(b, p, hist, pred, nullptr, nullptr);
    const int npx = p.side_w * p.side_h;
    int gx = (npx + 256 * 16 - 1) / (256 * 16);
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    k_bb_hist<<<dim3(gx, b.B), 256, 0, s>>>(P);
    k_bb_pred<<<b.B, 256, 0, s>>>(P);
    launches += 2;
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

// b: frames / bkg / calib / minmax / lut / n_cols / flip / B filled in by the caller (imadjust must be 0)
int lm_launch_bbox_tm_de(const LmBatch &b, const lm_bb_de_params &p, uint32_t *hist, uint8_t *pred, double *bb_x, int32_t *lims,
                         cudaStream_t s) {
    int launches = lm_launch_bbox_pred(b, p, hist, pred, s);
    if (launches < 0) return -1;
    const BBoxDev P = bbox_dev(b, p, hist, pred, bb_x, lims);
    k_bb_cols<<<b.B, 256, 0, s>>>(P);
    launches += 1;
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}
