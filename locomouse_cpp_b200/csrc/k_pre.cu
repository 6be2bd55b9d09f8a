// k_pre.cu — frame pre-processing: background subtraction, per-frame min/max, normalisation,
// calibration gather, flip, imadjust, and the bounding-box crop of both views.
//
// Replaces LocoMouse::readFrame (LocoMouse_class.cpp:1273-1333: subtract 1304, normalize 1310,
// correctImage 1337-1406, flip 1323), LocoMouse_TM::readFrame's imadjust (LocoMouse_TM.cpp:243-249,
// LUT at LocoMouse_class.cpp:3223-3241), cropBoundingBox (1408-1478) and storePreviousImage (1508):
// nothing but the crop windows the detectors read is ever materialised, and the "previous image"
// is simply the previous raw frame, re-derived on the fly by the pairing kernel.
//
//  k_minmax : the only pass over whole raw frames (HBM bound): 16-byte loads of F and BKG,
//             per-byte saturating subtract with SIMD-in-word intrinsics, warp + atomic reduction.
//  k_lut    : per frame 256-entry LUT = imadjust ∘ saturate_u8(rint(fma(d, a, b))) with
//             a = float(255 * (1/(max-min))), b = float(-min * scale) (cv::normalize semantics).
//  k_prep   : per frame and view, window = crop + template halo; window pixel (r,c) in calibrated
//             image coordinates -> raw index through the calibration map (with the flip folded in) ->
//             lut[sat(F - BKG)]; zero outside the image (the reference's zero padded I_PAD).
#include "lm_internal.h"

namespace {

__device__ __forceinline__ const uint8_t *frame_ptr(const LmBatch &b, int i) {  // i in [-1, B)
    return i < 0 ? b.prev : b.frames + (int64_t)i * b.frame_bytes;
}

// ---- k_minmax ------------------------------------------------------------------------------------
// minmax[slot] = (255 - min, max) so that both reduce with atomicMax from a zero-initialised buffer.
__global__ void __launch_bounds__(256) k_minmax(const __grid_constant__ LmBatch b, int slot0) {
    const int slot = slot0 + blockIdx.y;  // slot 0 = halo frame (-1)
    const uint8_t *F = frame_ptr(b, slot - 1);
    if (!F) return;
    const int64_t n = b.frame_bytes;
    const uint8_t *K = b.bkg;
    uint32_t mn = 0xffffffffu, mx = 0u;
    const bool vec = ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(K)) & 15) == 0;
    const int64_t nvec = vec ? (n >> 4) : 0;
    const uint4 *F4 = reinterpret_cast<const uint4 *>(F);
    const uint4 *K4 = reinterpret_cast<const uint4 *>(K);
    // four independent 16-byte loads of the frame (and of the L2-resident background) in flight per thread: the pass is
    // HBM-bound and a dependent one-load-per-iteration loop leaves the memory pipeline half empty
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 f[4], k[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = __ldcs(F4 + i + q * stride);  // streamed once: do not keep in L2
#pragma unroll
        for (int q = 0; q < 4; ++q) k[q] = __ldg(K4 + i + q * stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t d0 = __vsubus4(f[q].x, k[q].x), d1 = __vsubus4(f[q].y, k[q].y), d2 = __vsubus4(f[q].z, k[q].z), d3 = __vsubus4(f[q].w, k[q].w);
            mn = __vminu4(mn, __vminu4(__vminu4(d0, d1), __vminu4(d2, d3)));
            mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(d0, d1), __vmaxu4(d2, d3)));
        }
    }
    for (; i < nvec; i += stride) {
        uint4 f = __ldg(F4 + i), k = __ldg(K4 + i);
        uint32_t d0 = __vsubus4(f.x, k.x), d1 = __vsubus4(f.y, k.y), d2 = __vsubus4(f.z, k.z), d3 = __vsubus4(f.w, k.w);
        mn = __vminu4(mn, __vminu4(__vminu4(d0, d1), __vminu4(d2, d3)));
        mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(d0, d1), __vmaxu4(d2, d3)));
    }
    for (int64_t t = (nvec << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        int d = (int)F[t] - (int)K[t];
        uint32_t u = d < 0 ? 0u : (uint32_t)d;
        mn = __vminu4(mn, u * 0x01010101u);
        mx = __vmaxu4(mx, u * 0x01010101u);
    }
    uint32_t lo = min(min(mn & 0xff, (mn >> 8) & 0xff), min((mn >> 16) & 0xff, mn >> 24));
    uint32_t hi = max(max(mx & 0xff, (mx >> 8) & 0xff), max((mx >> 16) & 0xff, mx >> 24));
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    __shared__ uint32_t slo[8], shi[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        slo[w] = lo;
        shi[w] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) {
            lo = min(lo, slo[q]);
            hi = max(hi, shi[q]);
        }
        atomicMax(&b.minmax[slot * 2 + 0], (int)(255u - lo));
        atomicMax(&b.minmax[slot * 2 + 1], (int)hi);
    }
}

// ---- k_lut ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lut(const __grid_constant__ LmBatch b, int slot0) {
    const int slot = slot0 + blockIdx.x;
    const int v = threadIdx.x;
    const int smin = 255 - b.minmax[slot * 2 + 0], smax = b.minmax[slot * 2 + 1];
    // cv::normalize(NORM_MINMAX, 0..255) then Mat::convertTo(CV_8U, scale, shift):
    const double range = (double)smax - (double)smin;
    const double scale = __dmul_rn(255.0, range > 2.2204460492503131e-16 ? __ddiv_rn(1.0, range) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn((double)smin, scale));
    const float a = __double2float_rn(scale), sh = __double2float_rn(shift);
    int q = __float2int_rn(__fmaf_rn((float)v, a, sh));  // cvRound: half to even
    q = min(255, max(0, q));
    if (b.imadjust) {
        // imadjust(I, I, 0, 0.6, 0, 1): round() is half away from zero (values are non-negative)
        const double high_in = __dmul_rn(0.6, 255.0);
        const double range_div = __ddiv_rn(255.0, high_in);
        double t;
        if ((double)q <= 0.0)
            t = 0.0;
        else if ((double)q >= high_in)
            t = 255.0;
        else
            t = __dmul_rn((double)q, range_div);
        q = (int)round(t);
        q = min(255, max(0, q));
    }
    b.lut[slot * 256 + v] = (uint8_t)q;
}

// ---- k_prep --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ LmBatch b) {
    const int f = blockIdx.y, v = blockIdx.z;
    const LmView &V = b.view[v];
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = b.lut[(f + 1) * 256 + threadIdx.x];
    __syncthreads();
    const uint8_t *F = b.frames + (int64_t)f * b.frame_bytes;
    const int x0 = (int)b.bb_x[f] - b.bb_w + 1 - V.halo_x;
    const int ypos = (int)(v == LM_BOTTOM ? b.bb_y_bottom[f] : b.bb_y_side[f]);
    const int y0 = ypos - V.box_h + 1 - V.halo_y;
    uint8_t *W = b.win[v] + (int64_t)f * V.win_stride;
    const int words_per_row = V.win_pitch >> 2;
    const int nwords = words_per_row * V.win_h;
    for (int wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nwords; wi += gridDim.x * blockDim.x) {
        const int r = wi / words_per_row, c4 = (wi - r * words_per_row) << 2;
        const int yy = y0 + r;
        uint32_t out = 0;
        if (yy >= 0 && yy < b.n_rows) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c4 + q, xx = x0 + c;
                if (c < V.win_w && xx >= 0 && xx < b.n_cols) {
                    const int xs = b.flip ? (b.n_cols - 1 - xx) : xx;
                    const int idx = __ldg(b.calib + (int64_t)yy * b.n_cols + xs);
                    int d = (int)__ldg(F + idx) - (int)__ldg(b.bkg + idx);
                    d = d < 0 ? 0 : d;
                    out |= (uint32_t)lut[d] << (8 * q);
                }
            }
        }
        reinterpret_cast<uint32_t *>(W + (int64_t)r * V.win_pitch)[c4 >> 2] = out;
    }
}

}  // namespace

int lm_launch_minmax(const LmBatch &b, cudaStream_t s) {
    // slots 0..B (slot 0 = halo frame).  Enough CTAs per frame to keep HBM busy at any batch size.
    const int slot0 = b.prev ? 0 : 1;
    const int nslots = b.B + 1 - slot0;
    if (nslots <= 0) return 0;
    int64_t nvec = b.frame_bytes >> 4;
    int bx = (int)((nvec + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    k_minmax<<<dim3(bx, nslots), 256, 0, s>>>(b, slot0);
    k_lut<<<nslots, 256, 0, s>>>(b, slot0);
    return 2;
}

int lm_launch_prep(const LmBatch &b, cudaStream_t s) {
    int maxwords = 0;
    for (int v = 0; v < 2; ++v) {
        int w = (b.view[v].win_pitch >> 2) * b.view[v].win_h;
        if (w > maxwords) maxwords = w;
    }
    int bx = (maxwords + 256 * 4 - 1) / (256 * 4);
    if (bx < 1) bx = 1;
    k_prep<<<dim3(bx, b.B, 2), 256, 0, s>>>(b);
    return 1;
}
