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
// The only pass over whole raw frames, hence HBM-bound.  A CTA owns an 8 kB pixel range (two 16-byte vectors per thread)
// and walks MM_GROUP consecutive frames with the background vectors of that range held in registers, so the L2-resident
// background costs 1/MM_GROUP of the frame traffic instead of doubling it; two frames' loads are in flight per thread
// (streaming loads: a frame is read once).  Per-frame partials go warp -> shared memory, one barrier per CTA at the end.
constexpr int MM_GROUP = 8;
constexpr int MM_VEC = 2;

// One 32-bit word = four pixels.  Bytes are widened to s16x2 lanes (even / odd bytes) and go through the DPX
// instructions: d = max(f + (-k), 0) is one VIADDMNMX.S16x2.RELU, the running min / max one VIMNMX3.S16x2 each -- six
// instructions per four pixels (the byte-SIMD intrinsics __vsubus4 / __vminu4 are emulated with dozens of instructions
// on this architecture and made the pass ALU-bound at 40 % of the HBM rate).  nke / nko: the background's even / odd
// bytes negated per 16-bit lane, prepared once per CTA.
__device__ __forceinline__ void mm_word(uint32_t &mn, uint32_t &mx, uint32_t f, uint32_t nke, uint32_t nko) {
    const uint32_t fe = f & 0x00ff00ffu, fo = (f >> 8) & 0x00ff00ffu;
    const uint32_t de = __viaddmax_s16x2_relu(fe, nke, 0u), dd = __viaddmax_s16x2_relu(fo, nko, 0u);
    mn = __vimin3_s16x2(mn, de, dd);
    mx = __vimax3_s16x2(mx, de, dd);
}
struct NegBkg {
    uint32_t e[4], o[4];
};
__device__ __forceinline__ NegBkg mm_neg(const uint4 &k) {
    NegBkg r;
    const uint32_t w[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        r.e[q] = __vneg2(w[q] & 0x00ff00ffu);
        r.o[q] = __vneg2((w[q] >> 8) & 0x00ff00ffu);
    }
    return r;
}
__device__ __forceinline__ void mm_acc(uint32_t &mn, uint32_t &mx, const uint4 &f, const NegBkg &k) {
    mm_word(mn, mx, f.x, k.e[0], k.o[0]);
    mm_word(mn, mx, f.y, k.e[1], k.o[1]);
    mm_word(mn, mx, f.z, k.e[2], k.o[2]);
    mm_word(mn, mx, f.w, k.e[3], k.o[3]);
}

__global__ void __launch_bounds__(256, 6) k_minmax(const __grid_constant__ LmBatch b, int slot0, int nslots) {
    __shared__ uint32_t slo[MM_GROUP][8], shi[MM_GROUP][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t n = b.frame_bytes;
    const uint8_t *K = b.bkg;
    const int g0 = blockIdx.y * MM_GROUP;
    const int ng = min(MM_GROUP, nslots - g0);
    // all frames of a batch share alignment (contiguous, frame_bytes apart) unless frame_bytes is odd; checked per frame
    const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * MM_VEC;  // first 16-byte vector of this thread
    const int64_t nvec = n >> 4;
    NegBkg k[MM_VEC];
    const bool kvec = (reinterpret_cast<uintptr_t>(K) & 15) == 0;
#pragma unroll
    for (int q = 0; q < MM_VEC; ++q)
        k[q] = mm_neg((kvec && v0 + q < nvec) ? __ldg(reinterpret_cast<const uint4 *>(K) + v0 + q) : make_uint4(0, 0, 0, 0));
    for (int g = 0; g < ng; ++g) {
        const int slot = slot0 + g0 + g;  // slot 0 = halo frame (-1)
        const uint8_t *F = frame_ptr(b, slot - 1);
        uint32_t mn = 0x00ff00ffu, mx = 0u;  // s16x2 lanes
        if (F) {
            if (kvec && (reinterpret_cast<uintptr_t>(F) & 15) == 0) {
                const uint4 *F4 = reinterpret_cast<const uint4 *>(F);
                uint4 f[MM_VEC];
#pragma unroll
                for (int q = 0; q < MM_VEC; ++q) f[q] = (v0 + q < nvec) ? __ldcs(F4 + v0 + q) : make_uint4(0, 0, 0, 0);  // streamed once
#pragma unroll
                for (int q = 0; q < MM_VEC; ++q)
                    if (v0 + q < nvec) mm_acc(mn, mx, f[q], k[q]);
                // bytes after the last whole vector: the thread that owns the vector slot right after them
                if (v0 <= nvec && nvec < v0 + MM_VEC)
                    for (int64_t t = nvec << 4; t < n; ++t) {
                        const int d = (int)F[t] - (int)K[t];
                        const uint32_t u = d < 0 ? 0u : (uint32_t)d;
                        mn = __vimin3_s16x2(mn, u * 0x00010001u, u * 0x00010001u);
                        mx = __vimax3_s16x2(mx, u * 0x00010001u, u * 0x00010001u);
                    }
            } else {  // unaligned frames (odd frame size or caller pointer): bytewise over this thread's range
                const int64_t t0 = v0 << 4, t1 = min(n, (v0 + MM_VEC) << 4);
                for (int64_t t = t0; t < t1; ++t) {
                    const int d = (int)F[t] - (int)K[t];
                    const uint32_t u = d < 0 ? 0u : (uint32_t)d;
                    mn = __vimin3_s16x2(mn, u * 0x00010001u, u * 0x00010001u);
                    mx = __vimax3_s16x2(mx, u * 0x00010001u, u * 0x00010001u);
                }
            }
        }
        uint32_t lo = min(mn & 0xffffu, mn >> 16);
        uint32_t hi = max(mx & 0xffffu, mx >> 16);
        lo = __reduce_min_sync(0xffffffffu, lo);   // REDUX: one instruction per warp reduction
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            slo[g][w] = lo;
            shi[g][w] = hi;
        }
    }
    __syncthreads();
    if (threadIdx.x < ng) {
        const int g = threadIdx.x, slot = slot0 + g0 + g;
        if (frame_ptr(b, slot - 1)) {
            uint32_t lo = slo[g][0], hi = shi[g][0];
            for (int q = 1; q < 8; ++q) {
                lo = min(lo, slo[g][q]);
                hi = max(hi, shi[g][q]);
            }
            atomicMax(&b.minmax[slot * 2 + 0], (int)(255u - lo));
            atomicMax(&b.minmax[slot * 2 + 1], (int)hi);
        }
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
// Per video, once (k_fold_calib): the calibration map with the optional mirror folded in, calib_flip[r][c] =
// calib[r][flip ? n_cols - 1 - c : c], and the background seen through it, bkg_warp[r][c] = bkg[calib_flip[r][c]].  A window
// pixel is then lut[max(F[calib_flip] - bkg_warp, 0)]: one gather (the frame) instead of three.
// k_prep: one CTA = PREP_ROWS window rows of one (frame, view).  A thread owns one 4-pixel word column and walks the rows
// (32-bit incremental indexing); words that lie wholly inside the image -- all but the box border -- take a path without
// per-pixel predicates: four map loads, four frame bytes, the background word by two aligned loads + funnel shift.
constexpr int PREP_ROWS = 64;
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(256) k_fold_calib(const int32_t *__restrict__ calib, const uint8_t *__restrict__ bkg, int n_rows, int n_cols,
                                                    int flip, int32_t *__restrict__ calib_flip, uint8_t *__restrict__ bkg_warp) {
    const int64_t n = (int64_t)n_rows * n_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n_cols), c = (int)(i - (int64_t)r * n_cols);
        const int32_t src = calib[(int64_t)r * n_cols + (flip ? n_cols - 1 - c : c)];
        calib_flip[i] = src;
        bkg_warp[i] = bkg[src];
    }
}

__global__ void __launch_bounds__(PREP_THREADS) k_prep(const __grid_constant__ LmBatch b) {
    const int f = blockIdx.y, v = blockIdx.z;
    const LmView &V = b.view[v];
    const int r0 = blockIdx.x * PREP_ROWS;
    if (r0 >= V.win_h) return;
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = b.lut[(f + 1) * 256 + i];
    __syncthreads();
    const int r1 = min(V.win_h, r0 + PREP_ROWS);
    const uint8_t *__restrict__ F = b.frames + (int64_t)f * b.frame_bytes;
    const uint8_t *__restrict__ Kw = b.bkg_warp;
    const int32_t *__restrict__ C2 = b.calib_flip;
    const int x0 = (int)b.bb_x[f] - b.bb_w + 1 - V.halo_x;
    const int ypos = (int)(v == LM_BOTTOM ? b.bb_y_bottom[f] : b.bb_y_side[f]);
    const int y0 = ypos - V.box_h + 1 - V.halo_y;
    uint8_t *W = b.win[v] + (int64_t)f * V.win_stride;
    const int wpr = V.win_pitch >> 2, n_cols = b.n_cols, n_rows = b.n_rows;
    // thread -> (row slot, word column); when a row has more words than the CTA has threads the words are looped over
    const int nslot = wpr <= PREP_THREADS ? PREP_THREADS / wpr : 1;
    const int slot = wpr <= PREP_THREADS ? (int)threadIdx.x / wpr : 0;
    if (slot >= nslot) return;
    for (int w = wpr <= PREP_THREADS ? (int)threadIdx.x - slot * wpr : (int)threadIdx.x; w < wpr; w += PREP_THREADS) {
        const int c4 = w << 2, xx0 = x0 + c4;
        // bit q: pixel q of the word is a window pixel that lies inside the image
        unsigned vm = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c4 + q < V.win_w && xx0 + q >= 0 && xx0 + q < n_cols) vm |= 1u << q;
        uint32_t *dst = reinterpret_cast<uint32_t *>(W + (int64_t)(r0 + slot) * V.win_pitch) + w;
        const int dst_step = nslot * wpr;  // words
        int yy = y0 + r0 + slot;
        if (vm == 0xFu) {
#pragma unroll 2
            for (int r = r0 + slot; r < r1; r += nslot, yy += nslot, dst += dst_step) {
                uint32_t out = 0;
                if (yy >= 0 && yy < n_rows) {
                    const int base = yy * n_cols + xx0;
                    const int32_t *cp = C2 + base;
                    const int i0 = __ldg(cp), i1 = __ldg(cp + 1), i2 = __ldg(cp + 2), i3 = __ldg(cp + 3);
                    const uint32_t *kp = reinterpret_cast<const uint32_t *>(Kw + (base & ~3));  // bkg_warp is 4-byte aligned and padded
                    const uint32_t kw4 = __funnelshift_r(__ldg(kp), __ldg(kp + 1), (base & 3) * 8);
                    const int d0 = max((int)__ldg(F + i0) - (int)(kw4 & 0xffu), 0);
                    const int d1 = max((int)__ldg(F + i1) - (int)((kw4 >> 8) & 0xffu), 0);
                    const int d2 = max((int)__ldg(F + i2) - (int)((kw4 >> 16) & 0xffu), 0);
                    const int d3 = max((int)__ldg(F + i3) - (int)(kw4 >> 24), 0);
                    out = (uint32_t)lut[d0] | ((uint32_t)lut[d1] << 8) | ((uint32_t)lut[d2] << 16) | ((uint32_t)lut[d3] << 24);
                }
                *dst = out;
            }
        } else {
            for (int r = r0 + slot; r < r1; r += nslot, yy += nslot, dst += dst_step) {
                uint32_t out = 0;
                if (vm && yy >= 0 && yy < n_rows) {
                    const int base = yy * n_cols + xx0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (vm & (1u << q)) {
                            const int d = max((int)__ldg(F + __ldg(C2 + base + q)) - (int)__ldg(Kw + base + q), 0);
                            out |= (uint32_t)lut[d] << (8 * q);
                        }
                }
                *dst = out;
            }
        }
    }
}

}  // namespace

int lm_launch_minmax(const LmBatch &b, cudaStream_t s) {
    // slots 0..B (slot 0 = halo frame).  Enough CTAs per frame to keep HBM busy at any batch size.
    const int slot0 = b.prev ? 0 : 1;
    const int nslots = b.B + 1 - slot0;
    if (nslots <= 0) return 0;
    const int64_t nvec = (b.frame_bytes >> 4) + 1;  // + 1: the slot that owns the bytes after the last whole vector
    const int bx = (int)((nvec + 256 * MM_VEC - 1) / (256 * MM_VEC));
    k_minmax<<<dim3(bx, (nslots + MM_GROUP - 1) / MM_GROUP), 256, 0, s>>>(b, slot0, nslots);
    k_lut<<<nslots, 256, 0, s>>>(b, slot0);
    return 2;
}

int lm_launch_prep(const LmBatch &b, cudaStream_t s) {
    int maxh = 0;
    for (int v = 0; v < 2; ++v) maxh = b.view[v].win_h > maxh ? b.view[v].win_h : maxh;
    const int bx = (maxh + PREP_ROWS - 1) / PREP_ROWS;
    k_prep<<<dim3(bx < 1 ? 1 : bx, b.B, 2), PREP_THREADS, 0, s>>>(b);
    return 1;
}

int lm_launch_fold_calib(const int32_t *calib, const uint8_t *bkg, int n_rows, int n_cols, int flip, int32_t *calib_flip, uint8_t *bkg_warp,
                         cudaStream_t s) {
    k_fold_calib<<<lm_sm_count() * 4, 256, 0, s>>>(calib, bkg, n_rows, n_cols, flip, calib_flip, bkg_warp);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
