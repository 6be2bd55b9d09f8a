// k_cost.cu — the cost builders that feed the host tracker, batched over all frames of a result set.
//   k_unary       : LocoMouse::unaryCostBox (LocoMouse_class.cpp:1909-1952), one thread per (frame, candidate)
//   k_pw_count    : LocoMouse::pairwisePotential (1954-2070) + MATSPARSE(const MyMat*) (MyMat.cpp:141-178): the dense D is
//   k_pw_scan       never built; one warp per frame transition derives, per column, which rows the reference stores
//   k_pw_fill       (count -> per-frame column starts -> exclusive scan over frames -> fill, rows ascending).
// Double precision throughout with the reference's operation order; products that feed an addition are kept unfused
// (__dmul_rn / __dadd_rn), so every value equals the CPU's bit for bit.
#include "lm_internal.h"

namespace {

struct PwDev {
    const lm_cand *cand;     // [n][2][cand_cap] bottom candidates
    const int32_t *ncand;    // [n][2]
    int64_t n;
    int cand_cap, feat;
    lm_pairwise_params p;
    int nong, jc_stride;     // jc_stride = cand_cap + nong + 1
    double occ;              // occluded_cost * alpha_vel
    int32_t *jc;             // [n][jc_stride]
    int64_t *nnz;            // [n] entries per frame, then (after the scan) offs[n + 1]
    int64_t *offs;
    int32_t *ir;
    double *pr;
    int64_t cap;
};

__device__ __forceinline__ int ong_index(const lm_pairwise_params &p, int x, int y) {
    int xc = (int)round((p.grid_x - (double)x) / p.grid_spacing);
    int yc = (int)round((p.grid_y - (double)y) / p.grid_spacing);
    xc = xc < 0 ? 0 : (xc > p.ong_w - 1 ? p.ong_w - 1 : xc);   // matchToRange (LocoMouse_class.hpp:364-374)
    yc = yc < 0 ? 0 : (yc > p.ong_h - 1 ? p.ong_h - 1 : yc);
    return yc * p.ong_w + xc;
}

// value stored at D(j, i) for candidates i of frame f-1 and j of frame f; returns false when nothing is put
__device__ __forceinline__ bool pw_value(const lm_pairwise_params &p, const lm_cand &a, const lm_cand &b, double *v) {
    const double dx = (double)b.x - (double)a.x, dy = (double)b.y - (double)a.y;
    const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    if (!(dist < p.max_displacement)) return false;
    double inv = 1 - (dist / p.max_displacement);
    inv = inv * p.alpha_vel;
    *v = inv;
    return true;
}

// Visits the stored entries of column c in row order.  FILL = false: returns their number.
template <bool FILL>
__device__ __forceinline__ int pw_column(const PwDev &P, const lm_cand *A, int ni, const lm_cand *B, int nip1, int c, int32_t *ir, double *pr) {
    int k = 0;
    if (c < ni) {
        const lm_cand a = A[c];
        for (int j = 0; j < nip1; ++j) {
            double v;
            if (pw_value(P.p, a, B[j], &v) && v != 0) {
                if (FILL) {
                    ir[k] = j;
                    pr[k] = v;
                }
                ++k;
            }
        }
        if (P.occ != 0) {
            if (FILL) {
                ir[k] = nip1 + ong_index(P.p, a.x, a.y);
                pr[k] = P.occ;
            }
            ++k;
        }
    } else {
        const int g = c - ni;
        if (ni > 0 && P.occ != 0)   // the reference writes these inside its i == 0 iteration only (1954-2070, quirk)
            for (int j = 0; j < nip1; ++j)
                if (ong_index(P.p, B[j].x, B[j].y) == g) {
                    if (FILL) {
                        ir[k] = j;
                        pr[k] = P.occ;
                    }
                    ++k;
                }
        if (P.occ != 0) {
            if (FILL) {
                ir[k] = nip1 + g;
                pr[k] = P.occ;
            }
            ++k;
        }
    }
    return k;
}

__global__ void __launch_bounds__(256) k_pw_count(const __grid_constant__ PwDev P) {
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= P.n) return;
    int32_t *jc = P.jc + f * P.jc_stride;
    if (f == 0) {  // no transition into the first frame
        for (int c = lane; c < P.jc_stride; c += 32) jc[c] = 0;
        if (lane == 0) P.nnz[0] = 0;
        return;
    }
    const int ni = P.ncand[(f - 1) * 2 + P.feat], nip1 = P.ncand[f * 2 + P.feat];
    const lm_cand *A = P.cand + ((f - 1) * 2 + P.feat) * P.cand_cap, *B = P.cand + (f * 2 + P.feat) * P.cand_cap;
    const int ncols = ni + P.nong;
    int base = 0;
    for (int c0 = 0; c0 < P.jc_stride - 1; c0 += 32) {
        const int c = c0 + lane;
        const int cnt = c < ncols ? pw_column<false>(P, A, ni, B, nip1, c, nullptr, nullptr) : 0;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (c < P.jc_stride - 1) jc[c + 1] = base + incl;   // columns past ncols repeat the total
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        jc[0] = 0;
        P.nnz[f] = base;
    }
}

// exclusive scan of nnz[0..n) into offs[0..n], one CTA (n is at most a few million)
__global__ void __launch_bounds__(1024) k_pw_scan(const __grid_constant__ PwDev P) {
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int64_t per = (P.n + 1023) / 1024, lo = t * per, hi = lo + per < P.n ? lo + per : P.n;
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += P.nnz[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const int64_t v = part[i];
            part[i] = run;
            run += v;
        }
        P.offs[P.n] = run;
    }
    __syncthreads();
    int64_t run = part[t];
    for (int64_t i = lo; i < hi; ++i) {
        P.offs[i] = run;
        run += P.nnz[i];
    }
}

__global__ void __launch_bounds__(256) k_pw_fill(const __grid_constant__ PwDev P) {
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= P.n || f == 0) return;
    const int ni = P.ncand[(f - 1) * 2 + P.feat], nip1 = P.ncand[f * 2 + P.feat];
    const lm_cand *A = P.cand + ((f - 1) * 2 + P.feat) * P.cand_cap, *B = P.cand + (f * 2 + P.feat) * P.cand_cap;
    const int ncols = ni + P.nong;
    const int32_t *jc = P.jc + f * P.jc_stride;
    const int64_t o = P.offs[f];
    if (o + jc[ncols] > P.cap) return;  // the caller reports the overflow from offs[n]
    for (int c = lane; c < ncols; c += 32) pw_column<true>(P, A, ni, B, nip1, c, P.ir + o + jc[c], P.pr + o + jc[c]);
}

struct UnaryDev {
    const lm_cand *cand;
    const int32_t *ncand;
    int64_t n;
    int cand_cap, feat, np;
    double bb_w, bb_h, norm_fact;
    const lm_location_prior *pri;
    double *out;  // [n][np][cand_cap]
};

__global__ void __launch_bounds__(256) k_unary(const __grid_constant__ UnaryDev U) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= U.n * U.cand_cap) return;
    const int64_t f = idx / U.cand_cap;
    const int i = (int)(idx - f * U.cand_cap);
    const bool live = i < U.ncand[f * 2 + U.feat];
    lm_cand c{};
    double cx = 0, cy = 0;
    if (live) {
        c = U.cand[(f * 2 + U.feat) * U.cand_cap + i];
        cx = (double)c.x / U.bb_w;
        cy = (double)c.y / U.bb_h;
    }
    for (int j = 0; j < U.np; ++j) {
        double m = 0.0;
        if (live) {
            const lm_location_prior P = U.pri[j];
            if (P.area_x <= cx && cx < P.area_x + P.area_w && P.area_y <= cy && cy < P.area_y + P.area_h) {
                const double dx = cx - P.pos_x, dy = cy - P.pos_y;
                const double val = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) * U.norm_fact;
                if (val <= P.max_distance) m = (1 - val) * c.s;
            }
        }
        U.out[(f * U.np + j) * U.cand_cap + i] = m;
    }
}

}  // namespace

int lm_launch_unary(const lm_cand *cand, const int32_t *ncand, int64_t n, int cand_cap, int feat, int bb_w, int bb_h,
                    const lm_location_prior *pri, int np, double *out, cudaStream_t s) {
    UnaryDev U{cand, ncand, n, cand_cap, feat, np, (double)bb_w, (double)bb_h, 1 / sqrt(2.0), pri, out};
    const int64_t threads = n * cand_cap;
    if (threads == 0) return 0;
    k_unary<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(U);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int lm_launch_pairwise(const lm_cand *cand, const int32_t *ncand, int64_t n, int cand_cap, int feat, const lm_pairwise_params &p,
                       int32_t *jc, int64_t *nnz, int64_t *offs, int32_t *ir, double *pr, int64_t cap, int phase, cudaStream_t s) {
    PwDev P{};
    P.cand = cand;
    P.ncand = ncand;
    P.n = n;
    P.cand_cap = cand_cap;
    P.feat = feat;
    P.p = p;
    P.nong = p.ong_w * p.ong_h;
    P.jc_stride = cand_cap + P.nong + 1;
    P.occ = p.occluded_cost * p.alpha_vel;
    P.jc = jc;
    P.nnz = nnz;
    P.offs = offs;
    P.ir = ir;
    P.pr = pr;
    P.cap = cap;
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + 7) / 8);
    if (phase == 0) {
        k_pw_count<<<blocks, 256, 0, s>>>(P);
        k_pw_scan<<<1, 1024, 0, s>>>(P);
    } else {
        k_pw_fill<<<blocks, 256, 0, s>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? (phase == 0 ? 2 : 1) : -1;
}
