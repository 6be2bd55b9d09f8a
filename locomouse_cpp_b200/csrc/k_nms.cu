// k_nms.cu — candidate extraction from the positive-pixel lists: one CTA per (frame, feature).
//
//  k_nms_bottom = nmsMax          (LocoMouse_class.cpp:1610-1747) after detectPointCandidatesBottom's
//                                  tail masking (783, 849)
//  k_nms_side   = peakClustering  (LocoMouse_class.cpp:1749-1905), skipped when the feature's bottom
//                                  list is empty (820, 828)
//
// Common front end: drop tail-masked pixels (bottom only), build 64-bit keys
// (~score_bits << 32 | pixel index) and bitonic-sort them in shared memory: ascending key order ==
// score descending, row-major index ascending == the oracle's total order.
//
// nmsMax's chain suppression (discarded detections keep suppressing, Q3) is equivalent to
//   parent(j) = the best-ranked i < j whose box overlaps j's by more than 0.5, root = parent chain end
// (j is discarded by the FIRST overlapping i met in rank order, discarded or not, and inherits that
// i's root), so parents are found independently per detection and roots by pointer chasing.
// peakClustering is truly greedy: maxima are confirmed one at a time in rank order, each confirmation
// followed by a parallel sweep that clusters the still-free detections it overlaps.
// Overlap predicate inter/(2wh - inter) > 0.5  <=>  3*(w-|dx|)*(h-|dy|) > 2*w*h (exact in integers).
// Cluster sums are accumulated in double IN RANK ORDER by one thread per cluster, as the reference's
// loops do (1731-1739, 1865-1869), because double addition is not associative.
#include "lm_internal.h"

namespace {

// 16 bytes of shared memory per list entry: the 64-bit sort key, whose storage is reused after the sort for the score
// (low word) and the slot / root scratch (high word); link; packed position.  The small size class (<= NMS_SMALL_P entries,
// NMS_SMALL_T threads) keeps a CTA below 10 kB so that several fit beside a resident k_screen2 CTA.
constexpr int NMS_SMALL_T = 256, NMS_BIG_T = 512;

struct NmsSmem {
    unsigned long long *key;  // [P]  sort keys; afterwards {score, slot} pairs
    uint32_t *xy;             // [P]  x | y << 16 (one load per overlap test)
    int *link;                // [P]  parent / root / cluster id
    __device__ __forceinline__ float &s(int i) const { return reinterpret_cast<float *>(key)[2 * i]; }
    __device__ __forceinline__ int &slot(int i) const { return reinterpret_cast<int *>(key)[2 * i + 1]; }
};

// shared memory in front of the list arrays: root_rank[cand_cap], rounded to the alignment of the 64-bit keys
__host__ __device__ __forceinline__ size_t nms_cand_bytes(int cand_cap) { return ((size_t)cand_cap * 4 + 15) & ~(size_t)15; }

__device__ __forceinline__ NmsSmem carve(unsigned char *base, int P) {
    NmsSmem m;
    m.key = reinterpret_cast<unsigned long long *>(base);
    m.link = reinterpret_cast<int *>(m.key + P);
    m.xy = reinterpret_cast<uint32_t *>(m.link + P);
    return m;
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Loads, filters, sorts.  Returns the number of detections n (<= det_cap); sets *overflow.
template <int NMS_THREADS>
__device__ int load_and_sort(const LmBatch &b, int f, int feat, int view, NmsSmem &m, int P, int *s_n,
                             bool *overflow) {
    const int tid = threadIdx.x;
    const int list = (f * 2 + feat) * 2 + view;
    int cnt = b.det_count[list];
    *overflow = cnt > b.det_cap;
    if (cnt > b.det_cap) cnt = b.det_cap;
    const LmDet *d = b.det + (int64_t)list * b.det_cap;
    const int bw = b.bb_w;
    const uint8_t *tm = (view == LM_BOTTOM && b.tail_w > 0)
                            ? b.tailmask + (int64_t)f * b.bb_h[LM_BOTTOM] * b.tail_pitch
                            : nullptr;
    if (tid == 0) *s_n = 0;
    for (int i = tid; i < P; i += NMS_THREADS) m.key[i] = ~0ull;
    __syncthreads();
    for (int i = tid; i < cnt; i += NMS_THREADS) {
        LmDet e = d[i];
        int y = e.idx / bw, x = e.idx - y * bw;
        if (tm && x < b.tail_w && tm[y * b.tail_pitch + x]) continue;  // setTo(255, TAIL_MASK)
        int o = atomicAdd(s_n, 1);
        m.key[o] = ((unsigned long long)(~__float_as_uint(e.score)) << 32) | e.idx;
    }
    __syncthreads();
    const int n = *s_n;
    const int Q = next_pow2(n < 2 ? 2 : n);
    // bitonic network, one thread per compare-exchange PAIR (i, i | j): every thread of a pass is active
    for (int k = 2; k <= Q; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (Q >> 1); t += NMS_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const unsigned long long a = m.key[i], c = m.key[l];
                if ((a > c) == ((i & k) == 0)) {
                    m.key[i] = c;
                    m.key[l] = a;
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += NMS_THREADS) {
        unsigned long long k = m.key[i];
        unsigned idx = (unsigned)(k & 0xffffffffu);
        int y = idx / bw;
        m.xy[i] = (uint32_t)(idx - y * bw) | ((uint32_t)y << 16);
        m.s(i) = __uint_as_float(~(unsigned)(k >> 32));  // overwrites this entry's own key
    }
    __syncthreads();
    return n;
}

__device__ __forceinline__ bool overlaps(int dx, int dy, int w, int h, int wh2) {
    dx = dx < 0 ? -dx : dx;
    dy = dy < 0 ? -dy : dy;
    return dx < w && dy < h && 3 * (w - dx) * (h - dy) > wh2;
}

// One WARP per cluster slot: the lanes scan the rank-ordered detections 32 at a time (ballot of the slot's members),
// lane 0 accumulates the members of each ballot in rank order in double, as the reference's loops do (1731-1739,
// 1865-1869; double addition is not associative, so the order is part of the result).  On entry link[i] holds the
// root of detection i and slot(r) the candidate slot of root r; link is first replaced by the root's candidate slot.
template <bool HALF_EVEN, int NMS_THREADS>
__device__ void write_candidates(const NmsSmem &m, int n, int ncand, int cap, lm_cand *out, const int *root_rank) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = NMS_THREADS / 32;
    for (int i = tid; i < n; i += NMS_THREADS) m.link[i] = m.slot(m.link[i]);
    __syncthreads();
    for (int k = warp; k < cap; k += NWARPS) {
        lm_cand c;
        c.x = -1;
        c.y = -1;
        c.s = -1.0;
        if (k < ncand) {
            double wx = 0.0, wy = 0.0, ss = 0.0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                unsigned mask = __ballot_sync(0xffffffffu, i < n && m.link[i] == k);
                if (lane == 0)
                    while (mask) {
                        const int q = base + __ffs(mask) - 1;
                        mask &= mask - 1;
                        const double s = (double)m.s(q);
                        const uint32_t pq = m.xy[q];
                        wx = __dadd_rn(wx, __dmul_rn((double)(int)(pq & 0xffffu), s));
                        wy = __dadd_rn(wy, __dmul_rn((double)(int)(pq >> 16), s));
                        ss = __dadd_rn(ss, s);
                    }
            }
            if (lane == 0) {
                const double qx = __ddiv_rn(wx, ss), qy = __ddiv_rn(wy, ss);
                if (HALF_EVEN) {
                    c.x = __double2int_rn(qx);
                    c.y = __double2int_rn(qy);
                } else {
                    c.x = (int)round(qx);
                    c.y = (int)round(qy);
                }
                c.s = (double)m.s(root_rank[k]);
            }
        }
        if (lane == 0) out[k] = c;
    }
}

// ---- nmsMax ------------------------------------------------------------------------------------------
template <int NMS_THREADS>
__global__ void __launch_bounds__(NMS_THREADS) k_nms_bottom(const __grid_constant__ LmBatch b, int P, int lo, int hi, unsigned char *scratch) {
    extern __shared__ __align__(16) unsigned char raw[];
    // lists of the small class live in shared memory; the rare longer ones in this CTA's slice of a global scratch array
    // (16 P bytes per list), so that no launch of the stage asks for more shared memory than fits beside a k_screen2 CTA
    NmsSmem m = carve(scratch ? scratch + (size_t)blockIdx.x * (size_t)P * 16 : raw + nms_cand_bytes(b.cand_cap), P);
    int *root_rank = reinterpret_cast<int *>(raw);  // [cand_cap]
    __shared__ int s_n, s_nc;
    const int f = blockIdx.x >> 1, feat = blockIdx.x & 1, tid = threadIdx.x;
    {   // size class of this list (the two launches partition the lists by raw count)
        int c = b.det_count[(f * 2 + feat) * 2 + LM_BOTTOM];
        c = c > b.det_cap ? b.det_cap : c;
        if (c <= lo || c > hi) return;
    }
    bool overflow;
    const int n = load_and_sort<NMS_THREADS>(b, f, feat, LM_BOTTOM, m, P, &s_n, &overflow);
    const LmTemplateDev &T = b.tmpl[LM_BOTTOM][feat];
    const int w = T.kw, h = T.kh, wh2 = 2 * w * h;
    if (tid == 0) s_nc = 0;
    // parent = first better-ranked overlapping detection.  One warp per detection j: the lanes test 32 better ranks at
    // a time and the first set ballot bit is the parent, so the cost is ceil(parent_rank / 32) and evenly spread.
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int j = warp; j < n; j += NMS_THREADS / 32) {
            const uint32_t pj = m.xy[j];
            const int xj = (int)(pj & 0xffffu), yj = (int)(pj >> 16);
            int par = j;
            for (int base = 0; base < j; base += 32) {
                const int i = base + lane;
                const uint32_t pi = m.xy[i < j ? i : j];
                const bool hit = i < j && overlaps((int)(pi & 0xffffu) - xj, (int)(pi >> 16) - yj, w, h, wh2);
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (bal) {
                    par = base + __ffs(bal) - 1;
                    break;
                }
            }
            if (lane == 0) m.link[j] = par;
        }
    }
    __syncthreads();
    // roots by pointer chasing (parents only point to better ranks, so chains end)
    for (int j = tid; j < n; j += NMS_THREADS) {
        int r = j;
        while (m.link[r] != r) r = m.link[r];
        m.slot(j) = r;  // temp: root of j
    }
    __syncthreads();
    for (int j = tid; j < n; j += NMS_THREADS) m.link[j] = m.slot(j);
    __syncthreads();
    // candidate slots = roots in rank order (serial prefix by one warp is enough: n is small)
    if (tid < 32) {
        int base = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            int j = j0 + tid;
            bool is_root = j < n && m.link[j] == j;
            unsigned bal = __ballot_sync(0xffffffffu, is_root);
            if (is_root) {
                int k = base + __popc(bal & ((1u << tid) - 1));
                m.slot(j) = k;
                if (k < b.cand_cap) root_rank[k] = j;
            }
            base += __popc(bal);
        }
        if (tid == 0) s_nc = base;
    }
    __syncthreads();
    const int nc = s_nc;
    const int ncw = nc < b.cand_cap ? nc : b.cand_cap;
    write_candidates<true, NMS_THREADS>(m, n, ncw, b.cand_cap, b.bottom + (int64_t)(f * 2 + feat) * b.cand_cap, root_rank);
    if (tid == 0) {
        b.n_bottom[f * 2 + feat] = ncw;
        unsigned fl = 0;
        if (overflow) fl |= LM_FLAG_DET_OVERFLOW;
        if (nc > b.cand_cap) fl |= LM_FLAG_CAND_OVERFLOW;
        if (fl) atomicOr(&b.flags[f], fl);
    }
}

// ---- peakClustering ----------------------------------------------------------------------------------
template <int NMS_THREADS>
__global__ void __launch_bounds__(NMS_THREADS) k_nms_side(const __grid_constant__ LmBatch b, int P, int lo, int hi, unsigned char *scratch) {
    extern __shared__ __align__(16) unsigned char raw[];
    // lists of the small class live in shared memory; the rare longer ones in this CTA's slice of a global scratch array
    // (16 P bytes per list), so that no launch of the stage asks for more shared memory than fits beside a k_screen2 CTA
    NmsSmem m = carve(scratch ? scratch + (size_t)blockIdx.x * (size_t)P * 16 : raw + nms_cand_bytes(b.cand_cap), P);
    int *root_rank = reinterpret_cast<int *>(raw);
    __shared__ int s_n, s_next;
    const int f = blockIdx.x >> 1, feat = blockIdx.x & 1, tid = threadIdx.x;
    lm_cand *out = b.side + (int64_t)(f * 2 + feat) * b.cand_cap;
    {
        int c = b.det_count[(f * 2 + feat) * 2 + LM_SIDE];
        c = c > b.det_cap ? b.det_cap : c;
        if (c <= lo || c > hi) return;
    }
    if (b.n_bottom[f * 2 + feat] == 0) {  // Q6
        for (int k = tid; k < b.cand_cap; k += NMS_THREADS) {
            lm_cand c;
            c.x = -1;
            c.y = -1;
            c.s = -1.0;
            out[k] = c;
        }
        if (tid == 0) b.n_side[f * 2 + feat] = 0;
        return;
    }
    bool overflow;
    const int n = load_and_sort<NMS_THREADS>(b, f, feat, LM_SIDE, m, P, &s_n, &overflow);
    const LmTemplateDev &T = b.tmpl[LM_SIDE][feat];
    const int w = T.kw, h = T.kh, wh2 = 2 * w * h;
    for (int j = tid; j < n; j += NMS_THREADS) m.link[j] = -1;  // -1 = free
    __syncthreads();
    // Greedy: the next maximum is the best-ranked detection that is still free; it is found with a block-wide
    // atomicMin over 512 candidates at a time instead of a serial scan of the (mostly clustered) list.
    const int INF = 0x7fffffff;
    if (tid == 0) s_next = INF;
    __syncthreads();
    int nc = 0, c = -1;
    for (;;) {
        int found = INF;
        for (int base = c + 1; base < n; base += NMS_THREADS) {
            const int idx = base + tid;
            if (idx < n && m.link[idx] < 0) atomicMin(&s_next, idx);
            __syncthreads();
            found = s_next;
            __syncthreads();  // everyone has read s_next before the next chunk (or the reset below) touches it
            if (found != INF) break;
        }
        if (found == INF) break;
        c = found;
        const int xc = (int)(m.xy[c] & 0xffffu), yc = (int)(m.xy[c] >> 16);
        if (tid == 0) {
            s_next = INF;
            m.link[c] = c;
            m.slot(c) = nc;
            if (nc < b.cand_cap) root_rank[nc] = c;
        }
        for (int j = c + 1 + tid; j < n; j += NMS_THREADS)
            if (m.link[j] < 0) {
                const uint32_t pq = m.xy[j];
                if (overlaps((int)(pq & 0xffffu) - xc, (int)(pq >> 16) - yc, w, h, wh2)) m.link[j] = c;
            }
        ++nc;
        __syncthreads();
    }
    const int ncw = nc < b.cand_cap ? nc : b.cand_cap;
    write_candidates<false, NMS_THREADS>(m, n, ncw, b.cand_cap, out, root_rank);
    if (tid == 0) {
        b.n_side[f * 2 + feat] = ncw;
        unsigned fl = 0;
        if (overflow) fl |= LM_FLAG_DET_OVERFLOW;
        if (nc > b.cand_cap) fl |= LM_FLAG_CAND_OVERFLOW;
        if (fl) atomicOr(&b.flags[f], fl);
    }
}


}  // namespace

int lm_launch_nms(const LmBatch &b, cudaStream_t s) {
    int P = 2;
    while (P < b.det_cap) P <<= 1;
    // Size classes: lists of up to SMALL positives are sorted and clustered in shared memory (16 B per entry: a CTA fits
    // beside a resident k_screen2 CTA of a neighbouring sub-batch); the rare longer lists use the same code on a slice of
    // a global scratch array (the other CTAs of that launch exit at once), so that no launch waits for screen-free SMs.
    int SMALL = 1024;
    if (const char *e = getenv("LM_NMS_SMALL")) SMALL = std::max(2, atoi(e));
    static LmDevOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_nms_bottom<NMS_BIG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_nms_side<NMS_BIG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        lm_prefer_max_shared(k_nms_bottom<NMS_BIG_T>);
        lm_prefer_max_shared(k_nms_side<NMS_BIG_T>);
        lm_prefer_max_shared(k_nms_bottom<NMS_SMALL_T>);
        lm_prefer_max_shared(k_nms_side<NMS_SMALL_T>);
    }
    int launches = 0;
    int cls[4], ncls = 0;
    for (int c : {SMALL, P})
        if (ncls == 0 || (c > cls[ncls - 1] && cls[ncls - 1] < P)) cls[ncls++] = std::min(c, P);
    for (int view = 0; view < 2; ++view) {
        int lo = -1;
        for (int q = 0; q < ncls; ++q) {
            const bool small = q == 0 && cls[q] <= 1024;   // few entries: half the threads, so that more CTAs are resident
            unsigned char *scratch = q == 0 ? nullptr : b.nms_scratch;   // the views' launches follow each other in the stream
            const size_t smem = nms_cand_bytes(b.cand_cap) + (q == 0 ? (size_t)cls[q] * 16 : 0);
            if (q > 0 && !scratch) return -1;
            if (view == 0) {
                if (small)
                    k_nms_bottom<NMS_SMALL_T><<<b.B * 2, NMS_SMALL_T, smem, s>>>(b, cls[q], lo, cls[q], scratch);
                else
                    k_nms_bottom<NMS_BIG_T><<<b.B * 2, NMS_BIG_T, smem, s>>>(b, cls[q], lo, cls[q], scratch);
            } else {
                if (small)
                    k_nms_side<NMS_SMALL_T><<<b.B * 2, NMS_SMALL_T, smem, s>>>(b, cls[q], lo, cls[q], scratch);
                else
                    k_nms_side<NMS_BIG_T><<<b.B * 2, NMS_BIG_T, smem, s>>>(b, cls[q], lo, cls[q], scratch);
            }
            ++launches;
            lo = cls[q];
        }
    }
    return launches;
}
