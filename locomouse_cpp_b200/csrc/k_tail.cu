// k_tail.cu — tail detection after the two tail correlations: one CTA per frame.
//
// Replaces LocoMouse::detectLineCandidates from the binarisation on (LocoMouse_class.cpp:2593-2742)
// and selectLargestRegion (2744-2767):
//   bottom (score > 0) map -> largest connected region (= TAIL_MASK, also consumed by the bottom
//   NMS kernel) -> column-max masks the side map -> largest side region -> the bottom extent
//   [first,last) is split into n_tail_points segments -> per-segment centroids (cv::moments of a
//   binary image = pixel counts and coordinate sums, (int) truncation) -> TRACKS_TAIL (x, y, z).
// Connected components: lock-free union-find on pixel indices (atomicMin hooks, min index = root),
// labels in an L2-resident scratch; "largest" = max area, ties -> the component OpenCV labels first
// (first 2x2 block in block-raster order for 8-connectivity, first pixel for 4-connectivity), see
// oracle/lm_oracle.cpp largest_region and tests/test_oracle_vs_cv2.py.
#include "lm_internal.h"

namespace {

constexpr int TAIL_THREADS = 512;

__device__ __forceinline__ int uf_find(volatile int *L, int p) {
    for (;;) {
        int q = L[p];
        if (q == p) return p;
        p = q;
    }
}

__device__ __forceinline__ void uf_union(int *L, int a, int b) {
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            int t = a;
            a = b;
            b = t;
        }
        int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// Largest connected component of the foreground {p : bin[p] != 0 && (colgate == null || colgate[x])}.
// Writes mask[p] (0/1) for every pixel, accumulates per-column pixel count / row sum of the winner
// in shared memory, returns nothing; *s_best (shared) holds the packed winner or 0.
__device__ void cc_largest(const uint8_t *bin, const int *colgate, int rows, int cols, int pitch, int conn,
                           int *L, int *area, int *key, uint8_t *mask, int *colcnt, int *colsum, int *colany,
                           unsigned long long *s_best) {
    const int tid = threadIdx.x;
    const int n = rows * cols;
    const int bcols = (cols + 1) >> 1;
    auto fg = [&](int r, int c) -> bool {
        return bin[r * pitch + c] != 0 && (colgate == nullptr || colgate[c] != 0);
    };
    if (tid == 0) *s_best = 0ull;
    for (int c = tid; c < cols; c += TAIL_THREADS) {
        colcnt[c] = 0;
        colsum[c] = 0;
        if (colany) colany[c] = 0;
    }
    for (int p = tid; p < n; p += TAIL_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (fg(r, c)) {
            L[p] = p;
            area[p] = 0;
            key[p] = 0x7fffffff;
        }
    }
    __syncthreads();
    for (int p = tid; p < n; p += TAIL_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (!fg(r, c)) continue;
        if (c > 0 && fg(r, c - 1)) uf_union(L, p, p - 1);
        if (r > 0) {
            if (fg(r - 1, c)) uf_union(L, p, p - cols);
            if (conn == 8) {
                if (c > 0 && fg(r - 1, c - 1)) uf_union(L, p, p - cols - 1);
                if (c + 1 < cols && fg(r - 1, c + 1)) uf_union(L, p, p - cols + 1);
            }
        }
    }
    __syncthreads();
    for (int p = tid; p < n; p += TAIL_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (!fg(r, c)) continue;
        int root = uf_find(L, p);
        L[p] = root;
        atomicAdd(&area[root], 1);
        int k = (conn == 8) ? (r >> 1) * bcols + (c >> 1) : p;
        atomicMin(&key[root], k);
    }
    __syncthreads();
    for (int p = tid; p < n; p += TAIL_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (!fg(r, c)) continue;
        if (((volatile int *)L)[p] == p) {
            // area (21 bits) | inverted key (21 bits) | pixel index + 1 (22 bits)
            unsigned long long v = ((unsigned long long)((volatile int *)area)[p] << 43) |
                                   ((unsigned long long)(0x1fffff - ((volatile int *)key)[p]) << 22) |
                                   (unsigned long long)(p + 1);
            atomicMax(s_best, v);
        }
    }
    __syncthreads();
    const unsigned long long best = *s_best;
    const int best_root = (int)(best & 0x3fffff) - 1;
    for (int p = tid; p < n; p += TAIL_THREADS) {
        int r = p / cols, c = p - r * cols;
        uint8_t m = 0;
        if (best_root >= 0 && fg(r, c) && ((volatile int *)L)[p] == best_root) {
            m = 1;
            atomicAdd(&colcnt[c], 1);
            atomicAdd(&colsum[c], r);
            if (colany) colany[c] = 1;
        }
        mask[r * pitch + c] = m;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const __grid_constant__ LmBatch b) {
    extern __shared__ int sm[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int cols = b.tail_w, pitch = b.tail_pitch, np = b.n_tail_points;
    int *colany = sm, *cnt_b = sm + cols, *sum_b = sm + 2 * cols, *cnt_s = sm + 3 * cols, *sum_s = sm + 4 * cols;
    int *tx = sm + 5 * cols;  // [np]
    __shared__ unsigned long long s_best;
    __shared__ int s_first, s_last;

    int *L = b.cc + (int64_t)f * 3 * b.cc_stride, *area = L + b.cc_stride, *key = area + b.cc_stride;
    const int hb = b.bb_h[LM_BOTTOM], hs = b.bb_h[LM_SIDE];
    const uint8_t *bin_b = b.tailbin[LM_BOTTOM] + (int64_t)f * hb * pitch;
    const uint8_t *bin_s = b.tailbin[LM_SIDE] + (int64_t)f * hs * pitch;
    uint8_t *mask_b = b.tailmask + (int64_t)f * hb * pitch;
    uint8_t *mask_s = b.sidemask + (int64_t)f * hs * pitch;
    int32_t *tr = b.tail + (int64_t)f * 3 * np;

    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
    }
    cc_largest(bin_b, nullptr, hb, cols, pitch, b.conn, L, area, key, mask_b, cnt_b, sum_b, colany, &s_best);
    for (int c = tid; c < cols; c += TAIL_THREADS)
        if (colany[c]) {
            atomicMin(&s_first, c);
            atomicMax(&s_last, c);
        }
    cc_largest(bin_s, colany, hs, cols, pitch, b.conn, L, area, key, mask_s, cnt_s, sum_s, nullptr, &s_best);
    // (cc_largest ends with __syncthreads, so s_first / s_last are final here)
    for (int i = tid; i < 3 * np; i += TAIL_THREADS) tr[i] = -1;
    __syncthreads();
    const int first = s_first, last = s_last;
    if (last < 0) return;  // no tail region: all -1 (class.cpp:2654-2661)
    const int width = last - first, rem = width % np, reg = (width - rem) / np;
    if (tid < np) {
        const int i = tid;
        const int segw = reg + (i < rem ? 1 : 0);
        const int segx = first + i * reg + (i < rem ? i : rem);
        long long m00 = 0, m10 = 0, m01 = 0;
        for (int c = 0; c < segw; ++c) {
            m00 += cnt_b[segx + c];
            m10 += (long long)c * cnt_b[segx + c];
            m01 += sum_b[segx + c];
        }
        int x = -1;
        if (m00 > 0) {
            x = (int)(m10 / m00) + segx;
            tr[i] = x;
            tr[np + i] = (int)(m01 / m00);
        }
        if (x > 0 && cnt_s[x] > 0) tr[2 * np + i] = sum_s[x] / cnt_s[x];
    }
    (void)tx;
}

}  // namespace

int lm_launch_tail(const LmBatch &b, cudaStream_t s) {
    if (b.tail_w <= 0) return 0;
    size_t smem = (size_t)(5 * b.tail_w + b.n_tail_points) * sizeof(int);
    k_tail<<<b.B, TAIL_THREADS, smem, s>>>(b);
    return 1;
}
