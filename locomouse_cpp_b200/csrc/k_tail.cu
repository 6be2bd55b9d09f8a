// k_tail.cu — tail detection after the two tail correlations: one CTA per frame.
//
// Replaces LocoMouse::detectLineCandidates from the binarisation on (LocoMouse_class.cpp:2593-2742)
// and selectLargestRegion (2744-2767):
//   bottom (score > 0) map -> largest connected region (= TAIL_MASK, also consumed by the bottom
//   NMS kernel) -> column-max masks the side map -> largest side region -> the bottom extent
//   [first,last) is split into n_tail_points segments -> per-segment centroids (cv::moments of a
//   binary image = pixel counts and coordinate sums, (int) truncation) -> TRACKS_TAIL (x, y, z).
//
// Connected components are labelled on RUNS, entirely in shared memory:
//   1. the u8 map is read once with 16-byte loads and packed to a row-major bit image;
//   2. run starts are  bits & ~(bits << 1 | carry) ; a block scan numbers the runs row-major, so the
//      runs of a row are contiguous and x-sorted;
//   3. every run is united with the runs of the previous row it touches (8-connectivity: [x0-1,x1+1],
//      4-connectivity: [x0,x1]) by lock-free union-find (atomicMin hooks, smaller id = root);
//   4. per root: area = sum of run lengths, OpenCV label order key = min over runs of the first 2x2 block
//      ((r>>1)*ceil(W/2) + (x0>>1)) for 8-connectivity / first pixel (r*W + x0) for 4-connectivity;
//      "largest" = max area, ties -> smallest key (oracle/lm_oracle.cpp largest_region, pinned vs cv2).
// Frames whose maps have more than RUNCAP runs take the pixel-based global-memory path (k_tail_slow).
#include <cstdlib>

#include "lm_internal.h"

namespace {

constexpr int TAIL_THREADS = 256;
constexpr int RUNCAP = 3072;               // largest run capacity the kernel's 16-bit run indices are used with
constexpr int TAIL_RUNCAP_DEFAULT = 1536;

struct TailSmem {
    uint32_t *bits;      // [rows][wpr] input bit image (row-major words)
    uint32_t *obits;     // [rows][wpr] winner mask
    int *rowfirst;       // [rows + 1] first run id of each row
    unsigned short *rrow, *rx0, *rx1;  // [RUNCAP]
    int *parent;         // [RUNCAP]
    int *area, *key;     // [RUNCAP]
    int *colany, *cnt, *sum;  // [cols] each (colany persists from bottom to side)
};

__device__ __forceinline__ int uf_find_s(volatile int *L, int p) {
    for (;;) {
        int q = L[p];
        if (q == p) return p;
        p = q;
    }
}

__device__ __forceinline__ void uf_union_s(int *L, int a, int b) {
    for (;;) {
        a = uf_find_s(L, a);
        b = uf_find_s(L, b);
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

// block-wide exclusive scan of one int per thread (TAIL_THREADS threads); returns exclusive prefix,
// *total gets the sum.  scratch: >= 32 ints.
__device__ int block_exscan(int v, int *scratch, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) scratch[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < (TAIL_THREADS / 32) ? scratch[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += t;
        }
        scratch[lane] = si - s;  // exclusive warp offsets
        if (lane == 31) scratch[32] = si;
    }
    __syncthreads();
    int ex = scratch[w] + incl - v;
    *total = scratch[32];
    __syncthreads();
    return ex;
}

// Largest region of the bit image in S.bits (rows x cols).  On return S.obits holds the winner mask,
// S.cnt / S.sum its per-column pixel count and row sum (and S.colany if want_colany).  Returns false when
// the run capacity is exceeded (caller falls back to the slow path).
__device__ bool largest_region_runs(TailSmem &S, int rows, int cols, int wpr, int conn, bool want_colany,
                                    int *scratch, unsigned long long *s_best, int runcap) {
    const int tid = threadIdx.x;
    const int nwords = rows * wpr;
    // ---- count run starts per word, scan -----------------------------------------------------------
    // each thread owns a contiguous chunk of words so that run ids are row-major
    const int chunk = (nwords + TAIL_THREADS - 1) / TAIL_THREADS;
    const int w0 = tid * chunk, w1 = min(nwords, w0 + chunk);
    int mine = 0;
    for (int wi = w0; wi < w1; ++wi) {
        const int c = wi % wpr;
        const uint32_t b = S.bits[wi];
        const uint32_t carry = (c > 0) ? (S.bits[wi - 1] >> 31) : 0u;
        mine += __popc(b & ~((b << 1) | carry));
    }
    int total;
    int base = block_exscan(mine, scratch, &total);
    if (total > runcap) return false;
    for (int i = tid; i <= rows; i += TAIL_THREADS) S.rowfirst[i] = total;  // default: end
    for (int c = tid; c < cols; c += TAIL_THREADS) {
        S.cnt[c] = 0;
        S.sum[c] = 0;
        if (want_colany) S.colany[c] = 0;
    }
    for (int i = tid; i < nwords; i += TAIL_THREADS) S.obits[i] = 0u;
    if (tid == 0) *s_best = 0ull;
    __syncthreads();
    // ---- emit runs ------------------------------------------------------------------------------------
    for (int wi = w0; wi < w1; ++wi) {
        const int r = wi / wpr, c = wi - r * wpr;
        const uint32_t b = S.bits[wi];
        const uint32_t carry = (c > 0) ? (S.bits[wi - 1] >> 31) : 0u;
        uint32_t starts = b & ~((b << 1) | carry);
        while (starts) {
            const int bit = __ffs(starts) - 1;
            starts &= starts - 1;
            const int x0 = c * 32 + bit;
            // run end: first zero bit at or after x0 (may continue into following words of the row)
            int x1;
            {
                int cw = c;
                uint32_t inv = ~S.bits[wi] & (0xffffffffu << bit);
                while (inv == 0u && cw + 1 < wpr) {
                    ++cw;
                    inv = ~S.bits[r * wpr + cw];
                }
                x1 = (inv ? cw * 32 + __ffs(inv) - 1 : wpr * 32) - 1;
                if (x1 >= cols) x1 = cols - 1;
            }
            const int id = base++;
            S.rrow[id] = (unsigned short)r;
            S.rx0[id] = (unsigned short)x0;
            S.rx1[id] = (unsigned short)x1;
            S.parent[id] = id;
            S.area[id] = 0;
            S.key[id] = 0x7fffffff;
            atomicMin(&S.rowfirst[r], id);
        }
    }
    __syncthreads();
    // rows without runs: rowfirst[r] = rowfirst of the next row that has one (suffix min)
    if (tid == 0) {
        int nxt = total;
        for (int r = rows; r >= 0; --r) {
            if (S.rowfirst[r] > nxt) S.rowfirst[r] = nxt;
            nxt = S.rowfirst[r];
        }
    }
    __syncthreads();
    // ---- unite with the previous row ---------------------------------------------------------------
    const int ext = (conn == 8) ? 1 : 0;
    for (int id = tid; id < total; id += TAIL_THREADS) {
        const int r = S.rrow[id];
        if (r == 0) continue;
        const int lo = (int)S.rx0[id] - ext, hi = (int)S.rx1[id] + ext;
        for (int q = S.rowfirst[r - 1]; q < S.rowfirst[r]; ++q) {
            if ((int)S.rx1[q] < lo) continue;
            if ((int)S.rx0[q] > hi) break;
            uf_union_s(S.parent, id, q);
        }
    }
    __syncthreads();
    const int bcols = (cols + 1) >> 1;
    for (int id = tid; id < total; id += TAIL_THREADS) {
        const int root = uf_find_s(S.parent, id);
        S.parent[id] = root;
        const int r = S.rrow[id], x0 = S.rx0[id];
        atomicAdd(&S.area[root], (int)S.rx1[id] - x0 + 1);
        atomicMin(&S.key[root], (conn == 8) ? (r >> 1) * bcols + (x0 >> 1) : r * cols + x0);
    }
    __syncthreads();
    for (int id = tid; id < total; id += TAIL_THREADS)
        if (((volatile int *)S.parent)[id] == id) {
            unsigned long long v = ((unsigned long long)S.area[id] << 43) |
                                   ((unsigned long long)(0x1fffff - S.key[id]) << 22) | (unsigned long long)(id + 1);
            atomicMax(s_best, v);
        }
    __syncthreads();
    const int best = (int)(*s_best & 0x3fffff) - 1;
    for (int id = tid; id < total; id += TAIL_THREADS) {
        if (best < 0 || ((volatile int *)S.parent)[id] != best) continue;
        const int r = S.rrow[id], x0 = S.rx0[id], x1 = S.rx1[id];
        for (int x = x0; x <= x1; ++x) {
            atomicAdd(&S.cnt[x], 1);
            atomicAdd(&S.sum[x], r);
            if (want_colany) S.colany[x] = 1;
        }
        for (int cw = x0 >> 5; cw <= (x1 >> 5); ++cw) {
            const int a = max(x0, cw * 32) - cw * 32, e = min(x1, cw * 32 + 31) - cw * 32;
            const uint32_t m = (e == 31 ? 0xffffffffu : ((1u << (e + 1)) - 1u)) & (0xffffffffu << a);
            atomicOr(&S.obits[r * wpr + cw], m);
        }
    }
    __syncthreads();
    return true;
}

// u8 map [rows][pitch] (global) -> bit image in shared memory, optionally gated by colany
__device__ void load_bits(const uint8_t *bin, int rows, int cols, int pitch, int wpr, const int *colgate,
                          uint32_t *bits) {
    const int tid = threadIdx.x;
    const int nwords = rows * wpr;
    for (int wi = tid; wi < nwords; wi += TAIL_THREADS) {
        const int r = wi / wpr, c = wi - r * wpr;
        const int x0 = c * 32;
        const uint8_t *src = bin + (int64_t)r * pitch + x0;
        uint32_t w = 0;
        if (x0 + 32 <= pitch && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                // 4 bytes (each 0/1) -> 4 bits
                uint32_t t = v[q] & 0x01010101u;
                t = (t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xfu;
                w |= t << (4 * q);
            }
        } else {
            for (int q = 0; q < 32 && x0 + q < cols; ++q) w |= (uint32_t)(src[q] != 0) << q;
        }
        // clear bits beyond the box and apply the column gate
        const int valid = min(32, cols - x0);
        if (valid < 32) w &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
        if (colgate) {
            uint32_t g = 0;
            for (int q = 0; q < valid; ++q) g |= (uint32_t)(colgate[x0 + q] != 0) << q;
            w &= g;
        }
        bits[wi] = w;
    }
}

// winner mask bits (shared) -> u8 0/1 map (global)
__device__ void store_mask(const uint32_t *obits, int rows, int cols, int pitch, int wpr, uint8_t *mask) {
    const int nq = rows * (pitch >> 2);  // 4-pixel groups
    const int qpr = pitch >> 2;
    for (int i = threadIdx.x; i < nq; i += TAIL_THREADS) {
        const int r = i / qpr, x = (i - r * qpr) << 2;
        uint32_t out = 0;
        if (x < cols) {
            const uint32_t w = obits[r * wpr + (x >> 5)] >> (x & 31);
            out = (w & 1u) | ((w & 2u) << 7) | ((w & 4u) << 14) | ((w & 8u) << 21);
        }
        reinterpret_cast<uint32_t *>(mask + (int64_t)r * pitch)[x >> 2] = out;
    }
}

__device__ void write_tracks(const LmBatch &b, int f, int first, int last, const int *cnt_b, const int *sum_b,
                             const int *cnt_s, const int *sum_s) {
    const int np = b.n_tail_points, tid = threadIdx.x;
    int32_t *tr = b.tail + (int64_t)f * 3 * np;
    for (int i = tid; i < 3 * np; i += blockDim.x) tr[i] = -1;
    __syncthreads();
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
}

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const __grid_constant__ LmBatch b, int *need_slow, int runcap) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int cols = b.tail_w, pitch = b.tail_pitch;
    const int hb = b.bb_h[LM_BOTTOM], hs = b.bb_h[LM_SIDE], hmax = max(hb, hs);
    const int wpr = (cols + 31) >> 5;
    TailSmem S;
    unsigned char *p = raw;
    S.bits = reinterpret_cast<uint32_t *>(p); p += (size_t)hmax * wpr * 4;
    S.obits = reinterpret_cast<uint32_t *>(p); p += (size_t)hmax * wpr * 4;
    S.rowfirst = reinterpret_cast<int *>(p); p += (size_t)(hmax + 1) * 4;
    const size_t rcap = (size_t)((runcap > 1 ? runcap : 1) + 63) & ~(size_t)63;  // run arrays are sized for the launch's threshold
    S.parent = reinterpret_cast<int *>(p); p += rcap * 4;
    S.area = reinterpret_cast<int *>(p); p += rcap * 4;
    S.key = reinterpret_cast<int *>(p); p += rcap * 4;
    S.colany = reinterpret_cast<int *>(p); p += (size_t)cols * 4;
    int *cnt_b = reinterpret_cast<int *>(p); p += (size_t)cols * 4;
    int *sum_b = reinterpret_cast<int *>(p); p += (size_t)cols * 4;
    int *cnt_s = reinterpret_cast<int *>(p); p += (size_t)cols * 4;
    int *sum_s = reinterpret_cast<int *>(p); p += (size_t)cols * 4;
    S.rrow = reinterpret_cast<unsigned short *>(p); p += rcap * 2;
    S.rx0 = reinterpret_cast<unsigned short *>(p); p += rcap * 2;
    S.rx1 = reinterpret_cast<unsigned short *>(p); p += rcap * 2;
    __shared__ int scratch[40];
    __shared__ unsigned long long s_best;
    __shared__ int s_first, s_last;

    const uint8_t *bin_b = b.tailbin[LM_BOTTOM] + (int64_t)f * hb * pitch;
    const uint8_t *bin_s = b.tailbin[LM_SIDE] + (int64_t)f * hs * pitch;
    uint8_t *mask_b = b.tailmask + (int64_t)f * hb * pitch;

    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
        need_slow[f] = 0;
    }
    load_bits(bin_b, hb, cols, pitch, wpr, nullptr, S.bits);
    __syncthreads();
    S.cnt = cnt_b;
    S.sum = sum_b;
    if (!largest_region_runs(S, hb, cols, wpr, b.conn, true, scratch, &s_best, runcap)) {
        if (tid == 0) need_slow[f] = 1;
        return;
    }
    store_mask(S.obits, hb, cols, pitch, wpr, mask_b);
    for (int c = tid; c < cols; c += TAIL_THREADS)
        if (S.colany[c]) {
            atomicMin(&s_first, c);
            atomicMax(&s_last, c);
        }
    load_bits(bin_s, hs, cols, pitch, wpr, S.colany, S.bits);
    __syncthreads();
    S.cnt = cnt_s;
    S.sum = sum_s;
    if (!largest_region_runs(S, hs, cols, wpr, b.conn, false, scratch, &s_best, runcap)) {
        if (tid == 0) need_slow[f] = 1;
        return;
    }
    write_tracks(b, f, s_first, s_last, cnt_b, sum_b, cnt_s, sum_s);
}

// ================= slow path: pixel union-find in global memory (any number of runs) =================
constexpr int SLOW_THREADS = 512;

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
    for (int c = tid; c < cols; c += SLOW_THREADS) {
        colcnt[c] = 0;
        colsum[c] = 0;
        if (colany) colany[c] = 0;
    }
    for (int p = tid; p < n; p += SLOW_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (fg(r, c)) {
            L[p] = p;
            area[p] = 0;
            key[p] = 0x7fffffff;
        }
    }
    __syncthreads();
    for (int p = tid; p < n; p += SLOW_THREADS) {
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
    for (int p = tid; p < n; p += SLOW_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (!fg(r, c)) continue;
        int root = uf_find(L, p);
        L[p] = root;
        atomicAdd(&area[root], 1);
        int k = (conn == 8) ? (r >> 1) * bcols + (c >> 1) : p;
        atomicMin(&key[root], k);
    }
    __syncthreads();
    for (int p = tid; p < n; p += SLOW_THREADS) {
        int r = p / cols, c = p - r * cols;
        if (!fg(r, c)) continue;
        if (((volatile int *)L)[p] == p) {
            unsigned long long v = ((unsigned long long)((volatile int *)area)[p] << 43) |
                                   ((unsigned long long)(0x1fffff - ((volatile int *)key)[p]) << 22) |
                                   (unsigned long long)(p + 1);
            atomicMax(s_best, v);
        }
    }
    __syncthreads();
    const unsigned long long best = *s_best;
    const int best_root = (int)(best & 0x3fffff) - 1;
    for (int p = tid; p < n; p += SLOW_THREADS) {
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

__global__ void __launch_bounds__(SLOW_THREADS) k_tail_slow(const __grid_constant__ LmBatch b, const int *need_slow) {
    extern __shared__ int sm[];
    const int f = blockIdx.x, tid = threadIdx.x;
    if (!need_slow[f]) return;
    const int cols = b.tail_w, pitch = b.tail_pitch;
    int *colany = sm, *cnt_b = sm + cols, *sum_b = sm + 2 * cols, *cnt_s = sm + 3 * cols, *sum_s = sm + 4 * cols;
    __shared__ unsigned long long s_best;
    __shared__ int s_first, s_last;
    int *L = b.cc + (int64_t)f * 3 * b.cc_stride, *area = L + b.cc_stride, *key = area + b.cc_stride;
    const int hb = b.bb_h[LM_BOTTOM], hs = b.bb_h[LM_SIDE];
    const uint8_t *bin_b = b.tailbin[LM_BOTTOM] + (int64_t)f * hb * pitch;
    const uint8_t *bin_s = b.tailbin[LM_SIDE] + (int64_t)f * hs * pitch;
    uint8_t *mask_b = b.tailmask + (int64_t)f * hb * pitch;
    uint8_t *mask_s = b.sidemask + (int64_t)f * hs * pitch;
    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
    }
    cc_largest(bin_b, nullptr, hb, cols, pitch, b.conn, L, area, key, mask_b, cnt_b, sum_b, colany, &s_best);
    for (int c = tid; c < cols; c += SLOW_THREADS)
        if (colany[c]) {
            atomicMin(&s_first, c);
            atomicMax(&s_last, c);
        }
    cc_largest(bin_s, colany, hs, cols, pitch, b.conn, L, area, key, mask_s, cnt_s, sum_s, nullptr, &s_best);
    write_tracks(b, f, s_first, s_last, cnt_b, sum_b, cnt_s, sum_s);
}

size_t tail_smem(const LmBatch &b, int runcap) {
    const size_t rcap = (size_t)((runcap > 1 ? runcap : 1) + 63) & ~(size_t)63;
    const int hmax = b.bb_h[0] > b.bb_h[1] ? b.bb_h[0] : b.bb_h[1];
    const int wpr = (b.tail_w + 31) >> 5;
    size_t s = (size_t)hmax * wpr * 4 * 2 + (size_t)(hmax + 1) * 4 + rcap * 4 * 3 + (size_t)b.tail_w * 4 * 5 +
               rcap * 2 * 3;
    return (s + 15) & ~(size_t)15;
}

}  // namespace

int lm_launch_tail(const LmBatch &b, cudaStream_t s) {
    if (b.tail_w <= 0) return 0;
    static LmDevOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    int *need_slow = b.cc_flag;
    // Run capacity of the shared-memory path; frames with more runs take k_tail_slow.  The default keeps the footprint
    // below a quarter of an SM (4 CTAs per SM: one wave for a 512-frame sub-batch).  LM_TAIL_RUNCAP overrides it (tests
    // lower it to exercise the slow path).
    int runcap = TAIL_RUNCAP_DEFAULT;
    if (const char *e = getenv("LM_TAIL_RUNCAP")) {
        int v = atoi(e);
        if (v >= 0 && v <= RUNCAP) runcap = v;
    }
    const size_t smem = tail_smem(b, runcap);
    k_tail<<<b.B, TAIL_THREADS, smem, s>>>(b, need_slow, runcap);
    k_tail_slow<<<b.B, SLOW_THREADS, (size_t)(5 * b.tail_w) * sizeof(int), s>>>(b, need_slow);
    return 2;
}
