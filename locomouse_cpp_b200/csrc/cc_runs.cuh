// cc_runs.cuh — largest connected component of a binary image, shared by the tail stage (k_tail.cu) and by pass 1 of the
// base class (k_bbox_base.cu).  Reference: cv::connectedComponentsWithStats + "largest area, ties -> lowest label"
// (LocoMouse_class.cpp:2744-2767 selectLargestRegion, 921-945 largestBWAreaObject); label order as OpenCV assigns it
// (oracle/lm_oracle.cpp largest_region, pinned against cv2).
//
// Fast path: components are labelled on RUNS, entirely in shared memory:
//   1. the map is packed to a row-major bit image;
//   2. run starts are  bits & ~(bits << 1 | carry) ; a block scan numbers the runs row-major, so the
//      runs of a row are contiguous and x-sorted;
//   3. every run is united with the runs of the previous row it touches (8-connectivity: [x0-1,x1+1],
//      4-connectivity: [x0,x1]) by lock-free union-find (atomicMin hooks, smaller id = root);
//   4. per root: area = sum of run lengths, OpenCV label order key = min over runs of the first 2x2 block
//      ((r>>1)*ceil(W/2) + (x0>>1)) for 8-connectivity / first pixel (r*W + x0) for 4-connectivity;
//      "largest" = max area, ties -> smallest key.
// Slow path (any number of runs): pixel union-find in global memory.
#pragma once
#include <cstdint>

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
    int *rowcnt = nullptr;    // [rows] pixels of the winner per row (optional)
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
    if (S.rowcnt)
        for (int r = tid; r < rows; r += TAIL_THREADS) S.rowcnt[r] = 0;
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
    // the input bits are not read again: S.obits may be the same memory as S.bits (k_tail does that to halve its footprint)
    for (int i = tid; i < nwords; i += TAIL_THREADS) S.obits[i] = 0u;
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
        if (S.rowcnt) atomicAdd(&S.rowcnt[r], x1 - x0 + 1);
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

}  // namespace
