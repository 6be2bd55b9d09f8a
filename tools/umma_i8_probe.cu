// umma_i8_probe.cu — validates, on a real B200, the tcgen05 pieces k_screen.cu relies on before the
// full kernel is trusted:
//   * kind::i8 (u8 x s8 -> s32) with no-swizzle K-major shared-memory descriptors,
//   * SBO = 128 B so that operand rows are linear at a 16-byte pitch, which makes "start address + 16*j"
//     a j-row shift of the A operand (the banded-Toeplitz correlation trick),
//   * LBO = distance between the two 16-byte K chunks of one K=32 instruction,
//   * accumulation over several MMAs into one TMEM accumulator, tcgen05.commit -> mbarrier, tcgen05.ld.
// It computes D = sum_j A[j : j+128, :] * B_j^T (3 row shifts, K = 64) and compares with the CPU, then
// times a long chain of MMAs.  Prints one JSON line.  Every wait is bounded (no hang on a bad descriptor).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

constexpr int R = 160, KB = 64, N = 64, M = 128, NJ = 3;
constexpr int PANEL_A = R * 16, PANEL_B = N * 16;
__device__ const int kShift[NJ] = {0, 1, 29};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) probe(const uint8_t *A, const int8_t *B, int32_t *D, long long *cycles, int *status, int chain) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;                      // [4 panels][R][16]
    uint8_t *sB = smem + 4 * PANEL_A;        // [NJ][4 chunks][N][16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < R * 4; i += 128) {  // A[r][16p..16p+15] -> panel p, row r
        int r = i >> 2, p = i & 3;
        *reinterpret_cast<int4 *>(sA + p * PANEL_A + r * 16) = *reinterpret_cast<const int4 *>(A + r * KB + p * 16);
    }
    for (int i = tid; i < NJ * N * 4; i += 128) {
        int j = i / (N * 4), rem = i - j * N * 4, n = rem >> 2, c = rem & 3;
        *reinterpret_cast<int4 *>(sB + j * 4 * PANEL_B + c * PANEL_B + n * 16) =
            *reinterpret_cast<const int4 *>(B + (size_t)j * N * KB + n * KB + c * 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    // u8 x s8 -> s32, K-major A and B, N = 64, M = 128
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t barp = smem_u32(&bar);
    int st = 0;
    if (tid == 0) {
        int first = 1;
        for (int j = 0; j < NJ; ++j)
            for (int ks = 0; ks < 2; ++ks) {
                uint64_t ad = make_desc(smem_u32(sA) + kShift[j] * 16 + ks * 2 * PANEL_A, PANEL_A, 128);
                uint64_t bd = make_desc(smem_u32(sB) + j * 4 * PANEL_B + ks * 2 * PANEL_B, PANEL_B, 128);
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                             "l"(ad), "l"(bd), "r"(idesc), "r"(first ? 0 : 1) : "memory");
                first = 0;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    }
    if (!mbar_wait(barp, 0)) st = 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (!st) {
        uint32_t v[32];
        for (int half = 0; half < 2; ++half) {
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 32; ++c) D[tid * N + half * 32 + c] = (int32_t)v[c];
        }
    }
    // ---- timing: `chain` back-to-back MMAs (accumulating garbage), one commit --------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0 && !st) {
        long long t0 = clock64();
        for (int i = 0; i < chain; ++i) {
            const int j = i % 30, ks = i & 1;
            uint64_t ad = make_desc(smem_u32(sA) + j * 16 + ks * 2 * PANEL_A, PANEL_A, 128);
            uint64_t bd = make_desc(smem_u32(sB) + ks * 2 * PANEL_B, PANEL_B, 128);
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad),
                         "l"(bd), "r"(idesc), "r"(1) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
        if (!mbar_wait(barp, 1)) st = 2;
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
    }
    if (st) atomicMax(status, st);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    std::vector<uint8_t> A(R * KB);
    std::vector<int8_t> B(NJ * N * KB);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto &a : A) a = (uint8_t)(rnd() & 255);
    for (auto &b : B) b = (int8_t)((int)(rnd() & 255) - 128);
    const int shifts[NJ] = {0, 1, 29};
    std::vector<int32_t> ref(M * N, 0), got(M * N, -1);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            long long acc = 0;
            for (int j = 0; j < NJ; ++j)
                for (int k = 0; k < KB; ++k) acc += (int)A[(m + shifts[j]) * KB + k] * (int)B[(size_t)j * N * KB + n * KB + k];
            ref[m * N + n] = (int32_t)acc;
        }
    uint8_t *dA; int8_t *dB; int32_t *dD; long long *dC; int *dS;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, got.size() * 4); cudaMalloc(&dC, 16); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, got.size() * 4); cudaMemset(dC, 0, 16); cudaMemset(dS, 0, 4);
    const int chain = 3000;
    const size_t smem = 4 * PANEL_A + NJ * 4 * PANEL_B;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, dC, dS, chain);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc[2] = {0, 0}; int st = -1;
    cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    int bad = 0, first_bad = -1;
    for (int i = 0; i < M * N; ++i)
        if (got[i] != ref[i]) { if (first_bad < 0) first_bad = i; ++bad; }
    printf("{\"cuda\": \"%s\", \"status\": %d, \"mismatches\": %d, \"first_bad\": %d, \"got\": %d, \"want\": %d, "
           "\"chain\": %d, \"issue_cycles\": %lld, \"total_cycles\": %lld, \"cycles_per_mma_128x64x32\": %.2f}\n",
           cudaGetErrorString(e), st, bad, first_bad, first_bad >= 0 ? got[first_bad] : 0, first_bad >= 0 ? ref[first_bad] : 0, chain,
           cyc[0], cyc[1], (double)cyc[1] / chain);
    return (e == cudaSuccess && st == 0 && bad == 0) ? 0 : 1;
}
