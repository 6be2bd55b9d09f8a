// umma_sw_probe.cu — second descriptor experiment for k_screen.cu (the no-swizzle layout of the first one,
// tools/umma_i8_probe.cu, turned out to cost ~110 cycles per instruction whatever N is):
//   A (window tile): SWIZZLE_128B K-major, 128-byte rows, 16-byte chunk c of row r stored at chunk c ^ (r & 7);
//                    kernel row j = start address + 128*j  -> is the swizzle phase taken from the absolute
//                    address (base_offset = 0) or from the row index relative to the start (base_offset = j & 7)?
//   B (Toeplitz)   : SWIZZLE_64B K-major, 64-byte rows, chunk c of row n stored at chunk c ^ ((n >> 1) & 3).
// For each row shift it runs both base_offset variants, checks D = A[j : j+128, 0:64] * B^T (K = 64, two
// kind::i8 instructions) against the CPU and then times a chain of instructions with all SMs busy.
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

constexpr int R = 160, KB = 64, N = 64, M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw(uint32_t addr, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t ok;
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) probe(const uint8_t *A, const int8_t *B, int32_t *D, int shift, int use_base_off, int chain,
                                             long long *cycles, int *status, int NN, int kind, int tcols, int pat) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;              // [R][128 B], swizzle 128B
    uint8_t *sB = smem + R * 128;    // [N][64 B], swizzle 64B  (R*128 = 20480 is 1024-aligned)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < R * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        int4 v = make_int4(0, 0, 0, 0);
        if (c < 4) v = *reinterpret_cast<const int4 *>(A + r * KB + c * 16);
        *reinterpret_cast<int4 *>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    for (int i = tid; i < N * 4; i += 128) {
        const int n = i >> 2, c = i & 3;
        *reinterpret_cast<int4 *>(sB + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = *reinterpret_cast<const int4 *>(B + n * KB + c * 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tcols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t barp = smem_u32(&bar);
    int st = 0;
    if (tid == 0) {
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t aaddr = smem_u32(sA) + shift * 128 + ks * 32;
            uint64_t ad = desc_sw(aaddr, 1024, 2, use_base_off ? ((aaddr >> 7) & 7) : 0);
            uint64_t bd = desc_sw(smem_u32(sB) + ks * 32, 512, 4, 0);
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    }
    if (!mbar_wait(barp, 0)) st = 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (!st && blockIdx.x == 0) {
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
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0 && !st && chain > 0) {
        const uint32_t idesc2 = kind == 0 ? ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | (8u << 24))
                                          : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | (8u << 24));
        const uint64_t ad0 = desc_sw(smem_u32(sA), 1024, 2, 0), bd0 = desc_sw(smem_u32(sB), 512, 4, 0);
        long long t0 = clock64();
        if (pat >= 4) {
            // as pattern 3 but consecutive instructions alternate between (pat - 2) independent accumulators
            const int nacc = pat - 2;
            for (int i = 0; i < chain; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t ad = ad0 + (uint64_t)(((u >> 1) * 128 + (u & 1) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)(((u & 1) * 32) >> 4);
                    const uint32_t d = tmem + (uint32_t)((u % nacc) * NN);
                    if (kind == 0)
                        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc2), "r"(1) : "memory");
                    else
                        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc2), "r"(1) : "memory");
                }
            }
        } else if (pat == 3) {
            // lean issue loop: 8 instructions per iteration, descriptors = base + compile-time constant
            for (int i = 0; i < chain; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t ad = ad0 + (uint64_t)(((u >> 1) * 128 + (u & 1) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)(((u & 1) * 32) >> 4);
                    if (kind == 0)
                        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc2), "r"(1) : "memory");
                    else
                        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc2), "r"(1) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
        if (!mbar_wait(barp, 1)) st = 2;
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    if (st) atomicMax(status, st);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tcols));
}

int main() {
    std::vector<uint8_t> A(R * KB);
    std::vector<int8_t> B(N * KB);
    uint32_t s = 777u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto &a : A) a = (uint8_t)(rnd() & 255);
    for (auto &b : B) b = (int8_t)((int)(rnd() & 255) - 128);
    uint8_t *dA; int8_t *dB; int32_t *dD; long long *dC; int *dS;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const size_t smem = R * 128 + N * 64 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    std::vector<int32_t> got(M * N);
    for (int shift : {0, 1, 3, 8, 13, 29})
        for (int ubo = 0; ubo < 2; ++ubo) {
            cudaMemset(dD, 0xff, M * N * 4); cudaMemset(dS, 0, 4); cudaMemset(dC, 0, 8);
            probe<<<1, 128, smem>>>(dA, dB, dD, shift, ubo, 0, dC, dS, 64, 0, 64, 0);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
            int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    int acc = 0;
                    for (int k = 0; k < KB; ++k) acc += (int)A[(m + shift) * KB + k] * (int)B[n * KB + k];
                    bad += got[m * N + n] != acc;
                }
            printf("{\"test\": \"sw128_rowshift\", \"shift\": %d, \"base_offset_from_addr\": %d, \"cuda\": \"%s\", \"status\": %d, \"mismatches\": %d}\n",
                   shift, ubo, cudaGetErrorString(e), st, bad);
            if (e != cudaSuccess) return 1;
        }
    const size_t smem2 = R * 128 + 256 * 64 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    for (int kind = 0; kind < 2; ++kind)
        for (int NN : {64, 128, 256})
            for (int tcols : {64, 128, 256, 512})
                for (int pat = 3; pat < 7; ++pat) {
                    if (tcols != 512 || (pat >= 4 && NN * (pat - 2) > 512)) continue;
                    const int chain = 4000, grid = nsm;
                    cudaMemset(dS, 0, 4); cudaMemset(dC, 0, 8);
                    probe<<<grid, 128, smem2>>>(dA, dB, dD, 0, 0, chain, dC, dS, NN, kind, tcols, pat);
                    probe<<<grid, 128, smem2>>>(dA, dB, dD, 0, 0, chain, dC, dS, NN, kind, tcols, pat);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c = 0; int st = 0;
                    cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
                    printf("{\"kind\": \"%s\", \"N\": %d, \"tmem_cols\": %d, \"pattern\": %d, \"cuda\": \"%s\", \"status\": %d, \"cycles_per_mma\": %.1f}\n",
                           kind ? "bf16" : "i8", NN, tcols, pat, cudaGetErrorString(e), st, (double)c / chain);
                }
    return 0;
}
