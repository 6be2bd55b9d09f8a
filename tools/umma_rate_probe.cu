// umma_rate_probe.cu — measures cycles per tcgen05.mma (M = 128, SS operands, no-swizzle K-major) for
// kind::i8 and kind::f16 (bf16) at N = 32..256 with ALL SMs busy (grid = #SMs), i.e. the rates that bound
// k_screen.cu.  One JSON line per configuration.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t ok;
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// kind: 0 = i8 (u8 x s8 -> s32), 1 = bf16 x bf16 -> f32
__global__ void __launch_bounds__(128) rate(int kind, int N, int chain, long long *cycles, int *status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int R = 160;
    const uint32_t panel_a = R * 16, panel_b = (uint32_t)N * 16;
    for (int i = tid; i < (4 * R * 16 + 4 * N * 16) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    uint32_t idesc;
    if (kind == 0) idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t barp = smem_u32(&bar);
    int st = 0;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 4 * panel_a;
        long long t0 = clock64();
        for (int i = 0; i < chain; ++i) {
            const int j = i & 15, ks = i & 1;
            uint64_t ad = make_desc(a0 + j * 16 + ks * 2 * panel_a, panel_a);
            uint64_t bd = make_desc(b0 + ks * 2 * panel_b, panel_b);
            if (kind == 0)
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            else
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
        if (!mbar_wait(barp, 0)) st = 1;
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
        if (st) atomicMax(status, st);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
    long long *dC; int *dS;
    cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int chain = 4000;
    for (int grid : {1, nsm})
        for (int kind = 0; kind < 2; ++kind)
            for (int N : {32, 64, 128, 256}) {
                cudaMemset(dC, 0, 8); cudaMemset(dS, 0, 4);
                const size_t smem = 4 * 160 * 16 + 4 * N * 16;
                rate<<<grid, 128, smem>>>(kind, N, chain, dC, dS);  // warm-up
                rate<<<grid, 128, smem>>>(kind, N, chain, dC, dS);
                cudaError_t e = cudaDeviceSynchronize();
                long long c = 0; int st = 0;
                cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
                const double per = (double)c / chain, k = kind == 0 ? 32.0 : 16.0;
                printf("{\"grid\": %d, \"kind\": \"%s\", \"N\": %d, \"cuda\": \"%s\", \"status\": %d, \"cycles_per_mma\": %.1f, \"mac_per_cycle_per_sm\": %.0f}\n",
                       grid, kind == 0 ? "i8" : "bf16", N, cudaGetErrorString(e), st, per, 128.0 * N * k / per);
            }
    return 0;
}
