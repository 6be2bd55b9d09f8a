// umma_ts_probe.cu — cycles per tcgen05.mma.kind::i8 (M = 128, K = 32) with the A operand in TENSOR MEMORY ("TS" form:
// tcgen05.mma [d], [a_tmem], b_desc, idesc, p) against the SS form k_screen2 uses, at N = 64..256, all SMs busy.
// Question it answers: an SS instruction is bound by streaming A (4 kB) + B (N * 32 B) from shared memory; if the
// Toeplitz weight images became the (resident) A operand in TMEM and the window rows the B operand, would the
// instruction run at the tensor pipe's rate?  One JSON line per configuration.
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

// mode 0 = SS (A and B from shared memory), 1 = TS (A from tensor memory, B from shared memory),
// mode 2 = "CP+TS": the k_screen2 orientation (A = shifted window rows) with every instruction's A tile first copied
//          shared -> tensor memory by tcgen05.cp.128x256b into a rotating set of four 8-column buffers (tcgen05.cp and
//          tcgen05.mma execute in issue order), then consumed in TS form: does the copy engine's shared-memory read overlap
//          the tensor pipe where the SS form's A read does not?
// mode 3 = numerical check of mode 2's operand layout: D(SS) in columns 0.., D(CP+TS) in columns 256.. on pseudo-random
//          operands; status 2 if any accumulator differs.
__global__ void __launch_bounds__(128) rate(int mode, int N, int chain, long long *cycles, int *status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int R = 160;
    const uint32_t panel_a = R * 16, panel_b = (uint32_t)(N + 32) * 16;   // B rows shift like the window rows do
    for (int i = tid; i < (int)((4 * panel_a + 4 * panel_b) / 4); i += 128)
        reinterpret_cast<uint32_t *>(smem)[i] = mode == 3 ? ((uint32_t)i * 2654435761u) >> 3 : 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t barp = smem_u32(&bar);
    int st = 0;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 4 * panel_a;
        long long t0 = clock64();
        for (int i = 0; i < chain; ++i) {
            const int j = i & 15, ks = i & 1;
            uint64_t bd = make_desc(b0 + j * 16 + ks * 2 * panel_b, panel_b);
            if (mode == 2) {
                uint64_t ad = make_desc(a0 + j * 16 + ks * 2 * panel_a, panel_a);
                const uint32_t at = tmem + 256u + (uint32_t)((i & 3) * 8);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(at), "l"(ad) : "memory");
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;}" ::"r"(tmem), "r"(at), "l"(bd), "r"(idesc), "r"(1) : "memory");
            } else if (mode == 3) {
                uint64_t ad = make_desc(a0 + j * 16 + ks * 2 * panel_a, panel_a);
                const uint32_t at = tmem + 500u;
                const int acc = i > 0;
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(at), "l"(ad) : "memory");
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;}" ::"r"(tmem + 256u), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            } else if (mode == 0) {
                uint64_t ad = make_desc(a0 + j * 16 + ks * 2 * panel_a, panel_a);
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            } else {
                // A: 128 lanes x 32 bytes = 8 tensor-memory columns per K step; 60 different weight tiles (30 rows x 2 steps)
                const uint32_t at = tmem + 256u + (uint32_t)((i % 30) * 8);
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;}" ::"r"(tmem), "r"(at), "l"(bd), "r"(idesc), "r"(1) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
        if (!mbar_wait(barp, 0)) st = 1;
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
        if (st) atomicMax(status, st);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (mode == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        int bad = 0;
        for (int c = 0; c < N; ++c) {
            uint32_t x, y;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(y) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            bad += (x != y) || (x == 0u);
        }
        if (bad) atomicMax(status, 2);
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long *dC; int *dS;
    cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int chain_t = 6000;
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {64, 128, 192, 256}) {
            if (mode == 3 && N > 192) continue;   // the check keeps its second accumulator at column 256 and A at 500
            cudaMemset(dC, 0, 8); cudaMemset(dS, 0, 4);
            const size_t smem = 4 * 160 * 16 + 4 * (N + 32) * 16;
            const int chain = mode == 3 ? 37 : chain_t;
            rate<<<nsm, 128, smem>>>(mode, N, chain, dC, dS);  // warm-up
            rate<<<nsm, 128, smem>>>(mode, N, chain, dC, dS);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0; int st = 0;
            cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            const double per = (double)c / chain;
            printf("{\"grid\": %d, \"operands\": \"%s\", \"N\": %d, \"cuda\": \"%s\", \"status\": %d, \"cycles_per_mma\": %.1f, \"mac_per_cycle_per_sm\": %.0f}\n",
                   nsm, mode == 0 ? "SS" : mode == 1 ? "TS" : mode == 2 ? "CP+TS" : "CP+TS check vs SS", N, cudaGetErrorString(e), st, per, 128.0 * N * 32.0 / per);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
