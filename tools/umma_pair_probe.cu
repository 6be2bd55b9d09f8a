// umma_pair_probe.cu — what one k_screen2 instruction costs, measured in isolation.
//
// The probe replays k_screen2's MMA issue loop (csrc/k_screen2.cu: elect.sync, uniform descriptors, KH kernel rows x ks = 2
// K steps per window tile, A = shifted rows of one shared-memory tile, B = the resident Toeplitz operand, no-swizzle K-major
// 16-byte panels) WITHOUT the TMA producer and the epilogue, on every SM at once, as
//   cg = 2 : tcgen05.mma.cta_group::2 (M = 256 across a CTA pair, each CTA holds N/2 rows of B) — k_screen2's form
//   cg = 1 : tcgen05.mma.cta_group::1 (M = 128, the whole B in one CTA)                         — k_screen's form
// with knobs that separate the candidate limiters: operand bytes (N), the B working set (all KH x 2 tiles vs one tile),
// the A row shift (16 j bytes vs none), the operand data (constant vs pseudo-random: tensor-pipe power), and how many
// SMs run (all vs one pair).  One JSON line per configuration: cycles per instruction and MAC / cycle / SM.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../locomouse_cpp_b200/csrc/umma_common.cuh"

struct Cfg {
    int cg, N, KH, units, b_all, a_shift, rnd, rows;
    int mix;       // 1: units alternate like k_screen2's x tiles: 8 at N, 5 at N = 128
    int commits;   // tcgen05.commit per unit to a second barrier nobody waits on (k_screen2 commits twice per unit)
    int lds;   // tcgen05.ld.32x32b.x32 per warp (four warps, one per lane quadrant) and unit, issued while the MMAs run
};

template <int CG>
__device__ __forceinline__ void body(const Cfg c, long long *cycles, int *status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int nhalf = c.N / CG, ks = 2, npanel = 2 * ks;
    const uint32_t panel_a = (uint32_t)c.rows * 16u, chunk_b = (uint32_t)nhalf * 16u;
    const uint32_t b_bytes = (uint32_t)c.KH * npanel * chunk_b, a_bytes = npanel * panel_a;
    uint8_t *sB = smem, *sA = smem + b_bytes;
    for (uint32_t i = tid; i < (b_bytes + a_bytes) / 4; i += blockDim.x) {
        uint32_t v = 0x01010101u;
        if (c.rnd) {
            v = (i + 977u * blockIdx.x) * 2654435761u;
            v ^= v >> 15;
            v *= 2246822519u;
            v ^= v >> 13;
        }
        reinterpret_cast<uint32_t *>(smem)[i] = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc_w = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const uint32_t idesc_n = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const uint64_t bdesc0 = umma_desc(smem_u32(sB), chunk_b);
        const uint32_t a_step = (2u * panel_a) >> 4, b_step = (2u * chunk_b) >> 4;
        const uint32_t b_row = c.b_all ? ((uint32_t)npanel * chunk_b) >> 4 : 0u;
        const uint32_t a_row = c.a_shift ? 1u : 0u;
        const uint32_t barp = smem_u32(&bar);
        long long t0 = clock64();
        for (int u = 0; u < c.units; ++u) {
            const uint32_t idesc = (c.mix && (u % 13) >= 8) ? idesc_n : idesc_w;
            if (elect_one()) {
                const uint32_t d = tmem + (uint32_t)(u & 1) * 256u;
                uint64_t adesc = umma_desc(smem_u32(sA), panel_a);
                uint64_t bdesc = bdesc0;
                uint32_t accum = 0;
                for (int j = 0; j < c.KH; ++j) {
                    uint64_t ad = adesc, bd = bdesc;
                    for (int k = 0; k < ks; ++k) {
                        if (CG == 2) umma_i8_2cta(d, ad, bd, idesc, accum);
                        else umma_i8(d, ad, bd, idesc, accum);
                        accum = 1;
                        ad += a_step;
                        bd += b_step;
                    }
                    adesc += a_row;
                    bdesc += b_row;
                }
                for (int q = 0; q < c.commits; ++q) {
                    if (CG == 2) umma_commit_2cta(smem_u32(&bar2));
                    else umma_commit(smem_u32(&bar2));
                }
                if (u == c.units - 1) {
                    if (CG == 2) umma_commit_2cta(barp);
                    else umma_commit(barp);
                }
            }
            __syncwarp();
        }
        mbar_wait(barp, 0);
        long long t1 = clock64();
        if (tid == 0) cycles[2 + blockIdx.x] = t1 - t0;   // every issuing CTA reports: the kernel ends with the slowest pair
    } else if (warp == 0 && CG == 2) {
        mbar_wait(smem_u32(&bar), 0);   // the multicast commit also lands here
    } else if (warp >= 1 && c.lds > 0) {
        // epilogue-like tensor-memory reads beside the running MMAs (both CTAs of a pair): does tcgen05.ld slow the tensor pipe?
        uint32_t sink = 0;
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const long long l0 = clock64();
        for (int i = 0; i < c.units * c.lds; ++i) {
            uint32_t v[32];
            tmem_ld32(ta + (uint32_t)((i * 32) % 448), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 32; ++q) sink ^= v[q];
        }
        const long long l1 = clock64();
        if (blockIdx.x == 0 && tid == 32) cycles[1] = l1 - l0;
        if (sink == 0x12345678u) status[0] = 7;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    (void)status;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) probe2(const Cfg c, long long *cycles, int *status) { body<2>(c, cycles, status); }
__global__ void __launch_bounds__(160, 1) probe1(const Cfg c, long long *cycles, int *status) { body<1>(c, cycles, status); }

// a memory-streaming kernel to run beside the MMA chain (k_minmax-like: 256 threads, reads `n16` 16-byte words once)
__global__ void __launch_bounds__(256) stream_beside(const uint4 *src, long long n16, int *sink) {
    uint32_t acc = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256) {
        const uint4 v = __ldg(src + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = 2;
}
// an ALU-only kernel beside the chain (no memory traffic): 256 threads spinning for `spin` cycles
__global__ void __launch_bounds__(256) alu_beside(long long spin, int *sink) {
    const long long t0 = clock64();
    int acc = 0;
    while (clock64() - t0 < spin) ++acc;
    if (acc == 123456789) *sink = acc;
}

int main(int argc, char **argv) {
    long long *dC;
    int *dS;
    cudaMalloc(&dC, 8 * 160);
    cudaMalloc(&dS, 4);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(probe1, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    const int rows = 160;
    auto run = [&](int cg, int N, int KH, int b_all, int a_shift, int rnd, int grid, const char *note, int lds = 0, int commits = 0, int units = 120, int mix = 0) {
        Cfg c{cg, N, KH, units, b_all, a_shift, rnd, rows, mix, commits, lds};
        const size_t smem = (size_t)KH * 4 * (N / cg) * 16 + 4 * rows * 16;
        if (smem > 224 * 1024) return;
        cudaMemset(dC, 0, 8 * 160);
        for (int rep = 0; rep < 2; ++rep) {
            if (cg == 2) probe2<<<grid, 160, smem>>>(c, dC, dS);
            else probe1<<<grid, 160, smem>>>(c, dC, dS);
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc2[160];
        cudaMemcpy(cyc2, dC, 8 * 160, cudaMemcpyDeviceToHost);
        long long cyc = 0, cmin = 1LL << 62;
        for (int b = 0; b < grid; b += cg) {
            cyc = cyc2[2 + b] > cyc ? cyc2[2 + b] : cyc;
            cmin = cyc2[2 + b] < cmin ? cyc2[2 + b] : cmin;
        }
        const double per = (double)cyc / ((double)c.units * KH * 2);
        printf("{\"cg\": %d, \"M\": %d, \"N\": %d, \"KH\": %d, \"b_tiles\": \"%s\", \"a_shift\": %d, \"data\": \"%s\", \"grid\": %d, \"smem_kb\": %.0f, \"cuda\": \"%s\", "
               "\"units\": %d, \"mix\": %d, \"commits_per_unit\": %d, \"tmem_ld_x32_per_warp_per_unit\": %d, \"cycles_per_tmem_ld_x32\": %.1f, \"cycles_per_mma\": %.1f, \"cycles_per_mma_fastest_cta\": %.1f, \"mac_per_cycle_per_sm\": %.0f, \"frac_of_8192\": %.3f, \"note\": \"%s\"}\n",
               cg, 128 * cg, N, KH, b_all ? "all" : "one", a_shift, rnd ? "random" : "const", grid, smem / 1024.0, cudaGetErrorString(e), units, mix, commits, lds, lds ? (double)cyc2[1] / ((double)c.units * lds) : 0.0, per, (double)cmin / ((double)c.units * KH * 2),
               128.0 * N * 32.0 / per, 128.0 * N * 32.0 / per / 8192.0, note);
        fflush(stdout);
        if (e != cudaSuccess) exit(1);
    };
    if (argc > 1 && !strcmp(argv[1], "beside")) {
        // Does another kernel make progress on the SMs while the pair instruction streams its operands from shared memory at
        // ~106 B/clk, and what does it cost the MMA chain?  A = the N = 192 chain (1200 units, ~3.5 ms) on a high-priority stream,
        // B = a streaming read of 348 MB (k_minmax's traffic per 512 frames) or an ALU spin, started ~100 us later on another stream.
        cudaStream_t sa, sb;
        int lo, hi;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi);
        cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
        cudaEvent_t a0, a1, b0, b1;
        cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
        uint4 *big;
        const long long nbytes = 348LL << 20;
        cudaMalloc(&big, nbytes);
        cudaMemset(big, 1, nbytes);
        Cfg c{2, 192, 30, 1200, 1, 1, 1, rows, 0, 2, 0};
        const size_t smem = (size_t)30 * 4 * 96 * 16 + 4 * rows * 16;
        for (int kind = 0; kind < 3; ++kind) {
            auto launch_b = [&]() {
                if (kind == 0) stream_beside<<<592, 256, 0, sb>>>(big, nbytes / 16, dS);
                if (kind == 1) stream_beside<<<2688, 256, 0, sb>>>(big, nbytes / 16, dS);
                if (kind == 2) alu_beside<<<592, 256, 0, sb>>>(200000, dS);
            };
            launch_b();                                     // warm (module load) and time alone
            cudaDeviceSynchronize();
            cudaEventRecord(b0, sb);
            launch_b();
            cudaEventRecord(b1, sb);
            cudaDeviceSynchronize();
            float alone = 0;
            cudaEventElapsedTime(&alone, b0, b1);
            probe2<<<nsm, 160, smem>>>(c, dC, dS);       // warm
            cudaDeviceSynchronize();
            cudaMemset(dC, 0, 8 * 160);
            cudaEventRecord(a0, sa);
            probe2<<<nsm, 160, smem, sa>>>(c, dC, dS);
            cudaEventRecord(a1, sa);
            alu_beside<<<1, 256, 0, sb>>>(200000, dS);      // ~100 us delay
            cudaEventRecord(b0, sb);
            launch_b();
            cudaEventRecord(b1, sb);
            cudaError_t e = cudaDeviceSynchronize();
            float ta = 0, tb0 = 0, tb1 = 0;
            cudaEventElapsedTime(&ta, a0, a1);
            cudaEventElapsedTime(&tb0, a0, b0);
            cudaEventElapsedTime(&tb1, a0, b1);
            long long cyc2[160];
            cudaMemcpy(cyc2, dC, 8 * 160, cudaMemcpyDeviceToHost);
            long long cyc = 0;
            for (int b = 0; b < nsm; b += 2) cyc = cyc2[2 + b] > cyc ? cyc2[2 + b] : cyc;
            printf("{\"beside\": \"%s\", \"cuda\": \"%s\", \"mma_chain_ms\": %.3f, \"cycles_per_mma_with_B\": %.1f, \"B_alone_ms\": %.3f, \"B_start_ms\": %.3f, \"B_end_ms\": %.3f, "
                   "\"B_ms_beside\": %.3f}\n",
                   kind == 0 ? "stream 348 MB, 592 CTAs" : kind == 1 ? "stream 348 MB, 2688 CTAs" : "ALU spin 100 us, 592 CTAs", cudaGetErrorString(e), ta,
                   (double)cyc / (1200.0 * 60.0), alone, tb0, tb1, tb1 - tb0);
        }
        return 0;
    }
    // k_screen2's own shapes first
    run(2, 192, 30, 1, 1, 1, nsm, "k_screen2 wide instruction (paw+snout+tail), real-like data");
    run(2, 128, 30, 1, 1, 1, nsm, "k_screen2 narrow instruction");
    run(2, 192, 30, 1, 1, 1, nsm, "ten times longer chain", 0, 0, 1200);
    run(2, 192, 30, 1, 1, 1, nsm, "k_screen2's unit mix: 8 wide, 5 narrow (ideal 83.7 cycles per instruction)", 0, 2, 1300, 1);
    run(2, 192, 30, 1, 1, 0, nsm, "constant data");
    run(2, 192, 30, 0, 1, 1, nsm, "one B tile re-used");
    run(2, 192, 30, 1, 0, 1, nsm, "A not shifted");
    run(2, 192, 30, 1, 1, 1, 2, "one CTA pair only");
    run(2, 256, 22, 1, 1, 1, nsm, "N = 256 (KH cut to fit shared memory)");
    run(2, 256, 22, 1, 1, 0, nsm, "N = 256 constant data");
    run(2, 64, 30, 1, 1, 1, nsm, "one-template job");
    for (int lds : {6, 12, 24, 48}) run(2, 192, 30, 1, 1, 1, nsm, "epilogue-like tcgen05.ld beside the MMAs (k_screen2 reads 6 x32 per warp and wide unit)", lds);
    run(2, 32, 1, 1, 1, 1, nsm, "tcgen05.ld almost alone (one small MMA per unit)", 48);
    for (int cm : {1, 2}) run(2, 192, 30, 1, 1, 1, nsm, "tcgen05.commit after every unit", 0, cm);
    run(2, 128, 30, 1, 1, 1, nsm, "narrow, two commits per unit", 0, 2);
    run(1, 192, 17, 1, 1, 1, nsm, "single CTA, two commits per unit", 0, 2);
    for (int lds : {4, 16}) run(2, 128, 30, 1, 1, 1, nsm, "narrow instruction with tcgen05.ld beside it (k_screen2: 4 per unit)", lds);
    for (int N : {64, 128, 192, 256}) run(1, N, N <= 64 ? 30 : (N <= 128 ? 26 : (N <= 192 ? 17 : 12)), 1, 1, 1, nsm, "single CTA (k_screen form)");
    for (int N : {64, 256}) run(1, N, N <= 64 ? 30 : 12, 1, 1, 0, nsm, "single CTA, constant data");
    return 0;
}
