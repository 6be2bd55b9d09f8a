// ffma_peak.cu — FP32 FMA micro-benchmark: the roofline denominator for the correlation kernel.
// Measures sustained FFMA throughput for (a) the register-operand pattern the correlation kernel uses
// (weight register reused across 8 accumulators, weights streamed from shared memory) and (b) a
// constant-bank operand pattern, at the kernel's launch shape (256 threads, 2 CTAs/SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_peak ffma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__constant__ float cw[64];

template <int MODE>
__global__ void __launch_bounds__(256, 2) k_ffma(float *out, int iters) {
    __shared__ __align__(16) float wsm[32 * 32];
    for (int i = threadIdx.x; i < 32 * 32; i += 256) wsm[i] = 1.0f + 1e-7f * i;
    __syncthreads();
    float acc[32], p[40];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = threadIdx.x * 1e-3f + k;
#pragma unroll
    for (int k = 0; k < 40; ++k) p[k] = 1.0f + 1e-6f * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float wv[32];
            if (MODE == 0) {
                const float4 *wr = reinterpret_cast<const float4 *>(wsm + ((it + t) & 31) * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 v = wr[q];
                    wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 30; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == 0)
                        acc[t * 8 + k] = __fmaf_rn(wv[i], p[i + k], acc[t * 8 + k]);
                    else
                        acc[t * 8 + k] = __fmaf_rn(cw[i], p[i + k], acc[t * 8 + k]);
                }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += acc[k];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
double run(int grid, int iters, float *d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_ffma<MODE><<<grid, 256>>>(d, 10);
    cudaDeviceSynchronize();
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_ffma<MODE><<<grid, 256>>>(d, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double fma = (double)grid * 256 * iters * 4 * 30 * 8;
    return fma / (best * 1e-3) / 1e12;  // TFMA/s
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-7f * i;
    cudaMemcpyToSymbol(cw, h, sizeof h);
    int grid = prop.multiProcessorCount * 2 * 8;
    float *d;
    cudaMalloc(&d, (size_t)grid * 256 * 4);
    int iters = 2000;
    double r0 = run<0>(grid, iters, d), r1 = run<1>(grid, iters, d);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"ffma_reg_tfma\": %.3f, \"ffma_const_tfma\": %.3f, "
           "\"ffma_reg_tflops\": %.2f, \"ffma_const_tflops\": %.2f, \"nominal_tflops_at_attr_clock\": %.2f}\n",
           prop.name, prop.multiProcessorCount, clk, r0, r1, 2 * r0, 2 * r1,
           2.0 * prop.multiProcessorCount * 128 * (clk * 1e3) / 1e12);
    return 0;
}
