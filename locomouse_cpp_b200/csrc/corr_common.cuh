// corr_common.cuh — inner-loop helpers shared by the dense (k_corr.cu) and the sparse (k_screen.cu) exact
// correlation kernels.  Both keep the oracle's per-output operation order: one fp32 accumulator
// per output, taps row-major, FFMA (one rounding) or FMUL+FADD (two roundings) per tap.
#pragma once

// NT consecutive taps of one kernel row applied to one output row of the thread's patch (TX
// accumulators), taps in increasing column order.  p[] holds the NP pixels those taps read, w points at
// the first of the NT weights (16-byte aligned in shared memory).
template <int NT, int TX, bool FMA, int NP>
__device__ __forceinline__ void corr_taps(float (&acc)[TX], const float (&p)[NP], const float *__restrict__ w) {
    constexpr int NT4 = ((NT + 3) / 4) * 4;
    const float4 *wr = reinterpret_cast<const float4 *>(w);
    float wv[NT4];
#pragma unroll
    for (int q = 0; q < NT4 / 4; ++q) {
        float4 v = wr[q];
        wv[4 * q + 0] = v.x;
        wv[4 * q + 1] = v.y;
        wv[4 * q + 2] = v.z;
        wv[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
#pragma unroll
        for (int k = 0; k < TX; ++k) {
            if (FMA)
                acc[k] = __fmaf_rn(wv[i], p[i + k], acc[k]);
            else
                acc[k] = __fadd_rn(acc[k], __fmul_rn(wv[i], p[i + k]));
        }
    }
}

template <int NP>
__device__ __forceinline__ void load_pixels(float (&p)[NP], const float *__restrict__ row) {
    const float4 *src = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int q = 0; q < NP / 4; ++q) {
        float4 v = src[q];
        p[4 * q + 0] = v.x;
        p[4 * q + 1] = v.y;
        p[4 * q + 2] = v.z;
        p[4 * q + 3] = v.w;
    }
}

