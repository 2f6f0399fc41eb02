// int_peak.cuh -- micro-benchmark of the integer issue rate (the roofline denominator for the DP kernels).
// MEASURED_PEAKS.json holds HBM and bf16 peaks only; the alignment kernels are bound by the INT32 ALU
// (VIMNMX / VIADDMNMX / PRMT) and FMA-pipe integer (IMAD) issue rates, so the library measures them.
#pragma once
#include <cuda_runtime.h>

// MODE 0: ALU pipe only (VIADDMNMX chains)   MODE 1: FMA pipe only (IMAD chains)
// MODE 2: both pipes (half the chains each)  -- 8 independent chains per thread hide the 4-cycle latency
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(int iters, int seed, int *out) {
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 8 + k + seed;
    const int c1 = seed | 1, c2 = seed + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0 || (MODE == 2 && (k & 1)))
                    a[k] = __viaddmax_s32(a[k], c1, c2);
                else
                    a[k] = a[k] * c1 + c2;
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= a[k];
    if (s == 0x7fffffff) out[0] = s;  // never true in practice; keeps the chains alive
}
