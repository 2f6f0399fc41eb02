// ubench_pipes.cu -- issue-rate and latency micro-benchmarks of the instructions the packed DP kernel is made of
// (VIMNMX3.U16x2, VIADDMNMX.U16x2, PRMT, IMAD, LDS, SHFL), alone and in the mixes of the step loops.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench_pipes.cu && build/ubench_pipes
//
// Prints warp-instructions per clock per SM sub-partition (scheduler) for every test: 1.0 = the issue limit,
// 0.5 = one 16-lane pipe.  Tuning aid only; not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__);     \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ unsigned max3u(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ unsigned lds(unsigned addr) {  // volatile: keeps the loads inside the timed loop
    unsigned r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

enum { T_MAX3 = 0, T_ADDMAX, T_PRMT, T_IMAD, T_IADD, T_LDS, T_SHFL, T_MAX3_CHAIN, T_MIX_PROF, T_MIX_FLANK, T_MAX3_IMAD,
       T_MAX3_PRMT, T_MAX2, T_MIX_PROF_CHAIN, T_MIX_FLANK_VREG, T_MIX_FLANK_PLAIN, T_MAX3_IMAD_VREG, T_MAX3_IADD, T_COUNT };
static const char *names[T_COUNT] = {"VIMNMX3.U16x2 (8 indep chains)", "VIADDMNMX.U16x2 (8 indep)", "PRMT (8 indep)",
                                     "IMAD r*r+r (8 indep)", "IADD r+r via IMAD.IADD/IADD3 (8 indep)", "LDS.32 (8 indep)",
                                     "SHFL.UP (8 indep)", "VIMNMX3.U16x2 one dependent chain (latency)",
                                     "mix prof: LDS + IMAD + VIMNMX3 (8 indep cells)",
                                     "mix flank: PRMT + 2 IMAD + VIMNMX3 (8 indep cells)", "VIMNMX3 + IMAD 1:1 (8 indep)",
                                     "VIMNMX3 + PRMT 1:1 (8 indep)", "VIMNMX.U16x2 2-input (8 indep)",
                                     "mix prof, VIMNMX3 as ONE serial chain of 8 (like the kernel)",
                                     "mix flank, multiplier in a vector register", "mix flank, plain C adds (compiler's choice)",
                                     "VIMNMX3 + IMAD (vector-register multiplier) 1:1", "VIMNMX3 + plain add 1:1"};
static const int inst_per_iter[T_COUNT] = {8, 8, 8, 8, 8, 8, 8, 8, 24, 32, 16, 16, 8, 24, 32, 32, 16, 16};

template <int T>
__global__ void __launch_bounds__(128) k(int iters, unsigned seed, unsigned one, unsigned *out, long long *cyc) {
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    unsigned a[8], b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = threadIdx.x * 8 + q + seed, b[q] = seed * (q + 3);
    const unsigned c1 = seed | 1u, c2 = seed + 7u;
    const unsigned *sp = sm + (threadIdx.x & 31);
    unsigned off = 0;
    const unsigned vone = out[1 + (threadIdx.x & 31)];  // = 1, in a vector register
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned pp = (unsigned)__cvta_generic_to_shared(sp) + (u * 256 + (it & 3) * 1024) * 4;
            if (T == T_MAX3) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = max3u(a[q], b[q], c1);
            } else if (T == T_MAX2) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = __vmaxu2(a[q] ^ c1, b[q]);
            } else if (T == T_ADDMAX) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = __viaddmax_u16x2(a[q], c1, b[q]);
            } else if (T == T_PRMT) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = prmt(a[q], b[q], c1);
            } else if (T == T_IMAD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = imad(a[q], one, b[q]);
            } else if (T == T_IADD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = (a[q] + b[q]) ^ c1;
            } else if (T == T_LDS) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] ^= lds(pp + q * 128);
            } else if (T == T_SHFL) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = __shfl_up_sync(0xffffffffu, a[q], 1);
            } else if (T == T_MAX3_CHAIN) {
#pragma unroll
                for (int q = 0; q < 8; ++q) a[0] = max3u(a[0], b[q], c1 + q);
            } else if (T == T_MIX_PROF) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned t = imad(b[q], one, lds(pp + q * 128));
                    a[q] = max3u(t, a[q], c2);
                }
            } else if (T == T_MIX_PROF_CHAIN) {
                unsigned t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = imad(q ? a[q - 1] : c2, one, lds(pp + q * 128));
                unsigned uu = c1;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a[q] = max3u(t[q], uu, a[q]);
                    uu = a[q];
                }
            } else if (T == T_MIX_FLANK) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned t = imad(imad(b[q], one, prmt(c1, c2, a[q] & 0x7777u)), one, b[(q + 1) & 7]);
                    a[q] = max3u(t, a[q], c2);
                }
            } else if (T == T_MIX_FLANK_VREG) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned t = imad(imad(b[q], vone, prmt(c1, c2, a[q] & 0x7777u)), vone, b[(q + 1) & 7]);
                    a[q] = max3u(t, a[q], c2);
                }
            } else if (T == T_MIX_FLANK_PLAIN) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned t = b[q] + prmt(c1, c2, a[q] & 0x7777u) + b[(q + 1) & 7];
                    a[q] = max3u(t, a[q], c2);
                }
            } else if (T == T_MAX3_IMAD_VREG) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a[q] = max3u(a[q], b[q], c1);
                    b[q] = imad(b[q], vone, c2);
                }
            } else if (T == T_MAX3_IADD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a[q] = max3u(a[q], b[q], c1);
                    b[q] = b[q] + a[(q + 3) & 7];
                }
            } else if (T == T_MAX3_IMAD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a[q] = max3u(a[q], b[q], c1);
                    b[q] = imad(b[q], one, c2);
                }
            } else if (T == T_MAX3_PRMT) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a[q] = max3u(a[q], b[q], c1);
                    b[q] = prmt(b[q], c2, c1);
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned s = off;
#pragma unroll
    for (int q = 0; q < 8; ++q) s ^= a[q] ^ b[q];
    if (s == 0x12345678u) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// Whole-GPU rate from event timing (the per-block clock64 figure only shows the highest-priority warp: the
// scheduler arbitrates by warp id, so block 0 is not slowed down by its neighbours).
template <int T>
static void run(int warps_per_smsp, unsigned *d_out, long long *d_cyc, int n_sm) {
    const int iters = 4000;
    const int ctas_per_sm = warps_per_smsp;  // 128 threads = 4 warps = 1 per scheduler
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<T><<<n_sm * ctas_per_sm, 128>>>(100, 3u, 1u, d_out, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<T><<<n_sm * ctas_per_sm, 128>>>(iters, 3u, 1u, d_out, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double inst_per_warp = (double)iters * 4 * inst_per_iter[T];
    const double clk = 1.965e9;  // SM clock under load on this pool (nvidia-smi clocks.sm during the bench)
    printf("%-62s w/sched %d: %.3f warp-inst/clk/sched (events, all SMs) | first block alone: %.2f clk/inst\n", names[T],
           warps_per_smsp, inst_per_warp * warps_per_smsp / (ms * 1e-3 * clk), (double)cyc / inst_per_warp);
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    unsigned *d_out;
    long long *d_cyc;
    CK(cudaMalloc(&d_out, 64 * 4));
    {
        unsigned h[64];
        for (int i = 0; i < 64; ++i) h[i] = 1u;
        CK(cudaMemcpy(d_out, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&d_cyc, 64));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {1, 5, 8}) {
        run<T_MAX3>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX2>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_ADDMAX>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_PRMT>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_IMAD>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_IADD>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_LDS>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_SHFL>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX3_CHAIN>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX3_IMAD>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX3_PRMT>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MIX_PROF>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MIX_PROF_CHAIN>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MIX_FLANK>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MIX_FLANK_VREG>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MIX_FLANK_PLAIN>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX3_IMAD_VREG>(w, d_out, d_cyc, p.multiProcessorCount);
        run<T_MAX3_IADD>(w, d_out, d_cyc, p.multiProcessorCount);
    }
    return 0;
}
