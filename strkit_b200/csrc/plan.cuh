// plan.cuh -- device-side validation and work planning of a batch.
//
// The host batcher hands over raw arrays; everything per-read (argument checks, locus index of each read,
// kernel / rows-per-lane class, the segmented work order, scratch sizing maxima) is computed here so that
// the host-buffer C-ABI call costs one H2D copy plus a few microsecond-scale kernels, not several passes
// over millions of reads on one CPU core.
#pragma once
#include "dp_packed.cuh"
#include "strk_common.cuh"

struct PlanStats {
    unsigned long long first_error;  // (index << 8 | kind), smallest wins; ~0 = none
    unsigned bin_cnt[STRK_PK_NBIN];  // reads per class: [0] general kernel only, [R] packed kernel with R rows per lane
    unsigned bin_cursor[STRK_PK_NBIN];
    int bin_mmax[STRK_PK_NBIN], bin_flank[STRK_PK_NBIN];
    int max_n1;
    int mb_cols_base;  // multi-pass (db > 512) reads: max of max(fl, fr) + m * est
    int mb_m;          //                              max motif length
    unsigned n_dup;    // reads that share an earlier read's table (dedupe.cuh)
};

enum PlanError {
    PLAN_ERR_NEG_LEN = 1, PLAN_ERR_EMPTY = 2, PLAN_ERR_TOO_LONG = 3, PLAN_ERR_PAST_ARENA = 4, PLAN_ERR_EST = 5,
    PLAN_ERR_MOTIF_EMPTY = 6, PLAN_ERR_MOTIF_PAST_ARENA = 7, PLAN_ERR_READ_BEGIN = 8
};

__device__ __forceinline__ void plan_report(PlanStats *st, long long index, int kind) {
    atomicMin(&st->first_error, ((unsigned long long)index << 8) | (unsigned long long)kind);
}

// one thread per locus: motif checks, read_begin monotonicity, read -> locus map
__global__ void plan_loci_kernel(const long long *__restrict__ read_begin, long long n_loci, long long n_reads,
                                 const unsigned long long *__restrict__ motif_off, const int *__restrict__ motif_len,
                                 unsigned long long arena_bytes, int *__restrict__ read_locus, PlanStats *st) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_loci) return;
    const long long r0 = read_begin[l], r1 = read_begin[l + 1];
    if (r1 < r0 || r0 < 0 || r1 > n_reads) {
        plan_report(st, l, PLAN_ERR_READ_BEGIN);
        return;
    }
    if (motif_len[l] <= 0)
        plan_report(st, l, PLAN_ERR_MOTIF_EMPTY);
    else if (motif_off[l] > arena_bytes || (unsigned long long)motif_len[l] > arena_bytes - motif_off[l])
        plan_report(st, l, PLAN_ERR_MOTIF_PAST_ARENA);
    for (long long r = r0; r < r1; ++r) read_locus[r] = (int)l;
}

// one thread per read: argument checks, class, per-class histogram and maxima
__global__ void plan_reads_scan_kernel(const unsigned long long *__restrict__ seq_off, const int *__restrict__ lens,
                                       const int *__restrict__ est_cn, const int *__restrict__ read_locus,
                                       const int *__restrict__ motif_len, long long n_reads,
                                       unsigned long long arena_bytes, int packed_ok, unsigned char *__restrict__ bin,
                                       PlanStats *st) {
    __shared__ unsigned s_cnt[STRK_PK_NBIN];
    __shared__ int s_mmax[STRK_PK_NBIN], s_flank[STRK_PK_NBIN], s_max_n1, s_mb_cols, s_mb_m;
    if (threadIdx.x < STRK_PK_NBIN) s_cnt[threadIdx.x] = 0, s_mmax[threadIdx.x] = 0, s_flank[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_max_n1 = 0, s_mb_cols = 0, s_mb_m = 0;
    __syncthreads();
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) {
        const int fl = lens[3 * r], tr = lens[3 * r + 1], fr = lens[3 * r + 2];
        const long long n1 = (long long)fl + tr + fr;
        const int est = est_cn[r];
        int b = 0;
        if (fl < 0 || tr < 0 || fr < 0)
            plan_report(st, r, PLAN_ERR_NEG_LEN);
        else if (n1 <= 0)
            plan_report(st, r, PLAN_ERR_EMPTY);
        else if (n1 > (1 << 24))
            plan_report(st, r, PLAN_ERR_TOO_LONG);
        else if (seq_off[r] > arena_bytes || (unsigned long long)n1 > arena_bytes - seq_off[r])  // (no wrap-around)
            plan_report(st, r, PLAN_ERR_PAST_ARENA);
        else if (est < 0 || est > (1 << 22) || (long long)motif_len[read_locus[r]] * (long long)est > (1ll << 24))
            plan_report(st, r, PLAN_ERR_EST);  // candidate columns (64-bit): a garbage estimate must not size scratch
        else {
            const int m = motif_len[read_locus[r]];
            int R = packed_ok ? strk_pick_rows_packed((int)n1 + 1) : 0;
            if (fl < 1 || fr < 1 || fl > PK_FLANK_MAX || fr > PK_FLANK_MAX || m * R > 128 || m <= 0) R = 0;
            b = R;
            atomicMax(&s_max_n1, (int)n1);
            if (b) {
                atomicMax(&s_mmax[b], m);
                atomicMax(&s_flank[b], fl > fr ? fl : fr);
            }
            if (n1 > 32 * 16 && m > 0) {
                atomicMax(&s_mb_cols, (fl > fr ? fl : fr) + m * est);
                atomicMax(&s_mb_m, m);
            }
        }
        bin[r] = (unsigned char)b;
        atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < STRK_PK_NBIN) {
        if (s_cnt[threadIdx.x]) atomicAdd(&st->bin_cnt[threadIdx.x], s_cnt[threadIdx.x]);
        if (s_mmax[threadIdx.x]) atomicMax(&st->bin_mmax[threadIdx.x], s_mmax[threadIdx.x]);
        if (s_flank[threadIdx.x]) atomicMax(&st->bin_flank[threadIdx.x], s_flank[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        atomicMax(&st->max_n1, s_max_n1);
        if (s_mb_cols) atomicMax(&st->mb_cols_base, s_mb_cols);
        if (s_mb_m) atomicMax(&st->mb_m, s_mb_m);
    }
}

// segmented work order: class k occupies order[bin_off[k] .. bin_off[k] + bin_cnt[k])
//   rep != nullptr: reads with rep[r] != r share an earlier read's table and get no work item
__global__ void plan_reads_scatter_kernel(const unsigned char *__restrict__ bin, long long n_reads,
                                          const unsigned *__restrict__ bin_off, PlanStats *st, int *__restrict__ order,
                                          const int *__restrict__ rep) {
    __shared__ unsigned s_cnt[STRK_PK_NBIN], s_base[STRK_PK_NBIN];
    if (threadIdx.x < STRK_PK_NBIN) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0;
    unsigned local = 0;
    const bool mine = r < n_reads && (!rep || rep[r] == (int)r);
    if (mine) {
        b = bin[r];
        local = atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < STRK_PK_NBIN && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&st->bin_cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (mine) order[bin_off[b] + s_base[b] + local] = (int)r;
}

// Nibble-packed host arenas (STRK_ARENA_NIBBLE): two symbols per byte, low nibble first, code = index into
// "ACGTRYSWKMBDHVNX" (align_matrix.py:25-26).  Halves the H2D bytes of a block; this kernel expands the packed copy
// into the byte-per-symbol arena every DP kernel stages from, right after the copy lands.  One 128-bit load and two
// 128-bit stores per thread, coalesced; memory-bound (0.45 GB per million reads, < 0.1 ms at HBM speed).
__global__ void expand_nibbles_kernel(const uint4 *__restrict__ packed, unsigned long long n_bytes,
                                      unsigned char *__restrict__ out) {
    __shared__ unsigned char lut[16];
    if (threadIdx.x < 16) lut[threadIdx.x] = (unsigned char)"ACGTRYSWKMBDHVNX"[threadIdx.x];
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n_quads = n_bytes >> 4;
    if (i < n_quads) {
        const uint4 v = packed[i];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        unsigned o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned x = w[k] >> (16 * h);
                o[2 * k + h] = (unsigned)lut[x & 15u] | ((unsigned)lut[(x >> 4) & 15u] << 8) |
                               ((unsigned)lut[(x >> 8) & 15u] << 16) | ((unsigned)lut[(x >> 12) & 15u] << 24);
            }
        }
        uint4 *dst = (uint4 *)(out + 32ull * i);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else if (i == n_quads) {  // tail: fewer than 16 packed bytes
        const unsigned char *pb = (const unsigned char *)packed;
        for (unsigned long long b = n_quads << 4; b < n_bytes; ++b) {
            out[2 * b] = lut[pb[b] & 15u];
            out[2 * b + 1] = lut[pb[b] >> 4];
        }
    }
}
