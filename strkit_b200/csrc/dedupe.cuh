// dedupe.cuh -- identical reads of a locus share one score table.
//
// The reference wraps get_repeat_count in lru_cache(maxsize=512) (repeats.py:47): a read whose flanks, tract and
// start estimate equal an earlier read's costs it nothing, which is common on HiFi data (a third of config 2's
// reads).  Score tables depend on the sequences and the window only, so the DP runs once per distinct read of a
// locus; the replay still runs per read (the start guess carries an offset) on the representative's table row.
//
//   hash_reads_kernel     one warp per read: 64-bit hash of the lengths, the estimate and the bytes
//   dedupe_loci_kernel    one warp per locus: rep[r] = first earlier read of the locus with the same lengths,
//                         estimate and bytes (candidates by hash, confirmed byte by byte), else r; a duplicate is
//                         taken out of its class histogram so that the scatter kernel emits no work item for it
#pragma once
#include "plan.cuh"

__global__ void hash_reads_kernel(const unsigned char *__restrict__ arena, const unsigned long long *__restrict__ seq_off,
                                  const int *__restrict__ lens, const int *__restrict__ est_cn, long long n_reads,
                                  const PlanStats *__restrict__ st, unsigned long long *__restrict__ hash) {
    if (st->first_error != ~0ull) return;  // invalid batch: the planner reports it, nothing may be dereferenced
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const int fl = lens[3 * r], tr = lens[3 * r + 1], fr = lens[3 * r + 2];
    const int n1 = fl + tr + fr;
    const unsigned char *p = arena + seq_off[r];
    // The hash only nominates candidates (dedupe_loci_kernel confirms them byte by byte), so it samples every third
    // base: a third of the loads of this latency-bound kernel, and two reads that differ in one unsampled base are
    // told apart by the confirmation instead.
    unsigned long long h = 0x9E3779B97F4A7C15ull * (unsigned long long)(lane + 1);
#pragma unroll 4
    for (int i = 3 * lane; i < n1; i += 96) {
        h ^= (unsigned long long)p[i] + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        h *= 0x100000001B3ull;
        h ^= h >> 29;
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o) * 0xD6E8FEB86659FD93ull;
    h ^= ((unsigned long long)(unsigned)fl << 40) ^ ((unsigned long long)(unsigned)fr << 20) ^ (unsigned long long)(unsigned)tr;
    h = (h ^ (unsigned long long)(unsigned)est_cn[r]) * 0xFF51AFD7ED558CCDull;
    if (lane == 0) hash[r] = h;
}

__global__ void dedupe_loci_kernel(const unsigned char *__restrict__ arena, const unsigned long long *__restrict__ seq_off,
                                   const int *__restrict__ lens, const int *__restrict__ est_cn,
                                   const long long *__restrict__ read_begin, long long n_loci,
                                   const unsigned long long *__restrict__ hash, const unsigned char *__restrict__ bin,
                                   PlanStats *st, int *rep) {
    if (st->first_error != ~0ull) return;
    const long long l = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (l >= n_loci) return;
    const long long r0 = read_begin[l], r1 = read_begin[l + 1];
    unsigned dups = 0;
    if (r1 - r0 <= 32) {
        // Up to 32 reads: lane k keeps the key of read r0 + k in registers, so the match loop runs on shuffles and
        // votes; global memory is touched again only to confirm a candidate (and once per read for the result).
        const int n = (int)(r1 - r0);
        unsigned long long hk = 0ull;
        int fl = 0, tr = 0, fr = 0, est = 0;
        if (lane < n) {
            const long long r = r0 + lane;
            hk = hash[r];
            fl = lens[3 * r], tr = lens[3 * r + 1], fr = lens[3 * r + 2], est = est_cn[r];
        }
        unsigned rep_mask = 0u;  // reads that are representatives so far
        int my_rep = lane;
        for (int k = 0; k < n; ++k) {
            const unsigned long long h = __shfl_sync(0xffffffffu, hk, k);
            const int kfl = __shfl_sync(0xffffffffu, fl, k), ktr = __shfl_sync(0xffffffffu, tr, k);
            const int kfr = __shfl_sync(0xffffffffu, fr, k), kest = __shfl_sync(0xffffffffu, est, k);
            const bool cand = lane < k && ((rep_mask >> lane) & 1u) && hk == h && fl == kfl && tr == ktr && fr == kfr && est == kest;
            unsigned m = __ballot_sync(0xffffffffu, cand);
            int mine = k;
            if (m) {
                const int n1 = kfl + ktr + kfr;
                const unsigned char *a = arena + seq_off[r0 + k];
                while (m && mine == k) {  // candidates in read order: the first confirmed one is the representative
                    const int q = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned char *bq = arena + seq_off[r0 + q];
                    unsigned diff = 0u;  // (no short circuit: the loads of a confirmation are independent and pipeline)
#pragma unroll 4
                    for (int i = lane; i < n1; i += 32) diff |= (unsigned)(a[i] ^ bq[i]);
                    if (__all_sync(0xffffffffu, diff == 0u)) mine = q;
                }
            }
            if (mine == k) rep_mask |= 1u << k;
            if (lane == k) my_rep = mine;
        }
        if (lane < n) {
            const long long r = r0 + lane;
            rep[r] = (int)(r0 + my_rep);
            if (my_rep != lane) atomicSub(&st->bin_cnt[bin[r]], 1u);
        }
        dups = (unsigned)(n - __popc(rep_mask));
        if (lane == 0 && dups) atomicAdd(&st->n_dup, dups);
        return;
    }
    for (long long r = r0; r < r1; ++r) {
        const unsigned long long h = hash[r];
        const int fl = lens[3 * r], tr = lens[3 * r + 1], fr = lens[3 * r + 2], est = est_cn[r];
        const int n1 = fl + tr + fr;
        const unsigned char *a = arena + seq_off[r];
        long long mine = r;
        for (long long base = r0; base < r && mine == r; base += 32) {
            const long long p = base + lane;
            bool cand = false;
            if (p < r)
                cand = rep[p] == (int)p && hash[p] == h && lens[3 * p] == fl && lens[3 * p + 1] == tr && lens[3 * p + 2] == fr &&
                       est_cn[p] == est;
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m && mine == r) {  // candidates in read order: the first confirmed one is the representative
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const long long q = base + k;
                const unsigned char *b = arena + seq_off[q];
                unsigned diff = 0u;
#pragma unroll 4
                for (int i = lane; i < n1; i += 32) diff |= (unsigned)(a[i] ^ b[i]);
                if (__all_sync(0xffffffffu, diff == 0u)) mine = q;
            }
        }
        if (lane == 0) {
            rep[r] = (int)mine;
            if (mine != r) {
                atomicSub(&st->bin_cnt[bin[r]], 1u);
                ++dups;
            }
        }
        __syncwarp();  // rep[r] is read by every lane in the iterations that follow
    }
    if (lane == 0 && dups) atomicAdd(&st->n_dup, dups);
}
