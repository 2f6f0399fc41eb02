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
    unsigned long long h = 0x9E3779B97F4A7C15ull * (unsigned long long)(lane + 1);
    for (int i = lane; i < n1; i += 32) {
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
                bool same = true;
                for (int i = lane; i < n1; i += 32) same = same && a[i] == b[i];
                if (__all_sync(0xffffffffu, same)) mine = q;
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
