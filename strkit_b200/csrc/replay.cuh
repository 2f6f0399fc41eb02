// replay.cuh -- exact replay of the reference's hill-climb over precomputed score tables.
//
// The DP kernels score a whole window of candidate sizes in parallel; the reference instead walks
// sizes serially (strkit_rust_ext.get_repeat_count, called at repeats.py:58-68; the in-tree
// statement of the same search is repeats.py:100-156) and its answer depends on the start size,
// the visitation order, the iteration budget and the tie-breaks -- it is NOT an arg-max.  These
// routines run that search verbatim with every "score this size" replaced by a table look-up, so
// the parallel sweep returns bit-identical (n, score, n_explored).  A look-up outside the table
// window is reported as a miss; the host widens the window and runs the locus again.
//
// Also replayed here: the per-locus read loop of call_locus.py:1079,1129-1161 (start guess with the
// carried offset fraction; float64, round-half-even like CPython's round()).
#pragma once
#include "strk_common.cuh"

#define STRK_MAX_WINDOW 4100  // widest table row the replay can track (bitmap of visited sizes)
#define STRK_SEEN_WORDS ((STRK_MAX_WINDOW + 31) / 32)

struct ClimbResult {
    int best_n, best_score, n_explored;
    int status;  // 0 ok, 1 window miss, 2 nothing scored
    double sum_n;  // sum of the sizes scored (for reference-equivalent cell counts)
    int lo_touched, hi_touched;  // extent of the sizes looked up
};

struct SeenSet {
    unsigned int w[STRK_SEEN_WORDS];
    int lo, hi;
    __host__ __device__ void reset(int lo_, int hi_) {
        lo = lo_;
        hi = hi_;
        int nw = (hi_ - lo_ + 32) / 32;
        for (int k = 0; k < nw; ++k) w[k] = 0u;
    }
    __host__ __device__ bool inside(int n) const { return n >= lo && n <= hi; }
    // sizes outside the window can never have been scored
    __host__ __device__ bool has(int n) const {
        if (!inside(n)) return false;
        int k = n - lo;
        return (w[k >> 5] >> (k & 31)) & 1u;
    }
    __host__ __device__ void add(int n) {
        int k = n - lo;
        w[k >> 5] |= 1u << (k & 31);
    }
};

// The same set for windows of at most 32 sizes: one register instead of a bitmap in local memory.
struct SeenSmall {
    unsigned int w;
    int lo, hi;
    __host__ __device__ void reset(int lo_, int hi_) { lo = lo_, hi = hi_, w = 0u; }
    __host__ __device__ bool inside(int n) const { return n >= lo && n <= hi; }
    __host__ __device__ bool has(int n) const { return inside(n) && ((w >> (n - lo)) & 1u); }
    __host__ __device__ void add(int n) { w |= 1u << (n - lo); }
};

// Single-score search (read path).  table[n - n_lo] for n in [n_lo, n_hi].
template <typename Seen>
__host__ __device__ inline ClimbResult climb_single(const int *table, int n_lo, int n_hi, int start_count,
                                                    int max_iters, int range, int step, int tie_flags,
                                                    Seen &seen) {
    ClimbResult res;
    res.best_n = 0;
    res.best_score = 0;
    res.n_explored = 0;
    res.status = 0;
    res.sum_n = 0.0;
    res.lo_touched = 0x7fffffff;
    res.hi_touched = -1;
    seen.reset(n_lo, n_hi);
    int st_size[4], st_dir[4];
    int top = 0;
    st_size[top] = start_count - step, st_dir[top++] = -1;
    st_size[top] = start_count + step, st_dir[top++] = 1;
    st_size[top] = start_count, st_dir[top++] = 0;
    bool have_best = false;
    // Search-policy switches (tie_flags bits 2-3, STRK_SEARCH_* in the header): the Rust body of this search is not in
    // the reference tree and repeat_count_params.py:13 says the range "can be narrowed within the get_repeat_count
    // fn".  0 = the in-tree statement (range fixed, repeats.py:100-151); the two narrowing hypotheses only ever shrink
    // the windows, so the score table of a pass always covers them.
    while (top > 0 && res.n_explored < max_iters) {
        --top;
        const int size = st_size[top], dir = st_dir[top];
        if (size < 0) continue;
        const bool wide = step > range;
        int start_size = size - ((dir < 1 || wide) ? range : 0);
        if (start_size < 0) start_size = 0;
        const int end_size = size + ((dir > -1 || wide) ? range : 0);
        if ((tie_flags & 4) && range > 1) range = 1;              // STRK_SEARCH_NARROW_FIRST: after the first window
        if ((tie_flags & 8) && range > 1) range = range / 2;      // STRK_SEARCH_NARROW_HALVE: after every window
        res.lo_touched = start_size < res.lo_touched ? start_size : res.lo_touched;
        res.hi_touched = end_size > res.hi_touched ? end_size : res.hi_touched;
        bool have = false;
        int mv_size = 0, mv_score = 0;
        for (int i = start_size; i <= end_size; ++i) {
            if (!seen.inside(i)) {
                res.status = 1;
                return res;
            }
            const int sc = table[i - n_lo];
            if (!seen.has(i)) {
                seen.add(i);
                ++res.n_explored;
                res.sum_n += (double)i;
                // final pick = first-inserted maximum (repeats.py:154) unless STRK_TIE_FINAL_LAST
                if (!have_best || sc > res.best_score || ((tie_flags & 2) && sc == res.best_score)) {
                    have_best = true;
                    res.best_n = i;
                    res.best_score = sc;
                }
            }
            if (!have || sc > mv_score || ((tie_flags & 1) && sc == mv_score)) {
                have = true;
                mv_size = i;
                mv_score = sc;
            }
        }
        // at most one of the two pushes can fire, so the stack never exceeds 3 entries
        if (mv_size > size) {
            const int new_rc = mv_size + step;
            if (!seen.has(new_rc) && new_rc >= 0) st_size[top] = new_rc, st_dir[top++] = 1;
        }
        if (mv_size < size) {
            const int new_rc = mv_size - step;
            if (!seen.has(new_rc) && new_rc >= 0) st_size[top] = new_rc, st_dir[top++] = -1;
        }
    }
    if (!have_best) res.status = 2;
    return res;
}

struct RefClimbResult {
    int l_offset, r_offset, n_offset_scores;
    int status;
};

// Dual-score search of get_ref_repeat_count (repeats.py:100-169).  tab[k] for k = n - n_lo holds the
// packed ARGMAX keys: fwd at tab[k], rev at tab[W + k]; score = key >> 32,
// end_query = 0x7fffffff - (key & 0xffffffff) - 1.
__host__ __device__ inline void ref_unpack(long long key, int &score, int &end_query) {
    score = (int)(key >> 32);
    end_query = 0x7fffffff - (int)(unsigned)(key & 0xffffffffll) - 1;
}

__host__ __device__ inline RefClimbResult climb_ref(const long long *tab, int n_lo, int n_hi, int start_count,
                                                   int max_iters, int range, int step, int n_fl, int n_fr,
                                                   int ref_size, int vcf_anchor_size, SeenSet &seen) {
    RefClimbResult res;
    res.l_offset = res.r_offset = res.n_offset_scores = 0;
    res.status = 0;
    const int W = n_hi - n_lo + 1;
    seen.reset(n_lo, n_hi);
    int st_size[4], st_dir[4];
    int top = 0;
    st_size[top] = start_count - step, st_dir[top++] = -1;
    st_size[top] = start_count + step, st_dir[top++] = 1;
    st_size[top] = start_count, st_dir[top++] = 0;
    bool have_best = false;
    int bf_score = 0, bf_adj = 0, br_score = 0, br_adj = 0;
    const bool wide = step > range;
    while (top > 0 && res.n_offset_scores < max_iters) {
        --top;
        const int size = st_size[top], dir = st_dir[top];
        if (size < 0) continue;
        int start_size = size - ((dir < 1 || wide) ? range : 0);
        if (start_size < 0) start_size = 0;
        const int end_size = size + ((dir > -1 || wide) ? range : 0);
        bool have = false;
        int mv_size = 0, mv_s = 0, mv_a = 0;
        // max((*fwd_scores, *rev_scores), key=(score, adj)): first maximal, fwd entries first (:135)
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = start_size; i <= end_size; ++i) {
                if (!seen.inside(i)) {
                    res.status = 1;
                    return res;
                }
                int fs, fe, rs, re;
                ref_unpack(tab[i - n_lo], fs, fe);
                ref_unpack(tab[W + i - n_lo], rs, re);
                const int r_adj = fe + 1 - n_fl - ref_size;  // :34
                const int l_adj = re + 1 - n_fr - ref_size;  // :41
                if (!seen.has(i)) {
                    seen.add(i);
                    ++res.n_offset_scores;
                    if (!have_best || fs > bf_score) bf_score = fs, bf_adj = r_adj;  // :154
                    if (!have_best || rs > br_score) br_score = rs, br_adj = l_adj;  // :156
                    have_best = true;
                }
                const int s = pass == 0 ? fs : rs, a = pass == 0 ? r_adj : l_adj;
                if (!have || s > mv_s || (s == mv_s && a > mv_a)) {
                    have = true;
                    mv_size = i;
                    mv_s = s;
                    mv_a = a;
                }
            }
        }
        if (mv_size > size) {
            const int new_rc = mv_size + step;
            if (!seen.has(new_rc) && new_rc >= 0) st_size[top] = new_rc, st_dir[top++] = 1;
        }
        if (mv_size < size) {
            const int new_rc = mv_size - step;
            if (!seen.has(new_rc) && new_rc >= 0) st_size[top] = new_rc, st_dir[top++] = -1;
        }
    }
    if (!have_best) {
        res.status = 2;
        return res;
    }
    res.l_offset = br_adj;  // :161
    res.r_offset = bf_adj;  // :162
    if (res.l_offset >= n_fl - vcf_anchor_size) res.l_offset = 0;  // :164-167
    if (res.r_offset >= n_fr) res.r_offset = 0;                    // :168-169
    return res;
}

// One thread per locus: the read loop of call_locus.py:1129-1161 over the score tables.
//   table row of slot s = table + s * W ; window of the slot = [max(0, est - wdr), est + wdr], wdr = strk_read_wd()
//   miss_count[0] += loci whose search left the window; miss_count[2] += loci that stayed inside it but went
//   beyond est +- wd, i.e. that only the wide_short margin saved from a second pass
//   locus_ids == nullptr: locus q is q and its slots are read_begin[q]..; otherwise the widening
//   pass lists the loci to redo and slot_begin[q] is the first slot of locus_ids[q].
__global__ void replay_reads_kernel(const int *__restrict__ table, int W, int wd, int wide_short,
                                    const int *__restrict__ locus_ids,
                                    const long long *__restrict__ slot_begin, int n_list,
                                    const long long *__restrict__ read_begin, const int *__restrict__ est_cn,
                                    const int *__restrict__ lens, const int *__restrict__ motif_len, int max_iters,
                                    int range, int step, int tie_flags, int *__restrict__ out,
                                    unsigned char *__restrict__ locus_status, unsigned int *miss_count,
                                    double *ref_cells, const int *__restrict__ rep, double *__restrict__ hint = nullptr) {
    // rep != nullptr (first pass only, slot == read): read r looks its scores up in the row of read rep[r]
    // hint != nullptr (widening passes): per locus {smallest, largest} offset fraction seen at a miss (strk_slot_window);
    // a locus that misses again adds the fraction it missed at
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_list) return;
    const int locus = locus_ids ? locus_ids[q] : q;
    const long long r0 = read_begin[locus], r1 = read_begin[locus + 1];
    long long slot = locus_ids ? slot_begin[q] : r0;
    SeenSet seen;
    double frac = 0.0;  // read_offset_frac_from_starting_guess (:1079)
    double cells = 0.0;
    const int m = motif_len[locus];
    const int wdr = strk_read_wd(wd, m, wide_short);
    int status = 0;
    bool margin_used = false;
    for (long long r = r0; r < r1; ++r, ++slot) {
        const int est = est_cn[r];
        int read_sc = est;
        const int off = (int)rint(frac * (double)read_sc);  // round(), :1130
        if (off < -read_sc)
            frac = 0.0;  // :1133
        else
            read_sc += off;  // :1136
        int n_lo, n_hi;
        strk_slot_window(est, m, wd, wide_short, hint ? hint + 2 * (size_t)locus : nullptr, n_lo, n_hi);
        ClimbResult cr = climb_single(table + (size_t)(rep ? (long long)rep[r] : slot) * (size_t)W, n_lo, n_hi, read_sc,
                                      max_iters, range, step, tie_flags, seen);
        if (cr.status) {
            status = cr.status;
            if (hint && cr.status == 1) {  // where the next pass has to look: the fraction this read started from
                double *h = hint + 2 * (size_t)locus;
                const double fr_used = read_sc == est ? 0.0 : frac;
                h[0] = fr_used < h[0] ? fr_used : h[0];
                h[1] = fr_used > h[1] ? fr_used : h[1];
            }
            break;
        }
        margin_used |= cr.lo_touched < est - wd || cr.hi_touched > est + wd;
        out[4 * r + 0] = cr.best_n;
        out[4 * r + 1] = cr.best_score;
        out[4 * r + 2] = cr.n_explored;
        out[4 * r + 3] = read_sc;
        frac += (double)(cr.best_n - read_sc) / (double)(cr.best_n > 1 ? cr.best_n : 1);  // :1161
        const double n1 = (double)(lens[3 * r] + lens[3 * r + 1] + lens[3 * r + 2]);
        cells += n1 * ((double)cr.n_explored * (double)(lens[3 * r] + lens[3 * r + 2]) + (double)m * cr.sum_n);
    }
    locus_status[locus] = (unsigned char)status;
    if (status) {
        atomicAdd(miss_count, 1u);
    } else {
        atomicAdd(ref_cells, cells);
        if (margin_used) atomicAdd(miss_count + 2, 1u);
    }
}

// The same loop for narrow windows (W <= REPLAY_WMAX: every first pass).  The search reads a dozen table entries per
// read one after the other, each a trip to L2 (~4 us per read, 0.11 ms for a 30-read locus whatever the batch size):
// here the row, estimate and lengths of read r + 1 are fetched as independent loads while read r is searched on a
// copy of its row in shared memory (stride REPLAY_WMAX + 1 words per thread: conflict-free).
#define REPLAY_WMAX 20
#define REPLAY_THREADS 128
__global__ void __launch_bounds__(REPLAY_THREADS)
    replay_reads_small_kernel(const int *__restrict__ table, int W, int wd, int wide_short,
                              const int *__restrict__ locus_ids, const long long *__restrict__ slot_begin, int n_list,
                              const long long *__restrict__ read_begin, const int *__restrict__ est_cn,
                              const int *__restrict__ lens, const int *__restrict__ motif_len, int max_iters, int range,
                              int step, int tie_flags, int *__restrict__ out, unsigned char *__restrict__ locus_status,
                              unsigned int *miss_count, double *ref_cells, const int *__restrict__ rep,
                              double *__restrict__ hint = nullptr) {
    // hint != nullptr: a locus that misses records the offset fraction it missed at (first pass: windows are est +- wdr)
    __shared__ int rows[REPLAY_THREADS * (REPLAY_WMAX + 1)];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_list) return;
    int *row = rows + threadIdx.x * (REPLAY_WMAX + 1);
    const int locus = locus_ids ? locus_ids[q] : q;
    const long long r0 = read_begin[locus], r1 = read_begin[locus + 1];
    long long slot = locus_ids ? slot_begin[q] : r0;
    SeenSmall seen;
    double frac = 0.0;
    double cells = 0.0;
    const int m = motif_len[locus];
    const int wdr = strk_read_wd(wd, m, wide_short);
    int status = 0;
    bool margin_used = false;
    int nrow[REPLAY_WMAX], nest = 0, nl0 = 0, nl1 = 0, nl2 = 0;
    auto fetch = [&](long long r, long long sl) {
        const int *src = table + (size_t)(rep ? (long long)rep[r] : sl) * (size_t)W;
#pragma unroll
        for (int k = 0; k < REPLAY_WMAX; ++k) nrow[k] = k < W ? src[k] : 0;
        nest = est_cn[r];
        nl0 = lens[3 * r], nl1 = lens[3 * r + 1], nl2 = lens[3 * r + 2];
    };
    if (r0 < r1) fetch(r0, slot);
    for (long long r = r0; r < r1; ++r, ++slot) {
#pragma unroll
        for (int k = 0; k < REPLAY_WMAX; ++k) row[k] = nrow[k];
        const int est = nest;
        const double n1 = (double)(nl0 + nl1 + nl2), fl_fr = (double)(nl0 + nl2);
        if (r + 1 < r1) fetch(r + 1, slot + 1);
        int read_sc = est;
        const int off = (int)rint(frac * (double)read_sc);  // round(), :1130
        if (off < -read_sc)
            frac = 0.0;  // :1133
        else
            read_sc += off;  // :1136
        const int n_lo = est - wdr > 0 ? est - wdr : 0;
        const int n_hi = est + wdr;
        ClimbResult cr = climb_single(row, n_lo, n_hi, read_sc, max_iters, range, step, tie_flags, seen);
        if (cr.status) {
            status = cr.status;
            if (hint && cr.status == 1) {
                double *h = hint + 2 * (size_t)locus;
                const double fr_used = read_sc == est ? 0.0 : frac;
                h[0] = fr_used < h[0] ? fr_used : h[0];
                h[1] = fr_used > h[1] ? fr_used : h[1];
            }
            break;
        }
        margin_used |= cr.lo_touched < est - wd || cr.hi_touched > est + wd;
        *(int4 *)(out + 4 * r) = make_int4(cr.best_n, cr.best_score, cr.n_explored, read_sc);
        frac += (double)(cr.best_n - read_sc) / (double)(cr.best_n > 1 ? cr.best_n : 1);  // :1161
        cells += n1 * ((double)cr.n_explored * fl_fr + (double)m * cr.sum_n);
    }
    locus_status[locus] = (unsigned char)status;
    if (status) {
        atomicAdd(miss_count, 1u);
    } else {
        atomicAdd(ref_cells, cells);
        if (margin_used) atomicAdd(miss_count + 2, 1u);
    }
}

// Builds the family descriptors of a pass on the device (no descriptor H2D traffic).
//   read_ids == nullptr: slot s is read s.
__global__ void plan_reads_kernel(const int *__restrict__ read_ids, long long n_slots,
                                  const unsigned long long *__restrict__ seq_off, const int *__restrict__ lens,
                                  const int *__restrict__ est_cn, const int *__restrict__ read_locus,
                                  const unsigned long long *__restrict__ motif_off, const int *__restrict__ motif_len,
                                  int wd, int wide_short, int W, FamDesc *__restrict__ fams, double *exec_cells,
                                  const int *__restrict__ rep, const double *__restrict__ hint = nullptr,
                                  unsigned int *w_needed = nullptr) {
    // w_needed != nullptr: only measure -- the widest window of the pass (atomicMax), no descriptor is written
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double cells = 0.0;
    if (s < n_slots) {
    const long long r = read_ids ? read_ids[s] : s;
    const int locus = read_locus[r];
    FamDesc f;
    f.db_off = seq_off[r];
    f.motif_off = motif_off[locus];
    f.out_off = (unsigned long long)s * (unsigned long long)W;
    f.n_fl = lens[3 * r];
    f.n_tr = lens[3 * r + 1];
    f.n_fr = lens[3 * r + 2];
    f.m = motif_len[locus];
    const int est = est_cn[r];
    strk_slot_window(est, f.m, wd, wide_short, hint ? hint + 2 * (size_t)locus : nullptr, f.n_lo, f.n_hi);
    if (w_needed) {
        atomicMax(w_needed, (unsigned int)(f.n_hi - f.n_lo + 1));
    } else {
        fams[s] = f;
        // executed DP cells: one forward sweep over fl + motif*n_hi plus one backward sweep over fr
        if (!rep || rep[r] == (int)r)  // a read that shares another read's table executes no cells
            cells = (double)(f.n_fl + f.n_tr + f.n_fr) * ((double)f.n_fl + (double)f.m * (double)f.n_hi + (double)f.n_fr);
    }
    }
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_down_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells > 0.0) atomicAdd(exec_cells, cells);
}
