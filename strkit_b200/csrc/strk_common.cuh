// strk_common.cuh -- shared device/host definitions for the B200 repeat-count kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define STRK_WARP 32
#define STRK_NSYM_ 17
#define STRK_PAD_FREE 17    // pad-row code when the top border is free   (score  0)
#define STRK_PAD_PEN 18     // pad-row code when the top border is penalised (score -2g)
#define STRK_SMAT_ROWS 19

// One alignment family: a read (or a reference window) against every candidate fl + motif*n + fr.
// db = fl + tr + fr is contiguous in the arena (reference strkit/call/repeats.py:91).
struct FamDesc {
    unsigned long long db_off;     // arena offset of fl + tr + fr
    unsigned long long motif_off;  // arena offset of the motif
    unsigned long long out_off;    // element offset of this family's row in the output table
    int n_fl, n_tr, n_fr, m;
    int n_lo, n_hi;                // candidate sizes [n_lo, n_hi]
};

// Scoring constants resident on the device (one copy per context).
struct ScoreConsts {
    unsigned char lut[256];                        // ASCII -> symbol code 0..16 (case-insensitive)
    signed char smat[STRK_SMAT_ROWS * STRK_NSYM_]; // [row symbol or pad][column symbol]
    int gap;
    int end_flags;
    // packed kernel: PRMT byte tables per column symbol (row classes A,C,G,T,N,X,other,pad), biased by +2g;
    // the forward (t8f) and backward (t8b) sweeps differ in the pad-row byte only
    unsigned long long t8f[STRK_NSYM_], t8b[STRK_NSYM_];
    unsigned char cls_of[STRK_SMAT_ROWS + 1];  // symbol code -> PRMT row class, 0x80 = not representable
    // packed kernel, per row-symbol code: everything the selector build needs in one word
    //   [3:0]  forward selector nibble (class 0-3, or 8 = PRMT zero byte)   [7:4] backward nibble (4 + class, or 8)
    //   [10:8] row class 0-7 (two-table path)   [11] not representable   [12] carries an addend (N / X / other)
    //   [23:16] the addend: biased score of the class against A/C/G/T (one-table path)
    unsigned rowinfo[STRK_SMAT_ROWS + 1];
    int packed_ok;                             // every biased score fits a positive byte
    int one_table_ok;                          // N / X / other rows score the same against A, C, G and T
    // = 1 in every entry, opaque to the compiler: multiplier of the FMA-pipe adds (IMAD d, one, s).  Read as a
    // uniform value by default; PK_ONE_VREG=1 reads it per lane into a vector register (tuning switch: the
    // micro-benchmark tools/ubench_pipes.cu favours the vector-register form, the real step loops do not).
    unsigned one_v[32];
};

// Half-width (in copies) of a read's score window.  wide_short = 1 gives short motifs a wider first window: a noisy
// read's length estimate is off by a few BASES, which is more copies the shorter the motif (ONT-like reads miss a
// +-6 window at 29 % of the 2-mer loci, 4 % of the 3-mers, 1 % of the 4-mers; none at +-8 / +-8 / +-7).
__host__ __device__ inline int strk_read_wd(int wd, int m, int wide_short) {
    return wd + (wide_short ? (m <= 3 ? 2 : (m == 4 ? 1 : 0)) : 0);
}

// Score window [lo, hi] (candidate sizes) of one read in a pass.  First pass (hint == nullptr): est +- wdr.  Widening
// passes: the start count of a read is its estimate plus a carried offset, round(frac * est) (call_locus.py:1129-1136),
// and on loci whose reads alternate between a short and an expanded allele that offset is tens of copies -- so the
// window is stretched towards the starts the locus' offset fraction has been seen to produce so far (hint[0] = smallest,
// hint[1] = largest fraction at a miss) instead of being widened blindly around the estimate: a 6 kb expansion that
// starts 75 copies off needs ~100 sizes, not 385.  Used by the table planner and by the replay: they must agree.
__host__ __device__ inline void strk_slot_window(int est, int m, int wd, int wide_short, const double *hint, int &lo, int &hi) {
    const int wdr = strk_read_wd(wd, m, wide_short);
    int s_lo = 0, s_hi = 0;
    if (hint) {
        const double a = rint(hint[0] * (double)est), b = rint(hint[1] * (double)est);
        // (an offset below -est is dropped by the reference: the read starts at its estimate)
        const int ia = a < -(double)est ? 0 : (int)a, ib = b < -(double)est ? 0 : (int)b;
        s_lo = ia < ib ? ia : ib;
        s_hi = ia < ib ? ib : ia;
        if (s_lo > 0) s_lo = 0;
        if (s_hi < 0) s_hi = 0;
    }
    lo = est + s_lo - wdr;
    if (lo < 0) lo = 0;
    hi = est + s_hi + wdr;
}

// rows per lane of the packed kernel (32*R >= n1, every R in 2..16 is instantiated); 0 = too long for it
#define STRK_PK_RMAX 16
#define STRK_PK_NBIN (STRK_PK_RMAX + 1)  // work classes of a batch: [0] general kernel only, [R] packed kernel with R rows per lane
__host__ __device__ inline int strk_pick_rows_packed(int n1) {
    int r = (n1 + 31) / 32;
    if (r < 2) r = 2;
    return r <= STRK_PK_RMAX ? r : 0;
}

__host__ __device__ inline int strk_pick_rows(int n1, int warps = 4) {
    // rows per lane of the strip layout: smallest R in the instantiated set with 32*R >= n1
    const int set[12] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16};
    for (int k = 0; k < 12; ++k)
        if (32 * set[k] >= n1) return set[k];
    // Several strips (pipelined over the warps of a CTA): R = 16.  A cost model that trades rounds of the pipeline
    // against the step length (e.g. 11 strips of 384 rows instead of 9 of 512 for a 4 140-row window) was measured on
    // config 4 and lost 8 %: every additional strip adds its own hand-over lag and synchronisation.
    // (Also measured, for the latency shape of api.cu -- more warps per CTA when a launch holds few reads: the smallest R
    // with 32 * R * warps >= n1, all strips of a read in one round.  Slower at 12 and 16 warps than R = 16: every strip
    // adds ~20 instructions of bookkeeping per step, and one read is bound by the issue slots of its SM.)
    (void)warps;
    return 16;
}
