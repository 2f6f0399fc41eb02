// dp_general.cuh -- general semi-global DP kernel (int32 lanes, any length, any alphabet).
//
// One warp per alignment family.  The rows of the DP matrix (db = fl + tr + fr, the sequence the
// reference profiles, repeats.py:91-93) are cut into strips: lane t owns R consecutive rows, the
// warp sweeps the columns (the candidate) as a skewed wavefront -- at step s lane t computes
// column s - t + 1 -- and the bottom cell of lane t-1 reaches lane t through __shfl_up_sync.
// Families longer than 32*R rows take several passes; the boundary row travels through a per-warp
// scratch row in global memory (L2-resident).
//
// The candidate fl + motif*n + fr is never materialised (reference builds it with
// f"{fl}{motif * i}{fr}"): columns are generated on the fly from the left flank and the motif
// period.  All candidate sizes of a window share ONE forward sweep: because column |fl| + m*n of
// the sweep against fl + motif*n_hi is the state after fl + motif*n, and the remaining |fr|
// columns of any candidate only depend on db's suffix, score(n) = max_i F[i][|fl|+m*n] + B[i],
// where B is one backward sweep of reverse(fr) against reverse(db).  Max-plus path
// decomposition: exact, not a heuristic.
//
// Rows are FRONT-padded to a multiple of 32*R with a pad symbol whose score reproduces the
// border row (0 everywhere when the top border is free, -g*j otherwise), so the last real row is
// always the last register of lane 31 and the init row needs no special case.
//
// Linear gaps only: the reference passes open = extend = 5 (repeats.py:33,40), so parasail's
// affine recurrence collapses to H = max(diag + s, max(up, left) - g).
#pragma once
#include "strk_common.cuh"

#define STRK_NEG_INF (-(1 << 29))

enum PassKind { PASS_DUMP = 0, PASS_COMBINE = 1, PASS_ARGMAX = 2 };

struct PassCfg {
    const unsigned char *s1;   // rows (db), n1 symbols
    const unsigned char *pre;  // column prefix
    const unsigned char *motif;
    int n1, n_pre, m;
    int rev_s1, rev_pre, rev_motif;
    int n_lo, n_hi;       // candidate sizes (COMBINE / ARGMAX)
    int ncols;            // n_pre + m * n_hi  (DUMP: n_pre)
    int s1_beg_free;      // column-0 border is 0
    int s1_last_special;  // backward sweep of a free s1 end: row n1 is not a start node
    int s2_beg_free;      // top border is 0
    int kind;
    int lastrow_term;     // COMBINE with a free s2 end: candidates may stop early on the last row
    int *B;               // [n1 + 2] DUMP writes, COMBINE reads; B[n1 + 1] = best score of a path that
                          // starts inside the suffix columns (free s2 begin), see dp_pass
    int *out;             // COMBINE: [n_hi - n_lo + 1], pre-set to STRK_NEG_INF
    long long *out64;     // ARGMAX: packed (score << 32 | 0x7fffffff - row), pre-set to LLONG_MIN
    int *row0, *row1;     // boundary-row scratch, ncols + 1 ints each (multi-pass families only)
};

struct SmemConsts {
    unsigned char lut[256];
    signed char smat[STRK_SMAT_ROWS * STRK_NSYM_];
};

__device__ __forceinline__ int border_col0(const PassCfg &c, int i, int g) {
    // H[i][0] for real row i (i <= 0: pad rows and the corner)
    if (i <= 0) return 0;
    if (c.s1_beg_free) return (c.s1_last_special && i == c.n1) ? -g : 0;
    return -g * i;
}
__device__ __forceinline__ int border_row0(const PassCfg &c, int j, int g) { return c.s2_beg_free ? 0 : -g * j; }

template <int R>
__device__ void dp_pass(const PassCfg &c, const SmemConsts &sc, int g) {
    const int lane = threadIdx.x & 31;
    const int RB = 32 * R;
    const int NB = c.n1 <= RB ? 1 : (c.n1 + RB - 1) / RB;
    const int off = NB * RB - c.n1;  // number of pad rows in front
    const int padcode = c.s2_beg_free ? STRK_PAD_FREE : STRK_PAD_PEN;
    const int W = c.n_hi - c.n_lo + 1;

    if (c.kind == PASS_DUMP && c.ncols == 0) {  // no columns: the last column is the border
        for (int i = lane; i <= c.n1; i += 32) c.B[i] = border_col0(c, i, g);
        if (lane == 0) c.B[c.n1 + 1] = border_col0(c, c.n1, g);
        __syncwarp();
        return;
    }
    // candidate column 0 (empty left flank and n = 0): scores come from the border column
    if (c.kind == PASS_COMBINE && c.n_pre == 0 && c.n_lo == 0) {
        int best = STRK_NEG_INF;
        for (int i = lane; i <= c.n1; i += 32) {
            int b = c.B[c.n1 - i];
            // node (n1, 0) is not an end node: with a free s2 end the path must take one more column
            if (i == c.n1 && c.lastrow_term) b = -g;
            best = max(best, border_col0(c, i, g) + b);
        }
        if (c.s2_beg_free && lane == 0) best = max(best, c.B[c.n1 + 1]);
        atomicMax(&c.out[0], best);
    }

    for (int b = 0; b < NB; ++b) {
        const int Ibase = b * RB + lane * R;  // padded row index of the row above my strip
        int sym[R], H[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = Ibase + r + 1 - off;  // real row 1..n1, <= 0 for pad rows
            int code = padcode;
            if (i >= 1) {
                int idx = c.rev_s1 ? c.n1 - i : i - 1;
                code = sc.lut[c.s1[idx]];
            }
            sym[r] = code * STRK_NSYM_;
            H[r] = border_col0(c, i, g);
        }
        int prev_up = border_col0(c, Ibase - off, g);  // H[row above][0]
        const int *top = b == 0 ? nullptr : ((b - 1) & 1 ? c.row1 : c.row0);
        int *bot = b < NB - 1 ? (b & 1 ? c.row1 : c.row0) : nullptr;
        int top_next = 0;
        if (lane == 0 && top) top_next = top[1];
        int kk = -1, ncop = 0;
        int pmax = STRK_NEG_INF;  // running max of the last row (lane 31, last pass)
        const int nsteps = c.ncols + 31;
        for (int s = 0; s < nsteps; ++s) {
            const int j = s - lane + 1;
            int up_in = __shfl_up_sync(0xffffffffu, H[R - 1], 1);
            const bool active = j >= 1 && j <= c.ncols;
            if (lane == 0) {
                if (top) {
                    up_in = top_next;
                    if (j + 1 <= c.ncols) top_next = top[j + 1];
                } else {
                    up_in = border_row0(c, j, g);
                }
            }
            if (!active) continue;
            // column symbol
            int colsym;
            if (j <= c.n_pre) {
                colsym = sc.lut[c.pre[c.rev_pre ? c.n_pre - j : j - 1]];
            } else {
                kk = kk + 1 == c.m ? 0 : kk + 1;
                colsym = sc.lut[c.motif[c.rev_motif ? c.m - 1 - kk : kk]];
            }
            const signed char *srow = sc.smat + colsym;
            int d = prev_up, u = up_in;
            prev_up = up_in;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int left = H[r];
                int h = max(d + (int)srow[sym[r]], max(u, left) - g);
                d = left;
                u = h;
                H[r] = h;
            }
            if (bot && lane == 31) bot[j] = H[R - 1];
            pmax = max(pmax, H[R - 1]);

            // candidate column?
            int n = -1;
            if (c.kind == PASS_DUMP) {
                if (j == c.ncols) n = 0;
            } else if (j == c.n_pre) {
                n = 0;
            } else if (j > c.n_pre && kk == c.m - 1) {
                n = ++ncop;
            }
            if (n < 0) continue;
            if (c.kind == PASS_DUMP) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int i = Ibase + r + 1 - off;
                    if (i >= 0) c.B[i] = H[r];
                }
                if (off == 0 && b == 0 && lane == 0) c.B[0] = border_row0(c, j, g);
            } else if (n >= c.n_lo && n <= c.n_hi) {
                if (c.kind == PASS_COMBINE) {
                    int best = STRK_NEG_INF;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int i = Ibase + r + 1 - off;
                        if (i >= 0) best = max(best, H[r] + c.B[c.n1 - i]);
                    }
                    if (off == 0 && b == 0 && lane == 0) best = max(best, border_row0(c, j, g) + c.B[c.n1]);
                    if (c.lastrow_term && b == NB - 1 && lane == 31) best = max(best, pmax);
                    // free s2 begin: a path may start on the top border to the right of this column
                    if (c.s2_beg_free && b == NB - 1 && lane == 31) best = max(best, c.B[c.n1 + 1]);
                    atomicMax(&c.out[n - c.n_lo], best);
                } else {  // ARGMAX over real rows i >= 1, smallest row on ties (parasail end_query = i - 1)
                    long long best = (long long)0x8000000000000000ull;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int i = Ibase + r + 1 - off;
                        if (i >= 1) {
                            long long key = ((long long)H[r] << 32) | (unsigned)(0x7fffffff - i);
                            best = key > best ? key : best;
                        }
                    }
                    atomicMax(&c.out64[n - c.n_lo], best);
                }
            }
        }
        // backward sweep: best value anywhere on its last row (= forward row 0, i.e. paths that skip the
        // whole prefix through a free s2 begin), including the border cell
        if (c.kind == PASS_DUMP && b == NB - 1 && lane == 31) c.B[c.n1 + 1] = max(pmax, border_col0(c, c.n1, g));
        __syncwarp();  // boundary row written by lane 31 is read by lane 0 in the next pass
    }
    (void)W;
}

// ---------------------------------------------------------------------------------------------
// Read family: B = backward sweep of reverse(fr) vs reverse(db); then the forward sweep of
// fl + motif*n_hi vs db combines at every candidate column.  `flags` = STRK_*_FREE end flags.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ void process_read_family(const FamDesc &f, const unsigned char *arena, const SmemConsts &sc, int g,
                                    int flags, int *table, int *scratch, int scratch_rowlen, int scratch_b_len) {
    const int lane = threadIdx.x & 31;
    const int n1 = f.n_fl + f.n_tr + f.n_fr;
    const int W = f.n_hi - f.n_lo + 1;
    int *out = table + f.out_off;
    for (int k = lane; k < W; k += 32) out[k] = STRK_NEG_INF;
    __syncwarp();

    PassCfg c;
    c.s1 = arena + f.db_off;
    c.n1 = n1;
    c.motif = arena + f.motif_off;
    c.m = f.m;
    c.B = scratch;
    c.row0 = scratch + scratch_b_len;
    c.row1 = c.row0 + scratch_rowlen;
    c.out = out;
    c.out64 = nullptr;
    c.n_lo = f.n_lo;
    c.n_hi = f.n_hi;

    // backward sweep: rows reverse(db), columns reverse(fr); its begin flags are the end flags
    c.pre = arena + f.db_off + f.n_fl + f.n_tr;
    c.n_pre = f.n_fr;
    c.rev_s1 = 1;
    c.rev_pre = 1;
    c.rev_motif = 0;
    c.ncols = f.n_fr;
    c.s1_beg_free = (flags & 2) != 0;
    c.s1_last_special = 1;
    c.s2_beg_free = (flags & 8) != 0;
    c.kind = PASS_DUMP;
    c.lastrow_term = 0;
    dp_pass<R>(c, sc, g);
    __syncwarp();

    // forward sweep
    c.pre = arena + f.db_off;
    c.n_pre = f.n_fl;
    c.rev_s1 = 0;
    c.rev_pre = 0;
    c.ncols = f.n_fl + f.m * f.n_hi;
    c.s1_beg_free = (flags & 1) != 0;
    c.s1_last_special = 0;
    c.s2_beg_free = (flags & 4) != 0;
    c.kind = PASS_COMBINE;
    c.lastrow_term = (flags & 8) != 0;
    dp_pass<R>(c, sc, g);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Reference-boundary family (score_ref_boundaries, repeats.py:23-43): two sg_qe sweeps.
//   fwd: columns fl + motif*n           rows db            -> out64[2*k]
//   rev: columns reverse(fr) + reverse(motif)*n   rows reverse(db)   -> out64[2*k + 1]
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ void process_ref_family(const FamDesc &f, const unsigned char *arena, const SmemConsts &sc, int g,
                                   long long *table, int *scratch, int scratch_rowlen, int scratch_b_len) {
    const int lane = threadIdx.x & 31;
    const int n1 = f.n_fl + f.n_tr + f.n_fr;
    const int W = f.n_hi - f.n_lo + 1;
    long long *out = table + 2 * f.out_off;
    for (int k = lane; k < 2 * W; k += 32) out[k] = (long long)0x8000000000000000ull;
    __syncwarp();

    PassCfg c;
    c.s1 = arena + f.db_off;
    c.n1 = n1;
    c.motif = arena + f.motif_off;
    c.m = f.m;
    c.B = nullptr;
    c.row0 = scratch + scratch_b_len;
    c.row1 = c.row0 + scratch_rowlen;
    c.out = nullptr;
    c.n_lo = f.n_lo;
    c.n_hi = f.n_hi;
    c.s1_beg_free = 0;  // sg_qe: both begins penalised, end of s1 free
    c.s1_last_special = 0;
    c.s2_beg_free = 0;
    c.kind = PASS_ARGMAX;
    c.lastrow_term = 0;

    c.pre = arena + f.db_off;
    c.n_pre = f.n_fl;
    c.rev_s1 = 0;
    c.rev_pre = 0;
    c.rev_motif = 0;
    c.ncols = f.n_fl + f.m * f.n_hi;
    c.out64 = out;
    dp_pass<R>(c, sc, g);
    __syncwarp();

    c.pre = arena + f.db_off + f.n_fl + f.n_tr;
    c.n_pre = f.n_fr;
    c.rev_s1 = 1;
    c.rev_pre = 1;
    c.rev_motif = 1;
    c.ncols = f.n_fr + f.m * f.n_hi;
    c.out64 = out + W;
    dp_pass<R>(c, sc, g);
    __syncwarp();
}

// Persistent kernel: warps pull families from a cost-sorted queue.
template <bool REF>
__global__ void __launch_bounds__(256) dp_general_kernel(const FamDesc *__restrict__ fams, const int *__restrict__ order,
                                                         int n_fams, const unsigned char *__restrict__ arena,
                                                         const ScoreConsts *__restrict__ consts, void *table,
                                                         int *scratch, int scratch_rowlen, int scratch_b_len,
                                                         unsigned int *queue,
                                                         const unsigned int *__restrict__ n_fams_dev) {
    if (n_fams_dev) n_fams = (int)*n_fams_dev;  // list length produced on the device (packed-kernel fallbacks)
    __shared__ SmemConsts sc;
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sc.lut[k] = consts->lut[k];
    for (int k = threadIdx.x; k < STRK_SMAT_ROWS * STRK_NSYM_; k += blockDim.x) sc.smat[k] = consts->smat[k];
    __syncthreads();
    const int g = consts->gap;
    const int flags = consts->end_flags;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int *my_scratch = scratch + (size_t)warp_global * (size_t)(scratch_b_len + 2 * scratch_rowlen);

    for (;;) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= (unsigned)n_fams) break;
        const FamDesc f = fams[order ? order[q] : (int)q];
        const int n1 = f.n_fl + f.n_tr + f.n_fr;
        const int R = strk_pick_rows(n1);
#define STRK_CASE(RR)                                                                                          \
    case RR:                                                                                                   \
        if (REF)                                                                                               \
            process_ref_family<RR>(f, arena, sc, g, (long long *)table, my_scratch, scratch_rowlen,            \
                                   scratch_b_len);                                                             \
        else                                                                                                   \
            process_read_family<RR>(f, arena, sc, g, flags, (int *)table, my_scratch, scratch_rowlen,          \
                                    scratch_b_len);                                                            \
        break;
        switch (R) {
            STRK_CASE(2)
            STRK_CASE(3)
            STRK_CASE(4)
            STRK_CASE(5)
            STRK_CASE(6)
            STRK_CASE(7)
            STRK_CASE(8)
            STRK_CASE(9)
            STRK_CASE(10)
            STRK_CASE(12)
            STRK_CASE(14)
            default:
                STRK_CASE(16)
        }
#undef STRK_CASE
    }
}
