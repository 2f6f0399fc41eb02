// dp_packed.cuh -- packed u16x2 DPX kernel for the bulk of the reads (db <= 512 bases).
//
// Same strip wavefront as dp_general.cuh (lane t owns R consecutive rows of db, columns swept with a
// one-column skew per lane, bottom cell handed down with __shfl_up_sync), but every 32-bit lane word
// carries TWO sweeps of the same read at once:
//     low  half: forward sweep   rows db            columns fl + motif*a          (a <= a_hi)
//     high half: backward sweep  rows reverse(db)   columns reverse(fr) + reverse(motif)*b0
// and score(n = a + b0) = max_i F[i][|fl| + m*a] + B[n1 - i][|fr| + m*b0]: the candidate is split in the
// MIDDLE of its repeat tract, so both halves sweep about half of the candidate and one pass of
// (|cand| / 2 + 31) steps scores the whole window of sizes.  Exact max-plus path decomposition.
//
// Cell update in 2 integer instructions.  With H' = H + g*(row + col) the linear-gap recurrence
//     H = max(diag + s, max(up, left) - g)      becomes      H' = max3(diag' + (s + 2g), up', left')
// i.e. one add (IADD3 / IMAD.IADD) and one VIMNMX3.U16x2; every H' is >= 0, so unsigned 16-bit lanes never
// underflow and a plain 32-bit add cannot carry between the halves.
//
// Substitution scores without a per-cell table walk:
//   * flank phase (columns still inside fl / fr): PRMT as an 8-entry byte LUT.  The 8-byte table comes
//     from the column symbol (one LDS.128 per step), the selector from the row symbol (fixed register),
//     the upper byte of each half is produced by PRMT's sign-replicate mode.  2 PRMT + IADD3 + VIMNMX3.
//   * motif phase (both halves inside the periodic tract): the pair of column symbols repeats with
//     period m, so a packed query profile prof[k][row] is built once per read in shared memory and the
//     step costs LDS + IADD + VIMNMX3 per cell pair.  The candidate is never materialised.
//
// Rows are front-padded to 32*R with a pad class whose score reproduces the border row (see
// dp_general.cuh), so row n1 is always the last register of the last lane.
//
// Reads this kernel cannot take (IUPAC codes inside the read, value range beyond u16, empty flank) are
// appended to a fallback list that the general int32 kernel processes afterwards -- still on the GPU.
#pragma once
#include "dp_general.cuh"
#include "strk_common.cuh"

#define PK_FLANK_MAX 160  // longest flank the packed kernel stages (reference default flank_size = 70)

struct PackedSmemDims {
    int colt_entries;  // uint4 entries of the per-column table  (>= max flank + 32)
    int prof_words;    // u32 words of the packed profile        (>= m_max * R * 32)
    int w_max;         // candidate sizes per read               (table row stride)
};

__device__ __forceinline__ unsigned pk_vimax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }

// Raw PRMT (generic mode).  NOT __byte_perm: that intrinsic masks the selector with 0x7777, which costs an
// extra LOP and drops bit 3 of each nibble -- the sign-replicate bit used here to produce the zero bytes.
__device__ __forceinline__ unsigned pk_prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

template <int R>
__global__ void __launch_bounds__(128)
dp_packed_kernel(const FamDesc *__restrict__ fams, const int *__restrict__ list, int n_list,
                 const unsigned char *__restrict__ arena, const ScoreConsts *__restrict__ consts,
                 int *__restrict__ table, PackedSmemDims dims, int *__restrict__ fallback_list,
                 unsigned int *__restrict__ fallback_count) {
    static_assert(R % 2 == 0 && R >= 2 && R <= 16, "R must be even");
    constexpr int N = 32 * R;
    extern __shared__ uint4 smem_raw[];
    __shared__ SmemConsts sc;
    __shared__ unsigned long long t8f[STRK_NSYM_], t8b[STRK_NSYM_];
    __shared__ unsigned char cls_of[STRK_SMAT_ROWS + 1];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sc.lut[k] = consts->lut[k];
    for (int k = threadIdx.x; k < STRK_SMAT_ROWS * STRK_NSYM_; k += blockDim.x) sc.smat[k] = consts->smat[k];
    for (int k = threadIdx.x; k < STRK_NSYM_; k += blockDim.x) {
        t8f[k] = consts->t8f[k];
        t8b[k] = consts->t8b[k];
    }
    for (int k = threadIdx.x; k < STRK_SMAT_ROWS; k += blockDim.x) cls_of[k] = consts->cls_of[k];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fam_idx = blockIdx.x * (blockDim.x >> 5) + warp;
    if (fam_idx >= n_list) return;
    const int fam_id = list[fam_idx];
    const FamDesc f = fams[fam_id];
    const int g = consts->gap;
    const int flags = consts->end_flags;
    const bool s1_beg = flags & 1, s1_end = flags & 2, s2_beg = flags & 4, s2_end = flags & 8;

    // per-warp shared memory carve-up (16-byte units)
    const int fcol_words = dims.w_max * (R / 2) * 32;
    const int bq_words = (R / 2) * 32;
    const int per_warp16 = dims.colt_entries + (dims.prof_words + fcol_words + bq_words + 2 * dims.w_max + 3) / 4 + 1;
    uint4 *colT = smem_raw + (size_t)warp * per_warp16;
    unsigned *prof = (unsigned *)(colT + dims.colt_entries);
    unsigned *fcols = prof + dims.prof_words;
    unsigned *bq = fcols + fcol_words;
    unsigned *flast = bq + bq_words;  // [w_max] last-row prefix maxima at the candidate columns

    const int n1 = f.n_fl + f.n_tr + f.n_fr;
    const int off = N - n1;
    const int m = f.m;
    const unsigned char *db = arena + f.db_off;
    const unsigned char *motif = arena + f.motif_off;

    // split of the tract: b0 copies go to the backward half
    int b0 = (m * f.n_hi + f.n_fl - f.n_fr + m) / (2 * m);
    b0 = b0 < 0 ? 0 : (b0 > f.n_lo ? f.n_lo : b0);
    const int a_lo = f.n_lo - b0, a_hi = f.n_hi - b0;
    const int nW = a_hi - a_lo + 1;
    const int colsF = f.n_fl + m * a_hi, colsB = f.n_fr + m * b0;
    const int ncols = colsF > colsB ? colsF : colsB;
    const int Lmax = f.n_fl > f.n_fr ? f.n_fl : f.n_fr;

    // ---- eligibility (warp-uniform): anything odd goes to the general kernel
    bool ok = n1 <= N && f.n_fl >= 1 && f.n_fr >= 1 && Lmax <= PK_FLANK_MAX && Lmax + 32 <= dims.colt_entries &&
              m * R * 32 <= dims.prof_words && nW <= dims.w_max && (g * (N + ncols + 2) + 2 * N + 256) < 65535;
    // row symbols
    int codeFB[R];  // forward code | backward code << 8   (setup only)
    unsigned selF[R], selB[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = lane * R + r + 1 - off;  // real row (1..n1), <= 0: pad
        int cf = s2_beg ? STRK_PAD_FREE : STRK_PAD_PEN, cb = s2_end ? STRK_PAD_FREE : STRK_PAD_PEN;
        if (i >= 1 && i <= n1) {
            cf = sc.lut[db[i - 1]];
            cb = sc.lut[db[n1 - i]];
        }
        codeFB[r] = cf | (cb << 8);
        const unsigned kf = cls_of[cf], kb = cls_of[cb];
        if ((kf | kb) & 0x80) ok = false;
        selF[r] = (kf & 7) | 0x8880u;
        selB[r] = ((kb & 7) << 8) | 0x8088u;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) {
        if (lane == 0) fallback_list[atomicAdd(fallback_count, 1u)] = fam_id;
        return;
    }

    // ---- per-column PRMT tables for the flank phase (columns 1 .. Lmax + 31)
    const int ct_n = Lmax + 31 < ncols ? Lmax + 31 : ncols;
    for (int j = lane + 1; j <= ct_n; j += 32) {
        int sf, sb;
        if (j <= f.n_fl)
            sf = sc.lut[db[j - 1]];
        else
            sf = sc.lut[motif[(j - f.n_fl - 1) % m]];
        if (j <= f.n_fr)
            sb = sc.lut[db[n1 - j]];
        else
            sb = sc.lut[motif[m - 1 - (j - f.n_fr - 1) % m]];
        const unsigned long long a = t8f[sf], b = t8b[sb];
        colT[j] = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
    }
    // ---- packed profile for the motif phase: prof[(k * R + r) * 32 + lane], column j = Lmax + 1 + k (mod m)
    const int g2 = 2 * g;
    for (int k = 0; k < m; ++k) {
        const int sf = sc.lut[motif[(k + Lmax - f.n_fl) % m]];
        const int sb = sc.lut[motif[m - 1 - (k + Lmax - f.n_fr) % m]];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int vf = sc.smat[(codeFB[r] & 0xff) * STRK_NSYM_ + sf] + g2;
            const int vb = sc.smat[(codeFB[r] >> 8) * STRK_NSYM_ + sb] + g2;
            prof[(k * R + r) * 32 + lane] = (unsigned)vf | ((unsigned)vb << 16);
        }
    }
    for (int k = lane; k < bq_words; k += 32) bq[k] = 0u;

    // ---- borders (biased by g * (row + col))
    unsigned H[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int I = lane * R + r + 1, i = I - off;
        int bf = 0, bb = 0;  // unbiased column-0 values
        if (i >= 1) {
            bf = s1_beg ? 0 : -g * i;
            bb = s1_end ? (i == n1 ? -g : 0) : -g * i;
        }
        H[r] = (unsigned)(bf + g * I) | ((unsigned)(bb + g * I) << 16);
    }
    unsigned prev_up;
    {
        const int I = lane * R, i = I - off;
        int bf = 0, bb = 0;
        if (i >= 1) {
            bf = s1_beg ? 0 : -g * i;
            bb = s1_end ? 0 : -g * i;  // i < n1 here
        }
        prev_up = (unsigned)(bf + g * I) | ((unsigned)(bb + g * I) << 16);
    }
    const unsigned tinc = (s2_beg ? (unsigned)g : 0u) | ((s2_end ? (unsigned)g : 0u) << 16);
    const unsigned ginc = (unsigned)g | ((unsigned)g << 16);
    unsigned topv = 0u;  // top border at the column lane 0 is about to compute
    unsigned pm = 0u;    // biased prefix maxima of the last row (both halves), lane 31
    int next_cand = f.n_fl + m * a_lo, w = 0;
    unsigned bextra = 0u, bq_top = 0u;
    __syncwarp();

    const int nsteps = ncols + 31;
    const int s_star = Lmax + 31 < nsteps ? Lmax + 31 : nsteps;  // first step of the motif phase (warp-uniform)

    // candidate-column / final-column bookkeeping shared by both phases
#define PK_STEP_TAIL()                                                                                         \
    pm = (j == 1) ? H[R - 1] : __viaddmax_u16x2(pm, ginc, H[R - 1]);                                           \
    if (j == next_cand) {                                                                                      \
        _Pragma("unroll") for (int q = 0; q < R / 2; ++q)                                                      \
            fcols[(w * (R / 2) + q) * 32 + lane] = __byte_perm(H[2 * q], H[2 * q + 1], 0x5410);                \
        if (lane == 31) flast[w] = pm & 0xffffu;                                                               \
        ++w;                                                                                                   \
        next_cand = w < nW ? next_cand + m : 0x7fffffff;                                                       \
    }                                                                                                          \
    if (j == colsB) {                                                                                          \
        _Pragma("unroll") for (int r = 0; r < R; ++r) {                                                        \
            const int Ib = lane * R + r + 1;                                                                   \
            const int If = N + off - Ib; /* forward row paired with this backward row */                       \
            if (Ib >= off && If >= 1) {                                                                        \
                const int lf = (If - 1) / R, rf = (If - 1) % R;                                                \
                ((unsigned short *)bq)[((rf >> 1) * 32 + lf) * 2 + (rf & 1)] = (unsigned short)(H[r] >> 16);   \
            }                                                                                                  \
        }                                                                                                      \
        if (lane == 31) {                                                                                      \
            bextra = pm >> 16;                                                                                 \
            bq_top = H[R - 1] >> 16;                                                                           \
        }                                                                                                      \
    }

    int s = 0;
    // ---- flank phase: PRMT look-ups
    for (; s < s_star; ++s) {
        const int j = s - lane + 1;
        unsigned up_in = __shfl_up_sync(0xffffffffu, H[R - 1], 1);
        if (lane == 0) {
            topv += tinc;
            up_in = topv;
        }
        if (j >= 1 && j <= ncols) {
            const uint4 ct = colT[j];
            unsigned d = prev_up, u = up_in;
            prev_up = up_in;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned left = H[r];
                const unsigned t = d + pk_prmt(ct.x, ct.y, selF[r]) + pk_prmt(ct.z, ct.w, selB[r]);
                const unsigned h = pk_vimax3(t, u, left);
                d = left;
                u = h;
                H[r] = h;
            }
            PK_STEP_TAIL()
        }
    }
    // ---- motif phase: packed profile from shared memory
    int k = 0;
    {
        const int j = s - lane + 1;  // >= Lmax + 1 for every lane here
        k = (j - Lmax - 1) % m;
        if (k < 0) k += m;
    }
    for (; s < nsteps; ++s) {
        const int j = s - lane + 1;
        unsigned up_in = __shfl_up_sync(0xffffffffu, H[R - 1], 1);
        if (lane == 0) {
            topv += tinc;
            up_in = topv;
        }
        if (j <= ncols) {
            const unsigned *pp = prof + k * (R * 32) + lane;
            unsigned d = prev_up, u = up_in;
            prev_up = up_in;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned left = H[r];
                const unsigned t = d + pp[r * 32];
                const unsigned h = pk_vimax3(t, u, left);
                d = left;
                u = h;
                H[r] = h;
            }
            PK_STEP_TAIL()
        }
        k = k + 1 == m ? 0 : k + 1;
    }
#undef PK_STEP_TAIL

    // ---- combine: score(n) for every candidate of the window
    bextra = __shfl_sync(0xffffffffu, bextra, 31);
    bq_top = __shfl_sync(0xffffffffu, bq_top, 31);
    __syncwarp();
    int *out = table + f.out_off;
    for (int ww = 0; ww < nW; ++ww) {
        unsigned acc = 0u;
#pragma unroll
        for (int q = 0; q < R / 2; ++q)
            acc = __viaddmax_u16x2(fcols[(ww * (R / 2) + q) * 32 + lane], bq[q * 32 + lane], acc);
        unsigned v = max(acc & 0xffffu, acc >> 16);
        v = __reduce_max_sync(0xffffffffu, v);
        if (lane == 0) {
            const int p = f.n_fl + m * (a_lo + ww);
            int best = (int)v - g * (N + off + p + colsB);
            if (off == 0) {  // forward row 0 is the top border, not a stored row
                const int f0 = s2_beg ? 0 : -g * p;
                best = max(best, f0 + (int)bq_top - g * (N + colsB));
            }
            if (s2_end) best = max(best, (int)flast[ww] - g * (N + p));
            if (s2_beg) {
                const int border = s1_end ? -g : -g * n1;  // backward cell (n1, 0)
                best = max(best, max((int)bextra - g * (N + colsB), border));
            }
            out[a_lo + ww + b0 - f.n_lo] = best;
        }
    }
}
