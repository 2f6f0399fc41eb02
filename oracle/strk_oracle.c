/*
 * strk_oracle.c -- CPU restatement of STRkit's repeat-count hot path.
 * TEST INFRASTRUCTURE ONLY (see strk_oracle.h header for the rules and the
 * parity status: "parity unpinned" for parasail / strkit_rust_ext internals).
 *
 * Plain C, int32 arithmetic, no SIMD, one full alignment per candidate size --
 * exactly the work the reference does (no prefix sharing), so that timing this
 * file is a fair port of the reference's CPU cost model.
 */
#include "strk_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (INT_MIN / 2)

/* ---------------------------------------------------------------------------------------------
 * Alphabet and matrix: strkit/call/align_matrix.py:25-44, strkit/iupac.py:9-21.
 * dna_bases_str = "ACGT" + "RYSWKMBDHVN" (dict order of IUPAC_NUCLEOTIDE_CODES) + "X".
 * parasail.matrix_create(alphabet, match, mismatch): case-insensitive mapper, one
 * extra wildcard row/column scoring 0 for any byte outside the alphabet.
 * ------------------------------------------------------------------------------------------- */
static const char ALPHABET[17] = "ACGTRYSWKMBDHVNX";

int strk_oracle_symbol(unsigned char c) {
    if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 'a' + 'A');
    for (int i = 0; i < 16; ++i)
        if ((unsigned char)ALPHABET[i] == c) return i;
    return 16;
}

void strk_oracle_dna_matrix(int8_t out[STRK_NSYM * STRK_NSYM]) {
    /* matrix_create("ACGT...X", 2, -7): align_matrix.py:15-17,34 */
    for (int i = 0; i < STRK_NSYM; ++i)
        for (int j = 0; j < STRK_NSYM; ++j)
            out[i * STRK_NSYM + j] = (i == 16 || j == 16) ? 0 : (i == j ? 2 : -7);
    /* iupac.py:9-21 -- note "D" lists (A, C, T), the same as "H" (quirk preserved). */
    static const struct {
        char code;
        const char *bases;
    } codes[] = {{'R', "AG"}, {'Y', "CT"}, {'S', "CG"}, {'W', "AT"},   {'K', "GT"},   {'M', "AC"},
                 {'B', "CGT"}, {'D', "ACT"}, {'H', "ACT"}, {'V', "ACG"}, {'N', "ACGT"}, {'X', "ACGT"}};
    /* align_matrix.py:36-39 */
    for (unsigned k = 0; k < sizeof(codes) / sizeof(codes[0]); ++k) {
        int ci = strk_oracle_symbol((unsigned char)codes[k].code);
        int v = codes[k].code != 'X' ? 2 : 0;
        for (const char *b = codes[k].bases; *b; ++b) {
            int bi = strk_oracle_symbol((unsigned char)*b);
            out[ci * STRK_NSYM + bi] = (int8_t)v;
            out[bi * STRK_NSYM + ci] = (int8_t)v;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Semi-global alignment, parasail "sg" family semantics (restated; library absent).
 * Rows i = 1..n1 over s1 (the profiled sequence), columns j = 1..n2 over s2.
 * Gap of length k costs open + (k-1)*extend.  Free begin => zero border; free end of s1 =>
 * max over the last column (smallest row on ties, end_query = i-1); free end of s2 => max over
 * the last row; the corner always counts.
 * Call sites: repeats.py:33,40 (sg_qe_scan_profile_sat, open = extend = indel_penalty = 5).
 * ------------------------------------------------------------------------------------------- */
int strk_oracle_sg_align(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                         const int8_t *matrix, int flags, int *score, int *end_query, int *end_ref) {
    if (n1 <= 0 || n2 <= 0 || !s1 || !s2 || !matrix) return 1;
    const int s1_beg = flags & STRK_S1_BEG_FREE, s1_end = flags & STRK_S1_END_FREE;
    const int s2_beg = flags & STRK_S2_BEG_FREE, s2_end = flags & STRK_S2_END_FREE;

    int *H = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    int *F = (int *)malloc(sizeof(int) * (size_t)(n2 + 1)); /* vertical gap state per column */
    uint8_t *c2 = (uint8_t *)malloc((size_t)n2);
    if (!H || !F || !c2) {
        free(H);
        free(F);
        free(c2);
        return 2;
    }
    for (int j = 0; j < n2; ++j) c2[j] = (uint8_t)strk_oracle_symbol((unsigned char)s2[j]);

    H[0] = 0;
    F[0] = NEG_INF;
    for (int j = 1; j <= n2; ++j) {
        H[j] = s2_beg ? 0 : -gap_open - (j - 1) * gap_extend;
        F[j] = NEG_INF;
    }

    int best = NEG_INF, bq = n1 - 1, br = n2 - 1;
    for (int i = 1; i <= n1; ++i) {
        const int8_t *mrow = matrix + STRK_NSYM * strk_oracle_symbol((unsigned char)s1[i - 1]);
        int diag = H[0];
        int left = s1_beg ? 0 : -gap_open - (i - 1) * gap_extend; /* H[i][0] */
        int E = NEG_INF;                                          /* horizontal gap state */
        H[0] = left;
        for (int j = 1; j <= n2; ++j) {
            int up = H[j];
            int f = F[j] - gap_extend;
            int fo = up - gap_open;
            if (fo > f) f = fo;
            F[j] = f;
            int e = E - gap_extend;
            int eo = left - gap_open;
            if (eo > e) e = eo;
            E = e;
            int h = diag + mrow[c2[j - 1]];
            if (e > h) h = e;
            if (f > h) h = f;
            diag = up;
            H[j] = h;
            left = h;
        }
        /* free end of s1: last column, strict '>' keeps the smallest row */
        if (s1_end && left > best) {
            best = left;
            bq = i - 1;
            br = n2 - 1;
        }
    }
    if (s2_end) {
        for (int j = 1; j <= n2; ++j)
            if (H[j] > best) {
                best = H[j];
                bq = n1 - 1;
                br = j - 1;
            }
    }
    if (H[n2] > best || (!s1_end && !s2_end)) {
        best = H[n2];
        bq = n1 - 1;
        br = n2 - 1;
    }
    free(H);
    free(F);
    free(c2);
    if (score) *score = best;
    if (end_query) *end_query = bq;
    if (end_ref) *end_ref = br;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * The same alignment vectorised the way parasail's "scan" kernels are (striped query profile built
 * once per profiled sequence, repeats.py:92-93; per column one pass for the diagonal / horizontal
 * terms, a prefix scan down the column for the vertical gap, a second pass to apply it), 16-bit
 * lanes, AVX2.  For the CPU BASELINE only: results are asserted identical to the scalar restatement
 * above (tests/test_oracle_golden.py) and the parity tests keep using the scalar one unless told
 * otherwise.  Linear gaps (open == extend, what the reference passes); sequences too long for
 * 16-bit lanes fall back to the scalar code, as parasail's _sat kernels fall back to wider lanes.
 * ------------------------------------------------------------------------------------------- */
#if defined(__AVX2__)
#include <immintrin.h>
#define STRK_HAVE_AVX2 1
#define P16_MAXLEN_OR_0 6400
#define P16_LANES 16
#define P16_MAXLEN 6400 /* g * (len + 1) must stay inside int16 for g = 5 */

typedef struct {
    int n1, seglen;
    __m256i *prof; /* [17][seglen]: lane l of vector k = score(s1[l * seglen + k], symbol) */
    __m256i *h, *ht; /* column buffers */
} prof16;

static void prof16_free(prof16 *p) {
    if (!p) return;
    free(p->prof);
    free(p->h);
    free(p->ht);
    free(p);
}

static prof16 *prof16_create(const char *s1, int n1, const int8_t *matrix) {
    if (n1 <= 0 || n1 > P16_MAXLEN) return NULL;
    prof16 *p = (prof16 *)calloc(1, sizeof(prof16));
    if (!p) return NULL;
    p->n1 = n1;
    p->seglen = (n1 + P16_LANES - 1) / P16_LANES;
    const size_t sl = (size_t)p->seglen;
    if (posix_memalign((void **)&p->prof, 32, sizeof(__m256i) * sl * STRK_NSYM) ||
        posix_memalign((void **)&p->h, 32, sizeof(__m256i) * sl) || posix_memalign((void **)&p->ht, 32, sizeof(__m256i) * sl)) {
        prof16_free(p);
        return NULL;
    }
    int16_t *w = (int16_t *)p->prof;
    for (int c = 0; c < STRK_NSYM; ++c)
        for (int k = 0; k < p->seglen; ++k)
            for (int l = 0; l < P16_LANES; ++l) {
                int r = l * p->seglen + k;
                w[((size_t)c * sl + (size_t)k) * P16_LANES + l] =
                    r < n1 ? matrix[STRK_NSYM * strk_oracle_symbol((unsigned char)s1[r]) + c] : 0;
            }
    return p;
}

/* lanes move up by 1 / 2 / 4 / 8 positions (lane l receives lane l - s); vacated lanes take `fill` */
static inline __m256i p16_shift1(__m256i v, __m256i fillv) {
    __m256i t = _mm256_permute2x128_si256(v, v, 0x08);
    __m256i r = _mm256_alignr_epi8(v, t, 14);
    return _mm256_blend_epi16(r, fillv, 0x01) /* lane 0 and lane 8 */;
}

static int sg_align_prof16(const prof16 *p, const char *s2, int n2, int g, int flags, int *score, int *end_query,
                           int *end_ref) {
    const int n1 = p->n1, seglen = p->seglen;
    const int s1_beg = flags & STRK_S1_BEG_FREE, s1_end = flags & STRK_S1_END_FREE;
    const int s2_beg = flags & STRK_S2_BEG_FREE, s2_end = flags & STRK_S2_END_FREE;
    __m256i *H = p->h, *Ht = p->ht;
    int16_t *lastrow = s2_end ? (int16_t *)malloc(sizeof(int16_t) * (size_t)(n2 + 1)) : NULL;
    if (s2_end && !lastrow) return 2;
    const int rq = n1 - 1, kq = rq % seglen, lq = rq / seglen; /* position of DP row n1 */
    const __m256i vg = _mm256_set1_epi16((short)g);
    const __m256i vneg = _mm256_set1_epi16(-32768);
    const __m256i vgseg1 = _mm256_set1_epi16((short)(g * seglen > 32767 ? 32767 : g * seglen));
    /* masks that keep lanes >= s after a lane shift by s (the others are refilled with -32768) */
    int16_t lane_id[16];
    for (int l = 0; l < 16; ++l) lane_id[l] = (int16_t)l;
    const __m256i vlane = _mm256_loadu_si256((const __m256i *)lane_id);
    /* column 0 */
    {
        int16_t *h = (int16_t *)H;
        for (int k = 0; k < seglen; ++k)
            for (int l = 0; l < P16_LANES; ++l) {
                int r = l * seglen + k;
                h[k * P16_LANES + l] = (int16_t)(s1_beg ? 0 : -g * (r + 1));
            }
    }
    for (int j = 1; j <= n2; ++j) {
        const __m256i *P = p->prof + (size_t)strk_oracle_symbol((unsigned char)s2[j - 1]) * (size_t)seglen;
        const int top_prev = s2_beg ? 0 : -g * (j - 1); /* H[0][j-1] */
        const int top = s2_beg ? 0 : -g * j;            /* H[0][j]   */
        /* diagonal of vector 0: previous column's last vector moved up one lane, H[0][j-1] into lane 0 */
        __m256i last = H[seglen - 1];
        __m256i t = _mm256_permute2x128_si256(last, last, 0x08);
        __m256i vdiag = _mm256_alignr_epi8(last, t, 14);
        vdiag = _mm256_insert_epi16(vdiag, (short)top_prev, 0);
        /* pass 1: diagonal / horizontal terms, and the vertical gap inside each lane's run of rows */
        __m256i run = vneg;
        for (int k = 0; k < seglen; ++k) {
            __m256i hp = H[k];
            __m256i v = _mm256_max_epi16(_mm256_adds_epi16(vdiag, P[k]), _mm256_subs_epi16(hp, vg));
            run = _mm256_max_epi16(v, _mm256_subs_epi16(run, vg));
            Ht[k] = run;
            vdiag = hp;
        }
        /* carry into each lane: value of the row just above its first row.  A[0] = H[0][j], A[l] = lane l-1's last
         * row (without its own carry); C[l] = max over l' <= l of A[l'] - g * seglen * (l - l') by doubling */
        __m256i t2 = _mm256_permute2x128_si256(run, run, 0x08);
        __m256i A = _mm256_alignr_epi8(run, t2, 14);
        A = _mm256_insert_epi16(A, (short)top, 0);
        __m256i dec = vgseg1;
        /* shift by 1 */
        {
            __m256i tt = _mm256_permute2x128_si256(A, A, 0x08);
            __m256i sh = _mm256_alignr_epi8(A, tt, 14);
            sh = _mm256_blendv_epi8(sh, vneg, _mm256_cmpgt_epi16(_mm256_set1_epi16(1), vlane));
            A = _mm256_max_epi16(A, _mm256_subs_epi16(sh, dec));
            dec = _mm256_adds_epi16(dec, dec);
        }
        {
            __m256i tt = _mm256_permute2x128_si256(A, A, 0x08);
            __m256i sh = _mm256_alignr_epi8(A, tt, 12);
            sh = _mm256_blendv_epi8(sh, vneg, _mm256_cmpgt_epi16(_mm256_set1_epi16(2), vlane));
            A = _mm256_max_epi16(A, _mm256_subs_epi16(sh, dec));
            dec = _mm256_adds_epi16(dec, dec);
        }
        {
            __m256i tt = _mm256_permute2x128_si256(A, A, 0x08);
            __m256i sh = _mm256_alignr_epi8(A, tt, 8);
            sh = _mm256_blendv_epi8(sh, vneg, _mm256_cmpgt_epi16(_mm256_set1_epi16(4), vlane));
            A = _mm256_max_epi16(A, _mm256_subs_epi16(sh, dec));
            dec = _mm256_adds_epi16(dec, dec);
        }
        {
            __m256i sh = _mm256_permute2x128_si256(A, A, 0x08);
            sh = _mm256_blendv_epi8(sh, vneg, _mm256_cmpgt_epi16(_mm256_set1_epi16(8), vlane));
            A = _mm256_max_epi16(A, _mm256_subs_epi16(sh, dec));
        }
        /* pass 2: apply the carry down each lane's rows */
        __m256i c = _mm256_subs_epi16(A, vg);
        for (int k = 0; k < seglen; ++k) {
            H[k] = _mm256_max_epi16(Ht[k], c);
            c = _mm256_subs_epi16(c, vg);
        }
        if (lastrow) lastrow[j] = ((const int16_t *)&H[kq])[lq];
    }
    int best = NEG_INF, bq = n1 - 1, br = n2 - 1;
    const int16_t *h = (const int16_t *)H;
    if (s1_end) {
        for (int r = 0; r < n1; ++r) {
            int v = h[(r % seglen) * P16_LANES + r / seglen];
            if (v > best) best = v, bq = r, br = n2 - 1;
        }
    }
    if (s2_end) {
        for (int j = 1; j <= n2; ++j)
            if (lastrow[j] > best) best = lastrow[j], bq = n1 - 1, br = j - 1;
    }
    {
        int corner = h[kq * P16_LANES + lq];
        if (corner > best || (!s1_end && !s2_end)) best = corner, bq = n1 - 1, br = n2 - 1;
    }
    free(lastrow);
    if (score) *score = best;
    if (end_query) *end_query = bq;
    if (end_ref) *end_ref = br;
    return 0;
}
#else
#define STRK_HAVE_AVX2 0
#define P16_MAXLEN_OR_0 0
typedef struct prof16 prof16;
static void prof16_free(prof16 *p) { (void)p; }
static prof16 *prof16_create(const char *s1, int n1, const int8_t *matrix) {
    (void)s1, (void)n1, (void)matrix;
    return NULL;
}
static int sg_align_prof16(const prof16 *p, const char *s2, int n2, int g, int flags, int *score, int *end_query,
                           int *end_ref) {
    (void)p, (void)s2, (void)n2, (void)g, (void)flags, (void)score, (void)end_query, (void)end_ref;
    return 1;
}
#endif

/* 1 = the batch / search entry points below align through the AVX2 kernel where it applies (CPU baseline);
 * 0 (default) = the scalar restatement everywhere (the checker). */
static int g_use_simd = 0;
int strk_oracle_set_simd(int on) {
    int prev = g_use_simd;
    g_use_simd = on && STRK_HAVE_AVX2;
    return prev;
}
int strk_oracle_have_simd(void) { return STRK_HAVE_AVX2; }

int strk_oracle_sg_align_simd(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                              const int8_t *matrix, int flags, int *score, int *end_query, int *end_ref) {
    if (n1 <= 0 || n2 <= 0 || !s1 || !s2 || !matrix) return 1;
    prof16 *p = (gap_open == gap_extend && n2 <= P16_MAXLEN_OR_0) ? prof16_create(s1, n1, matrix) : NULL;
    if (!p) return strk_oracle_sg_align(s1, n1, s2, n2, gap_open, gap_extend, matrix, flags, score, end_query, end_ref);
    int rc = sg_align_prof16(p, s2, n2, gap_open, flags, score, end_query, end_ref);
    prof16_free(p);
    return rc;
}

/* alignment of s2 against an optional prebuilt profile of s1 (NULL, or s2 too long for 16-bit lanes: scalar) */
static int sg_align_any(const prof16 *p, const char *s1, int n1, const char *s2, int n2, int gap, const int8_t *matrix,
                        int flags, int *score, int *end_query, int *end_ref) {
    if (p && n2 > 0 && n2 <= P16_MAXLEN_OR_0) return sg_align_prof16(p, s2, n2, gap, flags, score, end_query, end_ref);
    return strk_oracle_sg_align(s1, n1, s2, n2, gap, gap, matrix, flags, score, end_query, end_ref);
}

/* ---------------------------------------------------------------------------------------------
 * Candidate construction: f"{flank_left_seq}{motif * n}{flank_right_seq}" scored against the
 * profile of db = fl + tr + fr (the pre-Rust Python body of get_repeat_count; the Rust port is
 * called at repeats.py:58-68).  Mode (free ends) is a parameter: believed plain "sg".
 * ------------------------------------------------------------------------------------------- */
static char *build_candidate(const char *fl, int n_fl, const char *motif, int m, int n, const char *fr, int n_fr,
                             int reverse, int *len_out) {
    int len = n_fl + m * n + n_fr;
    char *c = (char *)malloc((size_t)len + 1);
    if (!c) return NULL;
    int p = 0;
    if (n_fl) memcpy(c + p, fl, (size_t)n_fl);
    p += n_fl;
    for (int k = 0; k < n; ++k, p += m) memcpy(c + p, motif, (size_t)m);
    if (n_fr) memcpy(c + p, fr, (size_t)n_fr);
    if (reverse)
        for (int a = 0, b = len - 1; a < b; ++a, --b) {
            char t = c[a];
            c[a] = c[b];
            c[b] = t;
        }
    *len_out = len;
    return c;
}

static int score_candidate_p(const prof16 *p, const char *db, int n_db, const char *fl, int n_fl, const char *fr,
                             int n_fr, const char *motif, int m, int n, int gap, const int8_t *matrix, int flags,
                             int *score) {
    int len;
    char *cand = build_candidate(fl, n_fl, motif, m, n, fr, n_fr, 0, &len);
    if (!cand) return 2;
    int rc = sg_align_any(p, db, n_db, cand, len, gap, matrix, flags, score, NULL, NULL);
    free(cand);
    return rc;
}

int strk_oracle_score_candidate(const char *db, int n_db, const char *fl, int n_fl, const char *fr, int n_fr,
                                const char *motif, int m, int n, int gap, const int8_t *matrix, int flags,
                                int *score) {
    return score_candidate_p(NULL, db, n_db, fl, n_fl, fr, n_fr, motif, m, n, gap, matrix, flags, score);
}

/* Insertion-ordered int -> value map (Python dict semantics for repeats.py:103-104,154-156). */
typedef struct {
    int *keys;
    int *v0, *v1, *v2, *v3;
    int n, cap;
} omap;

static int omap_find(const omap *mp, int key) {
    for (int i = 0; i < mp->n; ++i)
        if (mp->keys[i] == key) return i;
    return -1;
}
static int omap_push(omap *mp, int key, int a, int b, int c, int d) {
    if (mp->n == mp->cap) {
        int nc = mp->cap ? mp->cap * 2 : 64;
        mp->keys = (int *)realloc(mp->keys, sizeof(int) * (size_t)nc);
        mp->v0 = (int *)realloc(mp->v0, sizeof(int) * (size_t)nc);
        mp->v1 = (int *)realloc(mp->v1, sizeof(int) * (size_t)nc);
        mp->v2 = (int *)realloc(mp->v2, sizeof(int) * (size_t)nc);
        mp->v3 = (int *)realloc(mp->v3, sizeof(int) * (size_t)nc);
        mp->cap = nc;
    }
    mp->keys[mp->n] = key;
    mp->v0[mp->n] = a;
    mp->v1[mp->n] = b;
    mp->v2[mp->n] = c;
    mp->v3[mp->n] = d;
    return mp->n++;
}
static void omap_free(omap *mp) {
    free(mp->keys);
    free(mp->v0);
    free(mp->v1);
    free(mp->v2);
    free(mp->v3);
    memset(mp, 0, sizeof(*mp));
}

typedef struct {
    int size, dir;
} explore_t;

/* ---------------------------------------------------------------------------------------------
 * strkit_rust_ext.get_repeat_count (repeats.py:58-68; Rust body not in tree).  Restated from the
 * in-tree statement of the same search, repeats.py:100-156, with one score per size:
 *   stack [(s-step,-1),(s+step,+1),(s,0)] popped from the end; negative sizes skipped; window
 *   bounds per :114-117; unseen sizes scored and counted :119-130; first maximal element :135;
 *   push rules :136-151; loop guard :106; final pick = first-inserted maximum :154.
 * ------------------------------------------------------------------------------------------- */
static int get_repeat_count_cells(int start_count, const char *db, int n_db, const char *fl, int n_fl,
                                  const char *fr, int n_fr, const char *motif, int m, int max_iters,
                                  int local_search_range, int step_size, int gap, const int8_t *matrix, int flags,
                                  int tie_flags, int32_t out4[4], double *cells) {
    omap seen;
    memset(&seen, 0, sizeof(seen));
    int cap = 16, top = 0;
    explore_t *stack = (explore_t *)malloc(sizeof(explore_t) * (size_t)cap);
    stack[top++] = (explore_t){start_count - step_size, -1};
    stack[top++] = (explore_t){start_count + step_size, 1};
    stack[top++] = (explore_t){start_count, 0};
    int n_explored = 0, rc = 0;
    /* the reference builds the profile of db once per search (profile_create_sat); so does the SIMD baseline */
    prof16 *prof = g_use_simd ? prof16_create(db, n_db, matrix) : NULL;

    while (top > 0 && n_explored < max_iters) {
        explore_t e = stack[--top];
        if (e.size < 0) continue;
        int wide = step_size > local_search_range;
        int start_size = e.size - ((e.dir < 1 || wide) ? local_search_range : 0);
        if (start_size < 0) start_size = 0;
        int end_size = e.size + ((e.dir > -1 || wide) ? local_search_range : 0);
        /* Hypotheses about the Rust body (repeat_count_params.py:13: the range "can be narrowed within the
         * get_repeat_count fn"); 0 = the in-tree statement, where it never changes. */
        if ((tie_flags & STRK_SEARCH_NARROW_FIRST) && local_search_range > 1) local_search_range = 1;
        if ((tie_flags & STRK_SEARCH_NARROW_HALVE) && local_search_range > 1) local_search_range /= 2;

        int have = 0, mv_size = 0, mv_score = 0;
        for (int i = start_size; i <= end_size; ++i) {
            int idx = omap_find(&seen, i);
            if (idx < 0) {
                int sc;
                rc = score_candidate_p(prof, db, n_db, fl, n_fl, fr, n_fr, motif, m, i, gap, matrix, flags, &sc);
                if (rc) goto done;
                if (cells) *cells += (double)n_db * (double)(n_fl + m * i + n_fr);
                idx = omap_push(&seen, i, sc, 0, 0, 0);
                ++n_explored;
            }
            int sc = seen.v0[idx];
            if (!have || sc > mv_score || ((tie_flags & STRK_TIE_WINDOW_LAST) && sc == mv_score)) {
                have = 1;
                mv_size = i;
                mv_score = sc;
            }
        }
        if (top + 2 > cap) {
            cap *= 2;
            stack = (explore_t *)realloc(stack, sizeof(explore_t) * (size_t)cap);
        }
        if (mv_size > e.size) {
            int new_rc = mv_size + step_size;
            if (omap_find(&seen, new_rc) < 0 && new_rc >= 0) stack[top++] = (explore_t){new_rc, 1};
        }
        if (mv_size < e.size) {
            int new_rc = mv_size - step_size;
            if (omap_find(&seen, new_rc) < 0 && new_rc >= 0) stack[top++] = (explore_t){new_rc, -1};
        }
    }
    if (seen.n == 0) { /* max() of an empty dict raises in the reference */
        rc = 3;
        goto done;
    }
    {
        int bi = 0;
        for (int i = 1; i < seen.n; ++i)
            if (seen.v0[i] > seen.v0[bi] || ((tie_flags & STRK_TIE_FINAL_LAST) && seen.v0[i] == seen.v0[bi])) bi = i;
        out4[0] = seen.keys[bi];
        out4[1] = seen.v0[bi];
        out4[2] = n_explored;
        out4[3] = seen.keys[bi] - start_count;
    }
done:
    free(stack);
    omap_free(&seen);
    prof16_free(prof);
    return rc;
}

static char *concat3(const char *a, int na, const char *b, int nb, const char *c, int nc) {
    char *s = (char *)malloc((size_t)(na + nb + nc) + 1);
    if (!s) return NULL;
    if (na) memcpy(s, a, (size_t)na);
    if (nb) memcpy(s + na, b, (size_t)nb);
    if (nc) memcpy(s + na + nb, c, (size_t)nc);
    return s;
}

int strk_oracle_get_repeat_count(int start_count, const char *tr, int n_tr, const char *fl, int n_fl,
                                 const char *fr, int n_fr, const char *motif, int m, int max_iters,
                                 int local_search_range, int step_size, int gap, const int8_t *matrix, int flags,
                                 int tie_flags, int32_t out4[4]) {
    if (m <= 0 || n_tr < 0 || n_fl < 0 || n_fr < 0 || n_fl + n_tr + n_fr <= 0) return 1;
    char *db = concat3(fl, n_fl, tr, n_tr, fr, n_fr);
    if (!db) return 2;
    int rc = get_repeat_count_cells(start_count, db, n_fl + n_tr + n_fr, fl, n_fl, fr, n_fr, motif, m, max_iters,
                                    local_search_range, step_size, gap, matrix, flags, tie_flags, out4, NULL);
    free(db);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * score_ref_boundaries, repeats.py:23-43.
 *   fwd: sg_qe(profile(db), fl + cand)                 -> (score, end_query + 1 - |fl| - ref_size)
 *   rev: sg_qe(profile(db[::-1]), (cand + fr)[::-1])   -> (score, end_query + 1 - |fr| - ref_size)
 * ------------------------------------------------------------------------------------------- */
static int score_ref_boundaries_rev(const prof16 *pf, const prof16 *pr, const char *db, const char *db_rev, int n_db,
                                    const char *fl, int n_fl, const char *fr, int n_fr, const char *motif, int m, int n,
                                    int ref_size, int gap, const int8_t *matrix, int32_t out4[4]) {
    int len, sc, eq, rc;
    char *ext_r = build_candidate(fl, n_fl, motif, m, n, NULL, 0, 0, &len);
    if (!ext_r) return 2;
    rc = sg_align_any(pf, db, n_db, ext_r, len, gap, matrix, STRK_MODE_SG_QE, &sc, &eq, NULL);
    free(ext_r);
    if (rc) return rc;
    out4[0] = sc;
    out4[1] = eq + 1 - n_fl - ref_size;
    char *ext_l = build_candidate(NULL, 0, motif, m, n, fr, n_fr, 1, &len);
    if (!ext_l) return 2;
    rc = sg_align_any(pr, db_rev, n_db, ext_l, len, gap, matrix, STRK_MODE_SG_QE, &sc, &eq, NULL);
    free(ext_l);
    if (rc) return rc;
    out4[2] = sc;
    out4[3] = eq + 1 - n_fr - ref_size;
    return 0;
}

static char *reversed(const char *s, int n) {
    char *r = (char *)malloc((size_t)n + 1);
    if (!r) return NULL;
    for (int i = 0; i < n; ++i) r[i] = s[n - 1 - i];
    return r;
}

int strk_oracle_score_ref_boundaries(const char *db, int n_db, const char *fl, int n_fl, const char *fr, int n_fr,
                                     const char *motif, int m, int n, int ref_size, int gap, const int8_t *matrix,
                                     int32_t out4[4]) {
    char *rev = reversed(db, n_db);
    if (!rev) return 2;
    int rc = score_ref_boundaries_rev(NULL, NULL, db, rev, n_db, fl, n_fl, fr, n_fr, motif, m, n, ref_size, gap, matrix,
                                      out4);
    free(rev);
    return rc;
}

/* Python tuple comparison (score, adj) > (score', adj') */
static int pair_gt(int s0, int a0, int s1, int a1) { return s0 > s1 || (s0 == s1 && a0 > a1); }

/* ---------------------------------------------------------------------------------------------
 * get_ref_repeat_count, repeats.py:73-192.
 * ------------------------------------------------------------------------------------------- */
int strk_oracle_get_ref_repeat_count(int start_count, const char *tr, int n_tr, const char *fl, int n_fl,
                                     const char *fr, int n_fr, const char *motif, int m, int ref_size,
                                     int vcf_anchor_size, int max_iters, int local_search_range, int step_size,
                                     int respect_coords, int gap, const int8_t *matrix, int flags, int tie_flags,
                                     int32_t out8[8]) {
    if (m <= 0 || n_tr < 0 || n_fl < 0 || n_fr < 0 || n_fl + n_tr + n_fr <= 0) return 1;
    const int n_db = n_fl + n_tr + n_fr;
    char *db = concat3(fl, n_fl, tr, n_tr, fr, n_fr); /* :91 */
    char *db_rev = db ? reversed(db, n_db) : NULL;    /* :93 */
    if (!db || !db_rev) {
        free(db);
        free(db_rev);
        return 2;
    }
    int l_offset = 0, r_offset = 0, n_offset_scores = 0, rc = 0;
    omap seen; /* v0=fwd score, v1=r_adj, v2=rev score, v3=l_adj (both dicts share their key order, :126-127) */
    memset(&seen, 0, sizeof(seen));
    explore_t *stack = NULL;
    prof16 *pf = g_use_simd ? prof16_create(db, n_db, matrix) : NULL;     /* :92 */
    prof16 *pr = g_use_simd ? prof16_create(db_rev, n_db, matrix) : NULL; /* :93 */

    if (!respect_coords) { /* :99 */
        int cap = 16, top = 0;
        stack = (explore_t *)malloc(sizeof(explore_t) * (size_t)cap);
        stack[top++] = (explore_t){start_count - step_size, -1}; /* :100-101 */
        stack[top++] = (explore_t){start_count + step_size, 1};
        stack[top++] = (explore_t){start_count, 0};

        while (top > 0 && n_offset_scores < max_iters) { /* :106 */
            explore_t e = stack[--top];
            if (e.size < 0) continue; /* :108-109 */
            int wide = step_size > local_search_range;
            int start_size = e.size - ((e.dir < 1 || wide) ? local_search_range : 0); /* :114-115 */
            if (start_size < 0) start_size = 0;
            int end_size = e.size + ((e.dir > -1 || wide) ? local_search_range : 0); /* :116-117 */

            /* max((*fwd_scores, *rev_scores), key=(score, adj)) -> first maximal, fwd list first (:135) */
            int have = 0, mv_size = 0, mv_s = 0, mv_a = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int i = start_size; i <= end_size; ++i) {
                    int idx = omap_find(&seen, i);
                    if (idx < 0) { /* :123-130 */
                        int32_t r4[4] = {0, 0, 0, 0};
                        rc = score_ref_boundaries_rev(pf, pr, db, db_rev, n_db, fl, n_fl, fr, n_fr, motif, m, i, ref_size,
                                                      gap, matrix, r4);
                        if (rc) goto done;
                        idx = omap_push(&seen, i, r4[0], r4[1], r4[2], r4[3]);
                        ++n_offset_scores;
                    }
                    int s = pass == 0 ? seen.v0[idx] : seen.v2[idx];
                    int a = pass == 0 ? seen.v1[idx] : seen.v3[idx];
                    if (!have || pair_gt(s, a, mv_s, mv_a)) {
                        have = 1;
                        mv_size = i;
                        mv_s = s;
                        mv_a = a;
                    }
                }
            }
            if (top + 2 > cap) {
                cap *= 2;
                stack = (explore_t *)realloc(stack, sizeof(explore_t) * (size_t)cap);
            }
            if (mv_size > e.size) { /* :136-143 */
                int new_rc = mv_size + step_size;
                if (omap_find(&seen, new_rc) < 0 && new_rc >= 0) stack[top++] = (explore_t){new_rc, 1};
            }
            if (mv_size < e.size) { /* :144-151 */
                int new_rc = mv_size - step_size;
                if (omap_find(&seen, new_rc) < 0 && new_rc >= 0) stack[top++] = (explore_t){new_rc, -1};
            }
        }
        if (seen.n == 0) { /* max() of empty dict raises ValueError in the reference (:154) */
            rc = 3;
            goto done;
        }
        int bf = 0, br = 0; /* :154-156: first-inserted maximum by score only */
        for (int i = 1; i < seen.n; ++i) {
            if (seen.v0[i] > seen.v0[bf]) bf = i;
            if (seen.v2[i] > seen.v2[br]) br = i;
        }
        l_offset = seen.v3[br]; /* :161 */
        r_offset = seen.v1[bf]; /* :162 */
        if (l_offset >= n_fl - vcf_anchor_size) l_offset = 0; /* :164-167 */
        if (r_offset >= n_fr) r_offset = 0;                   /* :168-169 */
    }
    {
        /* :171-176: chunks of the flanks move into the tract; all three remain slices of db */
        int mov_l = l_offset > 0 ? l_offset : 0, mov_r = r_offset > 0 ? r_offset : 0;
        int nfl2 = n_fl - mov_l, nfr2 = n_fr - mov_r;
        /* :182 round() = banker's rounding of the float quotient */
        int start2 = (int)nearbyint(((double)start_count * (double)m + (double)(mov_l + mov_r)) / (double)m);
        int32_t r4[4] = {0, 0, 0, 0};
        /* tr_seq.upper() (:183) cannot change a case-insensitive score */
        rc = get_repeat_count_cells(start2, db, n_db, db, nfl2, db + n_db - nfr2, nfr2, motif, m, max_iters,
                                    local_search_range, step_size, gap, matrix, flags, tie_flags, r4, NULL);
        if (rc) goto done;
        out8[0] = r4[0];
        out8[1] = r4[1];
        out8[2] = l_offset;
        out8[3] = r_offset;
        out8[4] = n_offset_scores;
        out8[5] = r4[2];
        out8[6] = nfl2;
        out8[7] = nfr2;
    }
done:
    free(stack);
    omap_free(&seen);
    prof16_free(pf);
    prof16_free(pr);
    free(db);
    free(db_rev);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * Read loop of call_locus, call_locus.py:1079 (offset fraction init), :1129-1136 (start guess),
 * :1148-1155 (call), :1161 (offset update).  Float64 arithmetic as in CPython.
 * ------------------------------------------------------------------------------------------- */
#define LOCI_BLOCK 16
#define LOCUS_MEMO 512 /* reads of a locus the per-locus call memo covers (the reference keeps at most 250 reads) */

typedef struct {
    const char *arena;
    const uint64_t *seq_off;
    const int32_t *lens;
    const int32_t *est_cn;
    const int64_t *read_begin;
    const uint64_t *motif_off;
    const int32_t *motif_len;
    int64_t n_loci;
    int64_t *next_block; /* shared work queue cursor: blocks of LOCI_BLOCK loci (cf. loci.py:193-207) */
    int max_iters, range, step, gap, flags, tie_flags;
    const int8_t *matrix;
    int32_t *out;
    double cells;
    int rc;
    double read_cells[LOCUS_MEMO]; /* cells scored by each memoised call of the current locus */
} loci_job;

static void *loci_worker(void *arg) {
    loci_job *jb = (loci_job *)arg;
    for (;;) {
    int64_t b0 = __atomic_fetch_add(jb->next_block, (int64_t)LOCI_BLOCK, __ATOMIC_RELAXED);
    if (b0 >= jb->n_loci) break;
    int64_t b1 = b0 + LOCI_BLOCK < jb->n_loci ? b0 + LOCI_BLOCK : jb->n_loci;
    for (int64_t l = b0; l < b1; ++l) {
        double frac = 0.0; /* read_offset_frac_from_starting_guess, :1079 */
        const char *motif = jb->arena + jb->motif_off[l];
        int m = jb->motif_len[l];
        for (int64_t r = jb->read_begin[l]; r < jb->read_begin[l + 1]; ++r) {
            int read_sc = jb->est_cn[r];
            int off = (int)nearbyint(frac * (double)read_sc); /* round(), :1130 */
            if (off < -read_sc)
                frac = 0.0; /* :1133 */
            else
                read_sc += off; /* :1136 */
            const char *db = jb->arena + jb->seq_off[r];
            int nfl = jb->lens[3 * r], ntr = jb->lens[3 * r + 1], nfr = jb->lens[3 * r + 2];
            int32_t r4[4] = {0, 0, 0, 0};
            /* The reference wraps get_repeat_count in lru_cache(maxsize=512) (repeats.py:47): a call whose start count
             * and sequences equal an earlier call's is answered from the cache.  Hits across loci need equal flanks,
             * so the memo here is per locus.  `cells` keeps counting what the call would have scored (the GPU side
             * reports reference-equivalent cells per read the same way). */
            int64_t hit = -1;
            for (int64_t p = jb->read_begin[l]; p < r && p - jb->read_begin[l] < LOCUS_MEMO && hit < 0; ++p) {
                if (jb->out[4 * p + 3] == read_sc && jb->lens[3 * p] == nfl && jb->lens[3 * p + 1] == ntr &&
                    jb->lens[3 * p + 2] == nfr &&
                    memcmp(jb->arena + jb->seq_off[p], db, (size_t)(nfl + ntr + nfr)) == 0)
                    hit = p;
            }
            if (hit >= 0) {
                r4[0] = jb->out[4 * hit + 0];
                r4[1] = jb->out[4 * hit + 1];
                r4[2] = jb->out[4 * hit + 2];
                r4[3] = r4[0] - read_sc;
                jb->cells += jb->read_cells[hit - jb->read_begin[l]];
            } else {
                const double before = jb->cells;
                int rc = get_repeat_count_cells(read_sc, db, nfl + ntr + nfr, db, nfl, db + nfl + ntr, nfr, motif, m,
                                                jb->max_iters, jb->range, jb->step, jb->gap, jb->matrix, jb->flags,
                                                jb->tie_flags, r4, &jb->cells);
                if (rc) {
                    jb->rc = rc;
                    return NULL;
                }
                if (r - jb->read_begin[l] < LOCUS_MEMO) jb->read_cells[r - jb->read_begin[l]] = jb->cells - before;
            }
            jb->out[4 * r + 0] = r4[0];
            jb->out[4 * r + 1] = r4[1];
            jb->out[4 * r + 2] = r4[2];
            jb->out[4 * r + 3] = read_sc;
            int cn1 = r4[0] > 1 ? r4[0] : 1;
            frac += (double)r4[3] / (double)cn1; /* :1161 */
        }
    }
    }
    return NULL;
}

int strk_oracle_count_loci(const char *arena, const uint64_t *seq_off, const int32_t *lens, const int32_t *est_cn,
                           const int64_t *read_begin, const uint64_t *motif_off, const int32_t *motif_len,
                           int64_t n_loci, int max_iters, int local_search_range, int step_size, int gap,
                           const int8_t *matrix, int flags, int tie_flags, int n_threads, int32_t *out,
                           double *cells_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    if ((int64_t)n_threads > n_loci) n_threads = n_loci > 0 ? (int)n_loci : 1;
    loci_job *jobs = (loci_job *)calloc((size_t)n_threads, sizeof(loci_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!jobs || !th) {
        free(jobs);
        free(th);
        return 2;
    }
    /* dynamic queue of locus blocks, as the reference's worker pool pulls blocks (call_sample.py:103-111) */
    int64_t next_block = 0;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (loci_job){arena, seq_off, lens, est_cn, read_begin, motif_off, motif_len, n_loci, &next_block,
                             max_iters, local_search_range, step_size, gap, flags, tie_flags, matrix, out, 0.0, 0, {0.0}};
    }
    if (n_threads == 1) {
        loci_worker(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, loci_worker, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    int rc = 0;
    double cells = 0.0;
    for (int t = 0; t < n_threads; ++t) {
        if (jobs[t].rc) rc = jobs[t].rc;
        cells += jobs[t].cells;
    }
    if (cells_out) *cells_out = cells;
    free(jobs);
    free(th);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * Soft-clip realignment: parasail.sg_dx_trace_scan_16(ref_window, read, 7, 0, dna_matrix) as
 * strkit/call/realign.py:56-63 calls it -- s1 = the reference window (aligned end to end, both of its ends
 * penalised), s2 = the read (both ends free: "d", "x"), affine gaps: a gap of length k costs
 * open + (k - 1) * extend (7 and 0 there), traceback to a CIGAR that starts at cell (0, 0).
 *
 * PARITY UNPINNED: parasail is not in the reference tree.  Restated from the published recurrences (Gotoh, three
 * states) and parasail's documented result fields; what cannot be checked without the library is a switch:
 *   STRK_TRACE_OPEN_ON_TIE   a gap state that can be opened or extended at equal score is recorded as opened
 *                            (default: extended)
 *   STRK_TRACE_INS_FIRST     equal-scoring predecessors of a cell are tried diagonal, vertical (s1 only, 'I'),
 *                            horizontal (s2 only, 'D') (default: diagonal, horizontal, vertical)
 *   STRK_TRACE_END_LAST      the LAST best column of the last row ends the alignment (default: the first)
 * CIGAR encoding = parasail's / BAM's: (length << 4) | op with M=0 I=1 D=2 '='=7 X=8; 'I' consumes s1 only, 'D' s2 only
 * (parasail's query = s1); diagonal steps are '=' when the matrix scores the pair > 0, else 'X'.
 * Returns 0, or 1 bad arguments, 2 out of memory, 4 cigar_cap too small (needed length in *cigar_len).
 * ------------------------------------------------------------------------------------------- */
int strk_oracle_realign(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                        const int8_t *matrix, int trace_flags, int *score, int *end_ref, uint32_t *cigar, int cigar_cap,
                        int *cigar_len) {
    if (n1 <= 0 || n2 <= 0 || !s1 || !s2 || !matrix) return 1;
    const size_t W = (size_t)n2 + 1;
    int *H = (int *)malloc(sizeof(int) * W), *F = (int *)malloc(sizeof(int) * W);
    uint8_t *T = (uint8_t *)malloc((size_t)(n1 + 1) * W); /* bits 0-1: H source (0 diag, 1 horizontal, 2 vertical); 2: E extended; 3: F extended */
    uint8_t *c2 = (uint8_t *)malloc((size_t)n2);
    if (!H || !F || !T || !c2) {
        free(H), free(F), free(T), free(c2);
        return 2;
    }
    for (int j = 0; j < n2; ++j) c2[j] = (uint8_t)strk_oracle_symbol((unsigned char)s2[j]);
    const int open_tie = trace_flags & 1, ins_first = trace_flags & 2, end_last = trace_flags & 4;
    for (int j = 0; j <= n2; ++j) H[j] = 0, F[j] = NEG_INF; /* s2 begin free: row 0 is all zero */
    for (int i = 1; i <= n1; ++i) {
        const int8_t *mrow = matrix + STRK_NSYM * strk_oracle_symbol((unsigned char)s1[i - 1]);
        int diag = H[0];
        int left = -gap_open - (i - 1) * gap_extend; /* H[i][0]: s1 begin penalised */
        int E = NEG_INF;
        H[0] = left;
        T[(size_t)i * W] = 2; /* column 0: vertical */
        for (int j = 1; j <= n2; ++j) {
            const int up = H[j];
            uint8_t t = 0;
            const int e_opn = left - gap_open, e_ext = E - gap_extend;
            if (e_ext > e_opn || (e_ext == e_opn && !open_tie)) E = e_ext, t |= 4; else E = e_opn;
            const int f_opn = up - gap_open, f_ext = F[j] - gap_extend;
            if (f_ext > f_opn || (f_ext == f_opn && !open_tie)) F[j] = f_ext, t |= 8; else F[j] = f_opn;
            const int d = diag + mrow[c2[j - 1]];
            int h = d, src = 0;
            if (ins_first) {
                if (F[j] > h) h = F[j], src = 2;
                if (E > h) h = E, src = 1;
            } else {
                if (E > h) h = E, src = 1;
                if (F[j] > h) h = F[j], src = 2;
            }
            T[(size_t)i * W + (size_t)j] = (uint8_t)(t | src);
            diag = up;
            H[j] = h;
            left = h;
        }
    }
    int best = H[0], bj = 0; /* H[n1][0] is a legal end (whole window against nothing) */
    for (int j = 1; j <= n2; ++j)
        if (H[j] > best || (end_last && H[j] == best)) best = H[j], bj = j;
    if (score) *score = best;
    if (end_ref) *end_ref = bj - 1;
    /* traceback from (n1, bj) to row 0, then the free leading part of s2 as one 'D' run */
    int n_ops = 0, rc = 0;
    uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n1 + n2 + 2));
    if (!rev) {
        free(H), free(F), free(T), free(c2);
        return 2;
    }
    {
        int i = n1, j = bj, state = 0; /* 0 = in H, 1 = in E (horizontal run), 2 = in F (vertical run) */
        while (i > 0) {
            const uint8_t t = j > 0 ? T[(size_t)i * W + (size_t)j] : 2;
            uint32_t op;
            if (state == 0) state = t & 3;
            if (state == 0) {
                const int sc = matrix[STRK_NSYM * strk_oracle_symbol((unsigned char)s1[i - 1]) + c2[j - 1]];
                op = sc > 0 ? 7u : 8u;
                --i, --j;
            } else if (state == 1) {
                op = 2u; /* 'D': s2 only */
                if (!(t & 4)) state = 0;
                --j;
            } else {
                op = 1u; /* 'I': s1 only */
                if (j == 0 || !(t & 8)) state = 0;
                --i;
            }
            if (n_ops && (rev[n_ops - 1] & 15u) == op) rev[n_ops - 1] += 16u; else rev[n_ops++] = 16u | op;
        }
        if (j > 0) rev[n_ops++] = ((uint32_t)j << 4) | 2u; /* read bases before the window: deletions from (0, 0) */
    }
    if (cigar_len) *cigar_len = n_ops;
    if (n_ops > cigar_cap || !cigar) rc = cigar ? 4 : 0;
    else for (int k = 0; k < n_ops; ++k) cigar[k] = rev[n_ops - 1 - k];
    free(rev), free(H), free(F), free(T), free(c2);
    return rc;
}
