/*
 * strk_oracle.h -- CPU oracle for STRkit's per-read repeat-count hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under strkit_b200/ may import, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported baseline.
 *
 * PARITY STATUS: "parity unpinned" for the parts whose arithmetic lives in
 * un-vendored third-party code (parasail 1.3.4 C library; strkit_rust_ext 0.29.0
 * get_repeat_count).  Those parts restate the published algorithm and are anchored
 * on the reference's call sites.  The in-tree control flow of
 * get_ref_repeat_count (reference strkit/call/repeats.py:73-192) IS pinned: see
 * tests/golden/gen_golden.py, which executes the reference's own repeats.py with a
 * stubbed DP and commits the vectors.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#ifndef STRK_ORACLE_H
#define STRK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Free-end flags of the semi-global alignment.  s1 = the profiled sequence
 * ("query" in parasail terms: db = flank_left+tr+flank_right, repeats.py:91-92),
 * s2 = the candidate.  parasail names: q = s1, d = s2, b = begin, e = end. */
#define STRK_S1_BEG_FREE 1 /* first column initialised to 0                */
#define STRK_S1_END_FREE 2 /* result = max over last column (sg_qe)        */
#define STRK_S2_BEG_FREE 4 /* first row initialised to 0                   */
#define STRK_S2_END_FREE 8 /* result = max over last row                   */
#define STRK_MODE_SG 15    /* parasail "sg": all four ends free            */
#define STRK_MODE_SG_QE 2  /* parasail "sg_qe" (repeats.py:33,40)          */
#define STRK_MODE_NW 0

/* Tie-break switches for the (unverifiable) Rust hill-climb.  0 = the in-tree
 * Python semantics of repeats.py:135,154-156 (first maximal element). */
#define STRK_TIE_WINDOW_LAST 1 /* window max picks the LAST maximal size   */
#define STRK_TIE_FINAL_LAST 2  /* final pick takes the LAST-inserted max   */

/* Search-policy switches, same flag word (hypotheses about the Rust body; 0 = range fixed as in repeats.py:100-151).
 * reference strkit/call/repeat_count_params.py:13: initial_local_search_range "can be narrowed within the
 * get_repeat_count fn". */
#define STRK_SEARCH_NARROW_FIRST 4 /* range becomes 1 after the first (direction 0) window   */
#define STRK_SEARCH_NARROW_HALVE 8 /* range is halved (not below 1) after every window       */

#define STRK_NSYM 17 /* 16-letter alphabet + parasail wildcard column      */

/* align_matrix.py:15-44 + iupac.py:9-21: 17x17 substitution matrix, row-major. */
void strk_oracle_dna_matrix(int8_t out[STRK_NSYM * STRK_NSYM]);

/* parasail matrix_create mapper: case-insensitive, unknown bytes -> 16. */
int strk_oracle_symbol(unsigned char c);

/* One semi-global alignment, exact int32 arithmetic (== parasail *_sat results).
 * Returns 0 on success, non-zero on invalid arguments (empty sequence). */
int strk_oracle_sg_align(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                         const int8_t *matrix, int flags, int *score, int *end_query, int *end_ref);

/* The same alignment through the AVX2 kernel (16-bit lanes, striped profile + column scan, the layout of
 * parasail's *_scan_profile kernels); identical results, falls back to the scalar code where it does not apply. */
int strk_oracle_sg_align_simd(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                              const int8_t *matrix, int flags, int *score, int *end_query, int *end_ref);
/* 1: the search / batch entry points below use the AVX2 kernel (the timed CPU baseline); 0 (default): scalar.
 * Returns the previous setting.  strk_oracle_have_simd: 1 when the library was built with AVX2. */
int strk_oracle_set_simd(int on);
int strk_oracle_have_simd(void);

/* Score of candidate fl + motif*n + fr against db = fl + tr + fr (read path). */
int strk_oracle_score_candidate(const char *db, int n_db, const char *fl, int n_fl, const char *fr, int n_fr,
                                const char *motif, int m, int n, int gap, const int8_t *matrix, int flags,
                                int *score);

/* strkit_rust_ext.get_repeat_count (repeats.py:58-68): hill-climb over n.
 * out4 = {best_n, best_score, n_explored, best_n - start_count}. */
int strk_oracle_get_repeat_count(int start_count, const char *tr, int n_tr, const char *fl, int n_fl,
                                 const char *fr, int n_fr, const char *motif, int m, int max_iters,
                                 int local_search_range, int step_size, int gap, const int8_t *matrix, int flags,
                                 int tie_flags, int32_t out4[4]);

/* score_ref_boundaries (repeats.py:23-43) for one candidate size n.
 * out4 = {fwd_score, r_adj, rev_score, l_adj}. */
int strk_oracle_score_ref_boundaries(const char *db, int n_db, const char *fl, int n_fl, const char *fr, int n_fr,
                                     const char *motif, int m, int n, int ref_size, int gap, const int8_t *matrix,
                                     int32_t out4[4]);

/* get_ref_repeat_count (repeats.py:73-192).
 * out8 = {cn, score, l_offset, r_offset, n_offset_scores, n_iters_final,
 *         new_fl_len, new_fr_len}; the adjusted (fl, tr, fr) are slices of the
 * same concatenation, so the two lengths describe them completely. */
int strk_oracle_get_ref_repeat_count(int start_count, const char *tr, int n_tr, const char *fl, int n_fl,
                                     const char *fr, int n_fr, const char *motif, int m, int ref_size,
                                     int vcf_anchor_size, int max_iters, int local_search_range, int step_size,
                                     int respect_coords, int gap, const int8_t *matrix, int flags, int tie_flags,
                                     int32_t out8[8]);

/* Per-locus read loop of call_locus.py:1079,1129-1161: start guess with the
 * carried offset fraction, one get_repeat_count per read, in read order.
 * Reads of locus l are read_begin[l] .. read_begin[l+1]-1.
 * seq_off[r] = offset of read r's fl+tr+fr concatenation in `arena`,
 * lens[3r..3r+2] = {n_fl, n_tr, n_fr}; motif_off/motif_len per locus.
 * out[4r..4r+3] = {cn, score, n_explored, start_count used}.
 * cells_out (optional) accumulates reference-equivalent DP cells.
 * n_threads > 1 splits loci over pthreads (the timed CPU baseline). */
int strk_oracle_count_loci(const char *arena, const uint64_t *seq_off, const int32_t *lens, const int32_t *est_cn,
                           const int64_t *read_begin, const uint64_t *motif_off, const int32_t *motif_len,
                           int64_t n_loci, int max_iters, int local_search_range, int step_size, int gap,
                           const int8_t *matrix, int flags, int tie_flags, int n_threads, int32_t *out,
                           double *cells_out);

/* Soft-clip realignment, strkit/call/realign.py:56-63: parasail.sg_dx_trace_scan_16(s1 = reference window, s2 = read,
 * open, extend, dna_matrix) -> score, end_ref (last read position aligned), CIGAR from cell (0, 0) in parasail / BAM
 * encoding ((len << 4) | op; I = s1 only, D = s2 only, 7 '=' / 8 'X' on the diagonal).  PARITY UNPINNED (see the .c). */
#define STRK_TRACE_OPEN_ON_TIE 1
#define STRK_TRACE_INS_FIRST 2
#define STRK_TRACE_END_LAST 4
int strk_oracle_realign(const char *s1, int n1, const char *s2, int n2, int gap_open, int gap_extend,
                        const int8_t *matrix, int trace_flags, int *score, int *end_ref, uint32_t *cigar, int cigar_cap,
                        int *cigar_len);

#ifdef __cplusplus
}
#endif
#endif
