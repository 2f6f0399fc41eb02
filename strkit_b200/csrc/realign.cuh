// realign.cuh -- soft-clip realignment: parasail.sg_dx_trace_scan_16(reference window, read, open 7,
// extend 0, dna_matrix) of strkit/call/realign.py:56-63, with traceback to a CIGAR.
//
// s1 = the reference window (flank + tract + flank, a few hundred to a few thousand bases), aligned end to end;
// s2 = the whole read (tens of kilobases), both ends free.  True affine gaps (Gotoh, three states): a gap of length k
// costs open + (k - 1) * extend -- the reference passes 7 and 0, so this is NOT the linear-gap recurrence of the
// repeat-count kernels.
//
//   realign_fill_kernel<R>   one warp per alignment.  Lane t owns R consecutive rows of s1 (front-padded to a multiple
//                            of 32R; windows longer than 32R rows run as consecutive strips that hand a boundary row
//                            of (H, F) through global memory), the columns of s2 are swept as a skewed wavefront
//                            (lane t at column s - t + 1 in step s, bottom cell handed down by __shfl_up_sync).
//                            int32 cells.  Every cell leaves one trace byte -- bits 0-1 where H came from (0 diagonal,
//                            1 horizontal = E, 2 vertical = F), bit 2 E extended, bit 3 F extended -- stored in
//                            WAVEFRONT order: a step of a strip writes 32 lanes x R contiguous bytes (coalesced),
//                            instead of 32R bytes scattered over 32R matrix rows.  Lane 31 of the last strip keeps the
//                            best cell of the last row (s2 end free).
//   realign_trace_kernel     one thread per alignment walks back from that cell to row 0 and writes the CIGAR in
//                            parasail's / BAM's encoding ((len << 4) | op; I = s1 only, D = s2 only, '=' / 'X').
//
// Bound: HBM writes of the trace (1 byte per cell); the integer work is ~20 instructions per cell.
// The three tie rules parasail's traceback applies and the tree cannot show are switches (trace_flags), the same ones as
// in the CPU checker: bit 0 open-on-tie, bit 1 vertical before horizontal, bit 2 last best end column.
#pragma once
#include "dp_general.cuh"
#include "strk_common.cuh"

#define RA_NEG (-(1 << 28))

struct RealignDesc {
    unsigned long long s1_off, s2_off;  // arena offsets (ASCII)
    unsigned long long trace_off;       // byte offset of this alignment's trace
    unsigned long long cigar_off;       // element offset into the cigar output
    int n1, n2;
    int cigar_cap;
    int R, NB;                          // rows per lane, strips
};

__host__ __device__ inline int ra_pick_rows(int n1) {
    if (n1 <= 32 * 4) return 4;
    if (n1 <= 32 * 8) return 8;
    if (n1 <= 32 * 12) return 12;
    return 16;
}
__host__ __device__ inline unsigned long long ra_trace_bytes(int n1, int n2) {
    const int R = ra_pick_rows(n1);
    const int NB = (n1 + 32 * R - 1) / (32 * R);
    return (unsigned long long)NB * (unsigned long long)(n2 + 31) * 32ull * (unsigned long long)R;
}

template <int R>
__device__ void realign_fill_one(const RealignDesc &d, const unsigned char *__restrict__ arena, const short *smat2,
                                 const unsigned char *lut, unsigned char *__restrict__ trace, int *__restrict__ bound,
                                 int gap_open, int gap_ext, int trace_flags, int *score_out, int *end_out) {
    const int lane = threadIdx.x & 31;
    const unsigned char *s1 = arena + d.s1_off, *s2 = arena + d.s2_off;
    const int n1 = d.n1, n2 = d.n2, NB = d.NB;
    const int RB = 32 * R;
    const int off = NB * RB - n1;  // pad rows in front: they replay DP row 0 (H = 0, F = -inf: s2 begin is free)
    const bool open_tie = trace_flags & 1, ins_first = trace_flags & 2, end_last = trace_flags & 4;
    const int nsteps = n2 + 31;
    unsigned *tw = (unsigned *)(trace + d.trace_off);
    int best = -gap_open - (n1 - 1) * gap_ext, bj = 0;  // H[n1][0]
    // boundary rows between strips: [cur | nxt] x [H | F] x (n2 + 1)
    int *bH[2] = {bound, bound + 2 * (n2 + 1)}, *bF[2] = {bound + (n2 + 1), bound + 3 * (n2 + 1)};

    for (int b = 0; b < NB; ++b) {
        const int Ibase = b * RB + lane * R;  // padded index of the row above my first row
        int H[R], E[R], rowc[R];
        bool pad[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = Ibase + r + 1 - off;  // real row, <= 0 for pad rows
            pad[r] = i <= 0;
            rowc[r] = pad[r] ? 0 : lut[s1[i - 1]] * STRK_NSYM_;
            H[r] = pad[r] ? 0 : -gap_open - (i - 1) * gap_ext;  // column 0
            E[r] = RA_NEG;
        }
        const int i_above = Ibase - off;  // real index of the row above my strip part
        int prev_up = i_above <= 0 ? 0 : -gap_open - (i_above - 1) * gap_ext;  // its column-0 value (the diagonal of column 1)
        int hlast = H[R - 1], flast = RA_NEG;
        const int *topH = b ? bH[(b - 1) & 1] : nullptr, *topF = b ? bF[(b - 1) & 1] : nullptr;
        int *botH = b < NB - 1 ? bH[b & 1] : nullptr, *botF = b < NB - 1 ? bF[b & 1] : nullptr;
        int th_cur = 0, tf_cur = 0, th_nxt = 0, tf_nxt = 0;
        if (topH && 1 + lane <= n2) th_nxt = topH[1 + lane], tf_nxt = topF[1 + lane];
        unsigned code_nxt = (1 - lane >= 1 && 1 - lane <= n2) ? lut[s2[-lane]] : 0u;
        for (int s = 0; s < nsteps; ++s) {
            const int j = s - lane + 1;
            int hup = __shfl_up_sync(0xffffffffu, hlast, 1), fup = __shfl_up_sync(0xffffffffu, flast, 1);
            if (topH) {
                if ((s & 31) == 0) {
                    th_cur = th_nxt, tf_cur = tf_nxt;
                    if (s + 33 + lane <= n2) th_nxt = topH[s + 33 + lane], tf_nxt = topF[s + 33 + lane];
                }
                const int t0 = __shfl_sync(0xffffffffu, th_cur, s & 31), t1 = __shfl_sync(0xffffffffu, tf_cur, s & 31);
                if (lane == 0) hup = t0, fup = t1;
            } else if (lane == 0) {
                hup = 0, fup = RA_NEG;  // DP row 0
            }
            const unsigned code = code_nxt;
            if (j + 1 >= 1 && j + 1 <= n2) code_nxt = lut[s2[j]];
            if (j < 1 || j > n2) continue;
            int dg = prev_up, uh = hup, uf = fup;
            prev_up = hup;
            unsigned word = 0u;
            unsigned *dst = tw + ((size_t)((size_t)b * nsteps + s) * 32 + lane) * (R / 4);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int e_opn = H[r] - gap_open, e_ext = E[r] - gap_ext;
                const bool eb = e_ext > e_opn || (e_ext == e_opn && !open_tie);
                const int e = eb ? e_ext : e_opn;
                const int f_opn = uh - gap_open, f_ext = uf - gap_ext;
                const bool fb = f_ext > f_opn || (f_ext == f_opn && !open_tie);
                int f = fb ? f_ext : f_opn;
                int h = dg + (int)smat2[rowc[r] + code];
                unsigned src = 0u;
                if (ins_first) {
                    if (f > h) h = f, src = 2u;
                    if (e > h) h = e, src = 1u;
                } else {
                    if (e > h) h = e, src = 1u;
                    if (f > h) h = f, src = 2u;
                }
                if (pad[r]) h = 0, f = RA_NEG;
                dg = H[r];
                H[r] = h;
                E[r] = e;
                uh = h;
                uf = f;
                word |= (src | (eb ? 4u : 0u) | (fb ? 8u : 0u)) << (8 * (r & 3));
                if ((r & 3) == 3) {
                    dst[r >> 2] = word;
                    word = 0u;
                }
            }
            hlast = uh;
            flast = uf;
            if (lane == 31) {
                if (botH) {
                    botH[j] = uh;
                    botF[j] = uf;
                } else if (uh > best || (end_last && uh == best)) {
                    best = uh;
                    bj = j;
                }
            }
        }
        __threadfence();  // the next strip reads the boundary row this one wrote
        __syncwarp();
    }
    if (lane == 31) {
        *score_out = best;
        *end_out = bj;
    }
}

#define RA_WARPS 4
__global__ void __launch_bounds__(RA_WARPS * 32) realign_fill_kernel(const RealignDesc *__restrict__ descs, int n,
                                                                     const unsigned char *__restrict__ arena,
                                                                     const ScoreConsts *__restrict__ consts,
                                                                     unsigned char *__restrict__ trace,
                                                                     int *__restrict__ bound, int bound_stride,
                                                                     int gap_open, int gap_ext, int trace_flags,
                                                                     int *__restrict__ score, int *__restrict__ end_col,
                                                                     unsigned int *queue) {
    __shared__ unsigned char lut[256];
    __shared__ short smat2[STRK_NSYM_ * STRK_NSYM_];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) lut[k] = consts->lut[k];
    for (int k = threadIdx.x; k < STRK_NSYM_ * STRK_NSYM_; k += blockDim.x) smat2[k] = consts->smat[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * RA_WARPS + (threadIdx.x >> 5);
    int *my_bound = bound + (size_t)warp_global * (size_t)bound_stride;
    for (;;) {
        unsigned q = 0;
        if (lane == 0) q = atomicAdd(queue, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= (unsigned)n) break;
        const RealignDesc d = descs[q];
        switch (d.R) {
            case 4: realign_fill_one<4>(d, arena, smat2, lut, trace, my_bound, gap_open, gap_ext, trace_flags, score + q, end_col + q); break;
            case 8: realign_fill_one<8>(d, arena, smat2, lut, trace, my_bound, gap_open, gap_ext, trace_flags, score + q, end_col + q); break;
            case 12: realign_fill_one<12>(d, arena, smat2, lut, trace, my_bound, gap_open, gap_ext, trace_flags, score + q, end_col + q); break;
            default: realign_fill_one<16>(d, arena, smat2, lut, trace, my_bound, gap_open, gap_ext, trace_flags, score + q, end_col + q); break;
        }
        __syncwarp();
    }
}

// One thread per alignment: walk back from (n1, end column) to row 0; runs are collected newest-first at the END of the
// alignment's cigar region and moved to its front in order.  cigar_len < 0 reports a region that was too small.
__global__ void realign_trace_kernel(const RealignDesc *__restrict__ descs, int n, const unsigned char *__restrict__ arena,
                                     const ScoreConsts *__restrict__ consts, const unsigned char *__restrict__ trace,
                                     const int *__restrict__ end_col, unsigned int *__restrict__ cigar,
                                     int *__restrict__ cigar_len) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const RealignDesc d = descs[q];
    const unsigned char *s1 = arena + d.s1_off, *s2 = arena + d.s2_off;
    const unsigned char *tr = trace + d.trace_off;
    const int R = d.R, RB = 32 * R, off = d.NB * RB - d.n1, nsteps = d.n2 + 31;
    unsigned int *out = cigar + d.cigar_off;
    int i = d.n1, j = end_col[q], state = 0, n_ops = 0, at = d.cigar_cap;  // runs grow downwards from out[cap - 1]
    unsigned cur = 0u;
    bool overflow = false;
    auto push = [&](unsigned op) {
        if (cur && (cur & 15u) == op) {
            cur += 16u;
            return;
        }
        if (cur) {
            if (at == 0) overflow = true; else out[--at] = cur;
            ++n_ops;
        }
        cur = 16u | op;
    };
    while (i > 0) {
        unsigned t = 2u;
        if (j > 0) {
            const int I = i + off - 1, b = I / RB, lane = (I % RB) / R, r = I % R;
            t = tr[((size_t)((size_t)b * nsteps + (j + lane - 1)) * 32 + lane) * R + r];
        }
        if (state == 0) state = (int)(t & 3u);
        if (state == 0) {
            const int sc = consts->smat[consts->lut[s1[i - 1]] * STRK_NSYM_ + consts->lut[s2[j - 1]]];
            push(sc > 0 ? 7u : 8u);
            --i, --j;
        } else if (state == 1) {
            push(2u);
            if (!(t & 4u)) state = 0;
            --j;
        } else {
            push(1u);
            if (j == 0 || !(t & 8u)) state = 0;
            --i;
        }
    }
    if (cur) {
        if (at == 0) overflow = true; else out[--at] = cur;
        ++n_ops;
    }
    if (j > 0) {  // read bases before the window: one run of deletions, so that the CIGAR starts at cell (0, 0)
        if (at == 0) overflow = true; else out[--at] = ((unsigned)j << 4) | 2u;
        ++n_ops;
    }
    if (overflow) {
        cigar_len[q] = -n_ops;
        return;
    }
    for (int k = 0; k < n_ops; ++k) out[k] = out[at + k];  // at >= k always: in-place move towards the front
    cigar_len[q] = n_ops;
}
