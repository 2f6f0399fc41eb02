"""Test harness: the reference's per-read loop (strkit/call/call_locus.py:1079-1288) with Python shims for the
strkit_rust_ext objects it touches (SURVEY appendix A, rows marked S).  TEST INFRASTRUCTURE ONLY.

The loop below follows the reference statement by statement for everything that involves the hot path: start guess
with the carried offset (:1129-1136), flank slicing (:1144-1146), the get_repeat_count call (:1148-1155), offset update
(:1161), the max-iterations log line (:1164-1170), calc_adj_score (:1172), the min-score filter and the
terrible-read abort (:1222-1250), read weight (:1259) and the read_dict entry (:1279-1288).  What it takes as
parameters is exactly what the drop-in must not care about: WHICH get_repeat_count it is handed.  The two Rust
formulas that are not in the reference tree (get_est_copy_num, calc_adj_score) are stand-ins; they sit outside the
boundary and are the same in both modes under comparison.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class SegmentShim:
    """STRkitAlignedSegment + STRkitAlignedSegmentSequenceDataForLocus, the members the loop reads."""
    name: str
    is_reverse: bool
    length: int
    tr_seq_wc: str
    flank_left_seq_wc: str
    flank_right_seq_wc: str
    motif_size: int

    def get_est_copy_num(self) -> int:
        return round(len(self.tr_seq_wc) / self.motif_size)   # stand-in (formula not in the reference tree)

    def calc_adj_score(self, score: int):
        if not self.tr_seq_wc:
            return None
        n = len(self.tr_seq_wc) + min(70, len(self.flank_left_seq_wc)) + min(70, len(self.flank_right_seq_wc))
        return score / (2.0 * n)                              # stand-in: a 0-1 fraction (docs/caller_usage.md:19-21)

    @property
    def tr_len_with_flank(self) -> int:
        return len(self.tr_seq_wc) + len(self.flank_left_seq_wc) + len(self.flank_right_seq_wc)


def read_loop(motif: str, segments, get_repeat_count, rc_params, flank_size: int = 70,
              min_read_align_score: float = 0.9, max_terrible_reads: int = 3,
              extremely_low_read_adj_score: float = 0.4):
    """Returns (read_dict | None when the locus is abandoned, log lines)."""
    read_dict = {}
    log = []
    extremely_poor_scoring_reads = []
    read_offset_frac_from_starting_guess = 0.0                                   # :1079
    for segment in segments:                                                     # :1082
        read_sc = segment.get_est_copy_num()                                     # :1129
        if (read_sc_offset := round(read_offset_frac_from_starting_guess * read_sc)) < -1 * read_sc:
            read_offset_frac_from_starting_guess = 0.0                           # :1133
        else:
            read_sc += read_sc_offset                                            # :1136
        fls = segment.flank_left_seq_wc[-1 * flank_size:]                        # :1145
        frs = segment.flank_right_seq_wc[:flank_size]                            # :1146
        (read_cn, read_cn_score), n_read_cn_iters, new_offset_from_starting_count = get_repeat_count(
            start_count=read_sc, tr_seq=segment.tr_seq_wc, flank_left_seq=fls, flank_right_seq=frs, motif=motif,
            rc_params=rc_params)                                                 # :1148-1155
        rn = segment.name
        read_offset_frac_from_starting_guess += new_offset_from_starting_count / max(read_cn, 1)   # :1161
        if n_read_cn_iters >= rc_params.max_iters:                               # :1164
            log.append(f"{rn}: read repeat counting exceeded maximum # iterations ({n_read_cn_iters})")
        read_adj_score = segment.calc_adj_score(read_cn_score)                   # :1172
        if read_adj_score is not None and read_adj_score < min_read_align_score:  # :1222
            log.append(f"skipping read {rn} (repeat count alignment scored {read_adj_score:.2f})")
            if read_adj_score < extremely_low_read_adj_score:                    # :1234
                extremely_poor_scoring_reads.append((rn, read_adj_score))
                if len(extremely_poor_scoring_reads) > max_terrible_reads:       # :1236
                    log.append("not calling locus due to extremely poor-aligning reads")
                    return None, log                                             # :1247-1250
            continue                                                             # :1252
        read_weight = 1.0 / max(1, len(segments))                                # stand-in for get_read_weight (:1259)
        read_dict[rn] = {"s": "-" if segment.is_reverse else "+", "cn": read_cn, "w": read_weight,
                         "sc": read_adj_score}                                   # :1279-1288
    return read_dict, log
