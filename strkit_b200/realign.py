"""Soft-clip realignment on the device: strkit/call/realign.py:34-72 (`realign_read`), i.e.
parasail.sg_dx_trace_scan_16(ref_seq, query_seq, 7, 0, dna_matrix) + get_aligned_pair_matches(cigar, left_flank_coord,
0, swap=True), batched over reads.

What is pinned and what is not: the call shape, the threshold (`min_realign_score_ratio * (flank_size * 2 *
match_score - realign_indel_open_penalty)`, realign.py:65) and the gap model are the reference's; the traceback's tie
rules (three switches, `trace_flags`) and the exact output of the Rust helper `get_aligned_pair_matches` are not in the
reference tree -- the pairs returned here are the positions of the diagonal steps ('=' and 'X') of the CIGAR, read
coordinate first (swap=True), which is what the name and the call site (call_locus.py:880-901) imply.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from ._native import check, lib
from .align_matrix import match_score
from .engine import Engine, default_engine

__all__ = ["realign_batch", "realign_read", "cigar_to_string", "aligned_pairs_from_cigar", "min_realign_score_ratio",
           "realign_indel_open_penalty"]

min_realign_score_ratio: float = 0.95      # realign.py:28
realign_indel_open_penalty: int = 7        # realign.py:29
_OPS = "MIDNSHP=XB"


def cigar_to_string(cigar: np.ndarray) -> str:
    return "".join(f"{int(c) >> 4}{_OPS[int(c) & 15]}" for c in cigar)


def aligned_pairs_from_cigar(cigar: np.ndarray, query_start: int, ref_start: int, swap: bool = False):
    """Coordinates of the diagonal steps of a CIGAR that starts at (query_start, ref_start); parasail's query is s1.
    Returns (coords of s1, coords of s2), or the two swapped."""
    q, r = int(query_start), int(ref_start)
    qs, rs = [], []
    for c in cigar:
        n, op = int(c) >> 4, int(c) & 15
        if op in (0, 7, 8):
            qs.append(np.arange(q, q + n, dtype=np.int64))
            rs.append(np.arange(r, r + n, dtype=np.int64))
            q += n
            r += n
        elif op == 1:
            q += n
        elif op == 2:
            r += n
    qa = np.concatenate(qs) if qs else np.zeros(0, dtype=np.int64)
    ra = np.concatenate(rs) if rs else np.zeros(0, dtype=np.int64)
    return (ra, qa) if swap else (qa, ra)


def realign_batch(pairs: Sequence[tuple[str, str]], engine: Engine | None = None, gap_open: int = realign_indel_open_penalty,
                  gap_extend: int = 0, trace_flags: int = 0):
    """[(ref_seq, query_seq)] -> [(score, end_ref, cigar uint32[])]: one C-ABI call (strk_realign) for all pairs."""
    if not pairs:
        return []
    eng = engine or default_engine()
    blobs, ref_off, ref_len, read_off, read_len, at = [], [], [], [], [], 0
    for ref_seq, query_seq in pairs:
        a, b = ref_seq.encode("ascii"), query_seq.encode("ascii")
        ref_off.append(at)
        ref_len.append(len(a))
        read_off.append(at + len(a))
        read_len.append(len(b))
        blobs += [a, b]
        at += len(a) + len(b)
    arena = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    n = len(pairs)
    ref_off, read_off = np.asarray(ref_off, dtype=np.uint64), np.asarray(read_off, dtype=np.uint64)
    ref_len, read_len = np.asarray(ref_len, dtype=np.int32), np.asarray(read_len, dtype=np.int32)
    cap = 2 * ref_len.astype(np.int64) + 4
    cigar_off = np.zeros(n + 1, dtype=np.uint64)
    cigar_off[1:] = np.cumsum(cap)
    cigar = np.zeros(int(cigar_off[-1]), dtype=np.uint32)
    score, end_ref, cigar_len = (np.zeros(n, dtype=np.int32) for _ in range(3))
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    with eng.lock:
        check(lib.strk_realign(eng._ctx, p(arena), arena.nbytes, p(ref_off), p(ref_len), p(read_off), p(read_len), n,
                               gap_open, gap_extend, trace_flags, p(score), p(end_ref), p(cigar), p(cigar_off), p(cigar_len)))
    return [(int(score[k]), int(end_ref[k]), cigar[int(cigar_off[k]):int(cigar_off[k]) + int(cigar_len[k])].copy())
            for k in range(n)]


def realign_read(ref_seq: str, query_seq: str, left_flank_coord: int, flank_size: int, q=None, read_log_str: str = "",
                 log_level: int = 0, engine: Engine | None = None):
    """Signature of strkit.call.realign.realign_read (realign.py:34-42).  Returns None when the alignment scores below
    the reference's threshold (:65), else (read coordinates, reference coordinates) of the aligned pairs (:71)."""
    (score, _end_ref, cigar), = realign_batch([(ref_seq, query_seq)], engine=engine)
    res = None
    if score >= min_realign_score_ratio * (flank_size * 2 * match_score - realign_indel_open_penalty):
        res = aligned_pairs_from_cigar(cigar, left_flank_coord, 0, swap=True)
    if q is not None:   # the reference hands the result to its parent process through a queue (:44-48)
        q.put(res)
        q.close()
    return res
