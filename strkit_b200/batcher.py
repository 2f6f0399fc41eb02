"""Host-side batcher: packs the hot-path tuples of many loci into flat arenas for the device.

Replaces the per-locus, per-read Python->Rust calls of the reference's worker pool
(strkit/call/call_sample.py:103-157 -> call_locus.py:1082-1161 -> repeats.py:47-70): instead of one
FFI call per read, the reads of a whole block of loci become ONE call across the C ABI.

A tuple is exactly what the reference passes to get_repeat_count (call_locus.py:1148-1155):
(start-count estimate, tr_seq_wc, flank_left_seq_wc[-flank:], flank_right_seq_wc[:flank], motif).
Sequences stay ASCII (any case, IUPAC, 'X' wildcards); the device encodes them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

__all__ = ["ReadBatch", "LocusReads", "pack_loci", "ARENA_ASCII", "ARENA_NIBBLE"]

ARENA_ASCII = 0    # one byte per symbol, as the reference passes its strings
ARENA_NIBBLE = 1   # two symbols per byte, low nibble first, code = index into "ACGTRYSWKMBDHVNX" (align_matrix.py:25-26)
_ALPHABET = b"ACGTRYSWKMBDHVNX"
_NIBBLE_OF = np.full(256, 16, dtype=np.uint8)
for _i, _c in enumerate(_ALPHABET):
    _NIBBLE_OF[_c] = _i
    _NIBBLE_OF[ord(chr(_c).lower())] = _i


@dataclass
class LocusReads:
    """The reads of one locus, in the order the reference iterates them (call_locus.py:1082)."""
    motif: str
    est_cn: Sequence[int]          # get_est_copy_num() per read (call_locus.py:1129)
    tr_seqs: Sequence[str]         # tr_seq_wc
    flank_left_seqs: Sequence[str]
    flank_right_seqs: Sequence[str]


@dataclass
class ReadBatch:
    arena: np.ndarray        # uint8, ASCII
    seq_off: np.ndarray      # uint64 [n_reads]   offset of fl + tr + fr
    lens: np.ndarray         # int32  [n_reads, 3]
    est_cn: np.ndarray       # int32  [n_reads]
    read_begin: np.ndarray   # int64  [n_loci + 1]
    motif_off: np.ndarray    # uint64 [n_loci]
    motif_len: np.ndarray    # int32  [n_loci]
    arena_format: int = ARENA_ASCII   # ARENA_NIBBLE: `arena` holds two symbols per byte; offsets / lengths stay in symbols

    @property
    def n_reads(self) -> int:
        return int(self.est_cn.shape[0])

    @property
    def n_loci(self) -> int:
        return int(self.motif_len.shape[0])

    def validate(self) -> None:
        assert self.arena.dtype == np.uint8 and self.arena.flags.c_contiguous
        assert self.seq_off.dtype == np.uint64 and self.lens.dtype == np.int32 and self.est_cn.dtype == np.int32
        assert self.read_begin.dtype == np.int64 and self.motif_off.dtype == np.uint64
        assert self.motif_len.dtype == np.int32
        assert self.lens.shape == (self.n_reads, 3) and self.read_begin.shape == (self.n_loci + 1,)

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in (self.arena, self.seq_off, self.lens, self.est_cn, self.read_begin,
                                          self.motif_off, self.motif_len)))

    def slice_loci(self, lo: int, hi: int, compact: bool = False) -> "ReadBatch":
        """Loci [lo, hi) as their own batch.  compact=False shares the whole arena (offsets stay valid; cheap, but
        every call that uploads the slice uploads the whole arena).  compact=True cuts the arena to the slice: the
        byte range its reads span (reads of consecutive loci are contiguous in every packer of this package) followed by
        its motifs, offsets rebased -- what a rank of a catalog partition should hold and upload."""
        r0, r1 = int(self.read_begin[lo]), int(self.read_begin[hi])
        if not compact or self.arena_format != ARENA_ASCII:
            return ReadBatch(self.arena, self.seq_off[r0:r1].copy(), self.lens[r0:r1].copy(), self.est_cn[r0:r1].copy(),
                             (self.read_begin[lo:hi + 1] - r0).copy(), self.motif_off[lo:hi].copy(),
                             self.motif_len[lo:hi].copy(), self.arena_format)
        seq_off, lens = self.seq_off[r0:r1].astype(np.int64), self.lens[r0:r1]
        if r1 > r0:
            a0 = int(seq_off.min())
            a1 = int((seq_off + lens.sum(axis=1, dtype=np.int64)).max())
        else:
            a0 = a1 = 0
        ml = self.motif_len[lo:hi].astype(np.int64)
        mo = self.motif_off[lo:hi].astype(np.int64)
        m_new = (a1 - a0) + np.concatenate([[0], np.cumsum(ml)[:-1]]) if hi > lo else np.zeros(0, dtype=np.int64)
        msrc = np.repeat(mo - m_new, ml) + (a1 - a0) + np.arange(int(ml.sum()), dtype=np.int64) if hi > lo else np.zeros(0, np.int64)
        arena = np.concatenate([self.arena[a0:a1], self.arena[msrc] if msrc.size else np.zeros(0, np.uint8)])
        return ReadBatch(arena, (seq_off - a0).astype(np.uint64), lens.copy(), self.est_cn[r0:r1].copy(),
                         (self.read_begin[lo:hi + 1] - r0).copy(), m_new.astype(np.uint64), self.motif_len[lo:hi].copy())

    def to_ascii(self) -> "ReadBatch":
        """Inverse of to_nibble (upper-case letters; offsets unchanged)."""
        if self.arena_format == ARENA_ASCII:
            return self
        letters = np.frombuffer(_ALPHABET, dtype=np.uint8)
        out = np.empty(2 * self.arena.shape[0], dtype=np.uint8)
        out[0::2] = letters[self.arena & 15]
        out[1::2] = letters[self.arena >> 4]
        return ReadBatch(out, self.seq_off, self.lens, self.est_cn, self.read_begin, self.motif_off, self.motif_len,
                         ARENA_ASCII)

    def to_nibble(self) -> "ReadBatch":
        """The same batch with a nibble-packed arena (half the host-to-device bytes).  Raises ValueError when the arena
        holds a byte outside the 16-letter alphabet: parasail's wildcard column has no nibble code, such a block stays
        ASCII."""
        if self.arena_format == ARENA_NIBBLE:
            return self
        codes = _NIBBLE_OF[self.arena]
        if codes.size and int(codes.max()) > 15:
            raise ValueError("arena holds bytes outside ACGTRYSWKMBDHVNX: not representable in the nibble format")
        if codes.size & 1:
            codes = np.concatenate([codes, np.zeros(1, dtype=np.uint8)])
        packed = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)
        return ReadBatch(packed, self.seq_off, self.lens, self.est_cn, self.read_begin, self.motif_off, self.motif_len,
                         ARENA_NIBBLE)


try:  # CPython helper built by __graft_entry__.build() (csrc/fastpack.c): two passes over the objects, no temporaries
    from . import _fastpack
except ImportError:  # not built: the pure-Python body below does the same, ~6x slower
    _fastpack = None


def pack_loci(loci: Iterable[LocusReads], use_helper: bool = True, nibble: bool = False, threads: int = 0) -> ReadBatch:
    """Pack per-locus string tuples into one arena.  With the C helper (csrc/fastpack.c): one walk over the Python
    objects under the GIL, then the bytes are moved by `threads` pthreads with the GIL released (0 = one per core, at
    most 16).  nibble=True emits the ARENA_NIBBLE layout (half the host-to-device bytes); a block holding a byte outside
    the 16-letter alphabet silently stays ASCII (check .arena_format).  Sequences may be str or bytes.  Without the
    helper the per-read Python work is kept to list building done by C-level iterators (zip / chain / map).
    (strkit_b200.synth builds fully vectorised synthetic batches.)"""
    from itertools import chain

    loci = list(loci)
    if _fastpack is not None and use_helper:
        fmt = ARENA_NIBBLE if nibble else ARENA_ASCII
        try:
            arena, seq_off, lens, est, read_begin, motif_off, motif_len = _fastpack.pack(loci, nibble, threads)
        except ValueError as exc:
            if not nibble or "nibble" not in str(exc):
                raise
            fmt = ARENA_ASCII
            arena, seq_off, lens, est, read_begin, motif_off, motif_len = _fastpack.pack(loci, False, threads)
        return ReadBatch(arena=np.frombuffer(arena, dtype=np.uint8), seq_off=np.frombuffer(seq_off, dtype=np.uint64),
                         lens=np.frombuffer(lens, dtype=np.int32).reshape(-1, 3),
                         est_cn=np.frombuffer(est, dtype=np.int32), read_begin=np.frombuffer(read_begin, dtype=np.int64),
                         motif_off=np.frombuffer(motif_off, dtype=np.uint64),
                         motif_len=np.frombuffer(motif_len, dtype=np.int32), arena_format=fmt)
    n_per = []
    for lr in loci:
        n = len(lr.tr_seqs)
        if not (len(lr.est_cn) == len(lr.flank_left_seqs) == len(lr.flank_right_seqs) == n):
            raise ValueError("LocusReads: per-read sequences must have equal lengths")
        n_per.append(n)
    n_reads = sum(n_per)
    # fl, tr, fr of every read, in read order
    _s = lambda x: x.decode("ascii") if isinstance(x, (bytes, bytearray)) else x  # noqa: E731
    seqs = list(map(_s, chain.from_iterable(chain.from_iterable(zip(lr.flank_left_seqs, lr.tr_seqs, lr.flank_right_seqs))
                                            for lr in loci)))
    lens_a = np.fromiter(map(len, seqs), dtype=np.int32, count=3 * n_reads).reshape(-1, 3)
    est = np.fromiter(chain.from_iterable(lr.est_cn for lr in loci), dtype=np.int32, count=n_reads)
    motifs = [_s(lr.motif) for lr in loci]
    motif_len = np.fromiter(map(len, motifs), dtype=np.int32, count=len(motifs))
    read_begin = np.zeros(len(loci) + 1, dtype=np.int64)
    np.cumsum(n_per, out=read_begin[1:])
    tot = lens_a.sum(axis=1, dtype=np.int64)
    seq_off = np.zeros(n_reads, dtype=np.uint64)
    if n_reads:
        seq_off[1:] = np.cumsum(tot)[:-1]
    seq_bytes = int(tot.sum())
    motif_off = np.zeros(len(motifs), dtype=np.uint64)
    if len(motifs):
        motif_off[:] = seq_bytes + np.concatenate([[0], np.cumsum(motif_len, dtype=np.int64)[:-1]])
    blob = ("".join(seqs) + "".join(motifs)).encode("ascii")
    if len(blob) != seq_bytes + int(motif_len.sum(dtype=np.int64)):
        raise ValueError("pack_loci: sequences must be ASCII")
    arena = np.frombuffer(blob, dtype=np.uint8)
    batch = ReadBatch(arena=arena, seq_off=seq_off, lens=lens_a, est_cn=est, read_begin=read_begin, motif_off=motif_off,
                      motif_len=motif_len)
    if nibble:
        try:
            return batch.to_nibble()
        except ValueError:
            pass
    return batch
