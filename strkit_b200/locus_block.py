"""Block-mode integration: the reads of a whole block of loci in ONE device call, fed back to the UNCHANGED
per-read bookkeeping of the reference.

What it replaces: the inner loop of the reference's worker (strkit/call/call_sample.py:103-157): for every locus of
a block, call_locus (call_locus.py:974) runs get_ref_repeat_count once (:799-810) and get_repeat_count once per read
(:1148-1155), each a Python -> Rust -> C round trip.  Here a worker makes one pre-pass over its block collecting
the argument tuples, one `BlockSession.run()` (pack -> C ABI -> CUDA), and then runs call_locus exactly as before
with the two names bound to the session's look-ups.

Why this is a drop-in and not a re-implementation of call_locus: both functions are pure in their arguments, so a
session is nothing but their cache, warmed on the GPU.  The only order-dependent part -- the start count of read k
carries an offset computed from the results of reads < k (call_locus.py:1079,1129-1136,1161) -- is replayed on the
device (replay.cuh) in float64, so the keys the session stores ARE the calls the unchanged loop will make; every
later filter of that loop (calc_adj_score / min_read_align_score :1172-1222, terrible-read abort :1224-1250, the
n_read_cn_iters log line :1164) consumes the returned tuple and runs untouched.  A call the pre-pass did not foresee
(different flank slicing, a read skipped upstream) is a cache miss and goes through the per-call path: same answer,
one launch-latency slower.  `hits` / `misses` make that visible.
"""
from __future__ import annotations

import contextlib
import importlib
from typing import Iterable, Sequence

import numpy as np

from . import repeats as _per_call
from .batcher import LocusReads, pack_loci
from .engine import Engine, default_engine
from .repeat_count_params import RepeatCountParams

__all__ = ["BlockSession"]

ReadTuple = tuple  # (est_cn, tr_seq_wc, flank_left_seq_wc, flank_right_seq_wc)


def _rc_key(p: RepeatCountParams) -> tuple:
    return (p.method, p.max_iters, p.initial_local_search_range, p.initial_step_size)


class BlockSession:
    """Collect -> run -> look up.  One session per block of loci per worker; not thread-safe."""

    def __init__(self, rc_params: RepeatCountParams, engine: Engine | None = None, flank_size: int = 70,
                 nibble: bool = True):
        if rc_params.method != "repalign":
            raise NotImplementedError("only rc_method='repalign' runs on the device (repeats.py:57-68)")
        self.rc_params = rc_params
        self.flank_size = flank_size
        self.nibble = nibble
        self._engine = engine
        self._loci: list[LocusReads] = []
        self._refs: list[tuple] = []
        self._reads_cache: dict[tuple, tuple] = {}
        self._ref_cache: dict[tuple, tuple] = {}
        self.hits = self.misses = self.ref_hits = self.ref_misses = 0
        self.n_reads = 0

    # ------------------------------------------------------------------ collect
    def add_locus(self, motif: str, reads: Iterable[ReadTuple]) -> int:
        """The reads of one locus in the order call_locus iterates its segments (:1082): per read
        (get_est_copy_num(), tr_seq_wc, flank_left_seq_wc, flank_right_seq_wc).  Flanks are cut to the flank_size bases
        next to the tract exactly as call_locus.py:1144-1146 does before the call."""
        est, trs, fls, frs = [], [], [], []
        for e, tr, fl, fr in reads:
            est.append(int(e))
            trs.append(tr)
            fls.append(fl[-self.flank_size:])
            frs.append(fr[:self.flank_size])
        self._loci.append(LocusReads(motif, est, trs, fls, frs))
        self.n_reads += len(est)
        return len(self._loci) - 1

    def add_reference(self, start_count: int, tr_seq: str, flank_left_seq: str, flank_right_seq: str, motif: str,
                      ref_size: int, vcf_anchor_size: int, rc_params: RepeatCountParams,
                      respect_coords: bool = False) -> None:
        """The get_ref_repeat_count call of one locus (call_locus.py:799-810), arguments as the reference passes them."""
        self._refs.append((int(start_count), tr_seq, flank_left_seq, flank_right_seq, motif, int(ref_size),
                           int(vcf_anchor_size), rc_params, bool(respect_coords)))

    # ------------------------------------------------------------------ run
    def run(self) -> None:
        eng = self._engine or default_engine()
        pkey = _rc_key(self.rc_params)
        with eng.lock:
            if self.n_reads:
                batch = pack_loci(self._loci, nibble=self.nibble)
                out = eng.count_reads(batch, self.rc_params)
                cache = self._reads_cache
                r = 0
                rows = out.tolist()
                for lr in self._loci:
                    motif = lr.motif
                    for tr, fl, fr in zip(lr.tr_seqs, lr.flank_left_seqs, lr.flank_right_seqs):
                        n, score, n_explored, start = rows[r]
                        cache[(start, tr, fl, fr, motif, pkey)] = ((n, score), n_explored, n - start)
                        r += 1
            # reference windows: one C-ABI call per (vcf_anchor_size, respect_coords) group -- normally one
            groups: dict[tuple, list[int]] = {}
            for i, ref in enumerate(self._refs):
                groups.setdefault((ref[6], ref[8]), []).append(i)
            for (anchor, respect), idx in groups.items():
                refs = [self._refs[i] for i in idx]
                batch = pack_loci([LocusReads(m, [sc], [tr], [fl], [fr]) for sc, tr, fl, fr, m, *_ in refs])
                rc = np.array([[p.max_iters, p.initial_local_search_range, p.initial_step_size] for *_, p, _ in refs],
                              dtype=np.int32)
                out = eng.ref_counts(batch, [x[0] for x in refs], [x[5] for x in refs], rc, anchor, respect).tolist()
                for ref, (cn, score, l_off, r_off, n_off, n_fin, nfl, nfr) in zip(refs, out):
                    sc, tr, fl, fr, m, ref_size, _, p, _ = ref
                    db = f"{fl}{tr}{fr}"
                    self._ref_cache[(sc, tr, fl, fr, m, ref_size, anchor, _rc_key(p), respect)] = (
                        (cn, score), l_off, r_off, (n_off, n_fin), (db[:nfl], db[nfl:len(db) - nfr], db[len(db) - nfr:]))
        self._loci, self._refs, self.n_reads = [], [], 0

    # ------------------------------------------------------------------ look up (drop-in signatures)
    def get_repeat_count(self, start_count: int, tr_seq: str, flank_left_seq: str, flank_right_seq: str, motif: str,
                         rc_params: RepeatCountParams) -> tuple[tuple[int, int], int, int]:
        """strkit.call.repeats.get_repeat_count (repeats.py:47-70)."""
        hit = self._reads_cache.get((start_count, tr_seq, flank_left_seq, flank_right_seq, motif, _rc_key(rc_params)))
        if hit is not None:
            self.hits += 1
            return hit
        self.misses += 1
        return _per_call.get_repeat_count(start_count, tr_seq, flank_left_seq, flank_right_seq, motif, rc_params)

    def get_ref_repeat_count(self, start_count: int, tr_seq: str, flank_left_seq: str, flank_right_seq: str, motif: str,
                             ref_size: int, vcf_anchor_size: int, rc_params: RepeatCountParams,
                             respect_coords: bool = False):
        """strkit.call.repeats.get_ref_repeat_count (repeats.py:73-192)."""
        hit = self._ref_cache.get((start_count, tr_seq, flank_left_seq, flank_right_seq, motif, ref_size, vcf_anchor_size,
                                   _rc_key(rc_params), bool(respect_coords)))
        if hit is not None:
            self.ref_hits += 1
            return hit
        self.ref_misses += 1
        return _per_call.get_ref_repeat_count(start_count, tr_seq, flank_left_seq, flank_right_seq, motif, ref_size,
                                              vcf_anchor_size, rc_params, respect_coords)

    @contextlib.contextmanager
    def installed(self, modules: Sequence[str] = ("strkit.call.repeats", "strkit.call.call_locus")):
        """Bind get_repeat_count / get_ref_repeat_count inside the given modules to this session for the duration of
        the block (call_locus from-imports both names, call_locus.py:32)."""
        saved = []
        try:
            for mod_name in modules:
                mod = importlib.import_module(mod_name)
                for name in ("get_repeat_count", "get_ref_repeat_count"):
                    if hasattr(mod, name):
                        saved.append((mod, name, getattr(mod, name)))
                        setattr(mod, name, getattr(self, name))
            yield self
        finally:
            for mod, name, fn in saved:
                setattr(mod, name, fn)
