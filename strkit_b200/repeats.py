"""Drop-in replacements for strkit.call.repeats.get_repeat_count / get_ref_repeat_count.

Same signatures, argument meaning and return tuples as the reference (strkit/call/repeats.py:47-70 and
:73-192); the work happens in CUDA behind the C ABI.  For throughput use strkit_b200.batcher +
Engine.count_reads (one call per block of loci); these per-call wrappers exist so that the unchanged
call_locus code keeps working when `strkit_b200.install()` rebinds the names.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .batcher import LocusReads, pack_loci
from .engine import default_engine
from .repeat_count_params import RepeatCountParams

__all__ = ["get_repeat_count", "get_ref_repeat_count"]


@lru_cache(maxsize=512)
def get_repeat_count(
    start_count: int,
    tr_seq: str,
    flank_left_seq: str,
    flank_right_seq: str,
    motif: str,
    rc_params: RepeatCountParams,
) -> tuple[tuple[int, int], int, int]:
    """returns: (best size, best score), n_explored, best size - start count   (repeats.py:55-56)"""
    if rc_params.method != "repalign":
        raise NotImplementedError("only rc_method='repalign' is implemented (repeats.py:57-68); 'comp' is the "
                                  "experimental composition counter and stays with the reference")
    batch = pack_loci([LocusReads(motif, [start_count], [tr_seq], [flank_left_seq], [flank_right_seq])])
    eng = default_engine()
    with eng.lock:
        out = eng.count_reads(batch, rc_params)
    n, score, n_explored, start = (int(v) for v in out[0])
    return (n, score), n_explored, n - start


def get_ref_repeat_count(
    start_count: int,
    tr_seq: str,
    flank_left_seq: str,
    flank_right_seq: str,
    motif: str,
    ref_size: int,
    vcf_anchor_size: int,
    rc_params: RepeatCountParams,
    respect_coords: bool = False,
) -> tuple[tuple[int, int], int, int, tuple[int, int], tuple[str, str, str]]:
    """Reference repeat count with boundary extension (repeats.py:73-192)."""
    batch = pack_loci([LocusReads(motif, [start_count], [tr_seq], [flank_left_seq], [flank_right_seq])])
    rc = np.array([[rc_params.max_iters, rc_params.initial_local_search_range, rc_params.initial_step_size]],
                  dtype=np.int32)
    eng = default_engine()
    with eng.lock:
        out = eng.ref_counts(batch, [start_count], [ref_size], rc, vcf_anchor_size, respect_coords)[0]
    cn, score, l_off, r_off, n_off, n_fin, nfl, nfr = (int(v) for v in out)
    db = f"{flank_left_seq}{tr_seq}{flank_right_seq}"
    # the reference returns the adjusted slices without upper-casing them (repeats.py:171-176,190-192)
    new_fl, new_tr, new_fr = db[:nfl], db[nfl:len(db) - nfr], db[len(db) - nfr:]
    return (cn, score), l_off, r_off, (n_off, n_fin), (new_fl, new_tr, new_fr)
