"""`strkit_rust_ext.get_repeat_count` with the signature of the PyO3 function the reference calls
(strkit/call/repeats.py:7,58-68):

    get_repeat_count(start_count, tr_seq, flank_left_seq, flank_right_seq, motif, max_iters, local_search_range,
                     step_size, use_shortcuts=False) -> ((best_n, best_score), n_explored, best_n - start_count)

A maintainer who wants the swap at the FFI line itself replaces `from strkit_rust_ext import get_repeat_count`
(repeats.py:7) by `from strkit_b200.strkit_rust_ext_shim import get_repeat_count`; everything above that line,
including the reference's own lru_cache and RepeatCountParams handling, stays as it is.  One C-ABI call
(strk_get_repeat_count) per Python call; for throughput use strkit_b200.locus_block.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from ._native import check, lib
from .engine import default_engine

__all__ = ["get_repeat_count"]

_tls = threading.local()


def get_repeat_count(start_count: int, tr_seq: str, flank_left_seq: str, flank_right_seq: str, motif: str,
                     max_iters: int, local_search_range: int, step_size: int,
                     use_shortcuts: bool = False) -> tuple[tuple[int, int], int, int]:
    if use_shortcuts:
        raise NotImplementedError("use_shortcuts=True is never passed by the reference (repeats.py:67) and its "
                                  "semantics are not in the reference tree")
    out = getattr(_tls, "out", None)
    if out is None:
        out = _tls.out = np.zeros(4, dtype=np.int32)
    tr, fl, fr, mo = (x.encode("ascii") for x in (tr_seq, flank_left_seq, flank_right_seq, motif))
    eng = default_engine()
    with eng.lock:
        check(lib.strk_get_repeat_count(eng._ctx, int(start_count), tr, len(tr), fl, len(fl), fr, len(fr), mo, len(mo),
                                        int(max_iters), int(local_search_range), int(step_size), 0,
                                        C.c_void_p(out.ctypes.data)))
        return (int(out[0]), int(out[1])), int(out[2]), int(out[3])
