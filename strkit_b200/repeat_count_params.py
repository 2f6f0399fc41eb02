"""Search hyper-parameters; same interface as the reference's strkit/call/repeat_count_params.py."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal

__all__ = ["RepeatCountMethod", "RepeatCountParams", "get_reference_rc_params"]

RepeatCountMethod = Literal["repalign", "comp"]


@dataclass(frozen=True)
class RepeatCountParams:
    method: str
    max_iters: int
    initial_local_search_range: int
    initial_step_size: int


def get_reference_rc_params(method: str, ref_est_cn: int, default_ref_max_iters: int) -> RepeatCountParams:
    """Tiers of repeat_count_params.py:17-42: large reference tracts search in bigger steps, fewer iterations."""
    max_iters, step, search_range = default_ref_max_iters, 1, 3
    if ref_est_cn >= 2000:
        max_iters, step, search_range = 50, 15, 1
    elif ref_est_cn >= 1000:
        max_iters, step = 150, 5
    elif ref_est_cn >= 200:
        max_iters, step = 200, 3
    return RepeatCountParams(method=method, max_iters=max_iters, initial_local_search_range=search_range,
                             initial_step_size=step)
