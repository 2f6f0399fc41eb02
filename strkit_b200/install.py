"""Rebind STRkit's hot-path functions to the B200 implementations.

call_locus binds the names with `from .repeats import get_repeat_count, get_ref_repeat_count`
(strkit/call/call_locus.py:32), so both strkit.call.repeats and strkit.call.call_locus are patched.
"""
from __future__ import annotations

import importlib

__all__ = ["install", "uninstall"]

_saved: dict[tuple[str, str], object] = {}
_TARGETS = ("strkit.call.repeats", "strkit.call.call_locus")
_NAMES = ("get_repeat_count", "get_ref_repeat_count")
_ALLELE_TARGETS = ("strkit.call.allele", "strkit.call.call_locus")  # call_locus.py:28 from-imports call_alleles


def install(alleles: bool = False) -> list[str]:
    """Patch an importable `strkit`; returns the list of 'module.name' bindings that were replaced.
    alleles=True also rebinds call_alleles (allele.py:176) to the GPU bootstrap / GMM caller: its random streams
    are not numpy's, so calls agree with the reference statistically rather than bit for bit (opt-in)."""
    from . import alleles as ours_alleles
    from . import repeats as ours

    patched = []
    plan = [(m, n, getattr(ours, n)) for m in _TARGETS for n in _NAMES]
    if alleles:
        plan += [(m, "call_alleles", ours_alleles.call_alleles) for m in _ALLELE_TARGETS]
    for mod_name, name, fn in plan:
        mod = importlib.import_module(mod_name)
        if hasattr(mod, name):
            _saved.setdefault((mod_name, name), getattr(mod, name))
            setattr(mod, name, fn)
            patched.append(f"{mod_name}.{name}")
    return patched


def uninstall() -> None:
    for (mod_name, name), fn in list(_saved.items()):
        setattr(importlib.import_module(mod_name), name, fn)
        del _saved[(mod_name, name)]
