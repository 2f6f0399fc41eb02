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


def install() -> list[str]:
    """Patch an importable `strkit`; returns the list of 'module.name' bindings that were replaced."""
    from . import repeats as ours

    patched = []
    for mod_name in _TARGETS:
        mod = importlib.import_module(mod_name)
        for name in _NAMES:
            if hasattr(mod, name):
                _saved.setdefault((mod_name, name), getattr(mod, name))
                setattr(mod, name, getattr(ours, name))
                patched.append(f"{mod_name}.{name}")
    return patched


def uninstall() -> None:
    for (mod_name, name), fn in list(_saved.items()):
        setattr(importlib.import_module(mod_name), name, fn)
        del _saved[(mod_name, name)]
