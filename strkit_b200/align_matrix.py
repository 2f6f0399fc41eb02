"""Scoring constants of the repeat-count alignments.

Same names and values as the reference's strkit/call/align_matrix.py:15-44 (match 2, mismatch 7,
indel 5; 16-letter alphabet ACGT + IUPAC codes + the low-quality wildcard X) and strkit/iupac.py:9-21.
`dna_matrix` here is a plain 17x17 numpy int8 array (row/column 16 = parasail's wildcard for bytes
outside the alphabet) that is handed to the device at context creation.
"""
from __future__ import annotations

import numpy as np

__all__ = ["dna_codes", "match_score", "mismatch_penalty", "indel_penalty", "dna_bases", "dna_bases_str", "dna_matrix"]

match_score: int = 2
mismatch_penalty: int = 7
indel_penalty: int = 5

# order of the reference's IUPAC_NUCLEOTIDE_CODES dict (iupac.py:9-21); "D" really lists A, C, T there
_IUPAC: dict[str, str] = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "ACT",
                          "H": "ACT", "V": "ACG", "N": "ACGT"}

dna_bases_str: str = "ACGT" + "".join(_IUPAC) + "X"
dna_bases: dict[str, int] = {b: i for i, b in enumerate(dna_bases_str)}
dna_codes: dict[str, tuple[str, ...]] = {**{k: tuple(v) for k, v in _IUPAC.items()}, "X": ("A", "C", "G", "T")}


def _make_matrix() -> np.ndarray:
    n = len(dna_bases_str)
    mat = np.zeros((n + 1, n + 1), dtype=np.int8)  # last row/column: unknown bytes score 0
    mat[:n, :n] = -mismatch_penalty
    mat[np.arange(n), np.arange(n)] = match_score
    for code, bases in dna_codes.items():
        v = match_score if code != "X" else 0
        for b in bases:
            mat[dna_bases[code], dna_bases[b]] = v
            mat[dna_bases[b], dna_bases[code]] = v
    mat.setflags(write=False)
    return mat


dna_matrix: np.ndarray = _make_matrix()
