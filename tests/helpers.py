"""Shared helpers for the parity tests (oracle side = tests/oracle_lib.py, the checker)."""
from __future__ import annotations

import numpy as np

from strkit_b200.batcher import LocusReads, ReadBatch, pack_loci


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n)) if n > 0 else ""


def mutate(rng, s, sub=0.02, ins=0.02, dele=0.02, alphabet="ACGT"):
    out = []
    for c in s:
        u = rng.random()
        if u < dele:
            continue
        if u < dele + sub:
            c = rng.choice(list(alphabet))
        out.append(c)
        if rng.random() < ins:
            out.append(rng.choice(list(alphabet)))
    return "".join(out)


def random_families(rng, n, max_k=12, flank_choices=(0, 1, 5, 20, 70), weird=True):
    """Small random read families: list of (motif, tr, fl, fr)."""
    fams = []
    alpha_weird = "ACGTRYSWKMBDHVNXacgtn-*"
    for i in range(n):
        m = int(rng.integers(1, 8))
        motif = rand_seq(rng, m, "ACGTRYN" if (weird and i % 5 == 0) else "ACGT")
        k = int(rng.integers(0, max_k + 1))
        concrete = "".join(c if c in "ACGT" else rng.choice(list("ACGT")) for c in motif)
        tr = mutate(rng, concrete * k, 0.05, 0.03, 0.03)
        fl = rand_seq(rng, int(rng.choice(flank_choices)))
        fr = rand_seq(rng, int(rng.choice(flank_choices)))
        if weird and i % 7 == 0:
            tr = "".join(rng.choice(list(alpha_weird)) if rng.random() < 0.15 else ch for ch in tr)
            fl = "".join(rng.choice(list(alpha_weird)) if rng.random() < 0.1 else ch for ch in fl)
        if len(fl) + len(tr) + len(fr) == 0:
            tr = "A"
        fams.append((motif, tr, fl, fr))
    return fams


def families_to_batch(fams, est=None) -> ReadBatch:
    loci = []
    for i, (motif, tr, fl, fr) in enumerate(fams):
        e = round(len(tr) / len(motif)) if est is None else est[i]
        loci.append(LocusReads(motif, [e], [tr], [fl], [fr]))
    return pack_loci(loci)


def oracle_tables(oracle, fams, n_lo, n_hi, flags):
    out = []
    for (motif, tr, fl, fr), lo, hi in zip(fams, n_lo, n_hi):
        out.extend(oracle.score_candidate(tr, fl, fr, motif, n, flags) for n in range(lo, hi + 1))
    return np.asarray(out, dtype=np.int32)
