#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's own strkit/call/repeats.py.

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/gen_golden.py

What is real and what is stubbed:

* REAL reference code, imported unmodified from /root/reference: strkit.call.repeats
  (get_ref_repeat_count, score_ref_boundaries, the get_repeat_count dispatcher),
  strkit.call.align_matrix (the matrix construction), strkit.iupac,
  strkit.call.repeat_count_params.
* STUBBED because the packages are not installable here (no network, no Rust):
  - `parasail`: a small numpy re-statement of matrix_create / profile_create_sat /
    sg*_scan_profile_sat (exact integer DP).  Independent of oracle/strk_oracle.c,
    so the two restatements check each other.
  - `strkit_rust_ext.get_repeat_count`: the hill-climb of repeats.py:100-156 with one
    score per size (plain "sg" alignment) -- a restatement, marked as such below.
  - importlib.metadata.version("strkit") (strkit/__init__.py:7).

So the committed vectors PIN the in-tree control flow of get_ref_repeat_count
(hill-climb order, tie-breaks, offset rules, rounding) and the matrix; the DP
semantics underneath remain a restatement ("parity unpinned").
"""
from __future__ import annotations

import importlib.metadata
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = "/root/reference"
OUT = Path(__file__).parent


# ------------------------------------------------------------------ parasail stub
class _Matrix:
    def __init__(self, alphabet: str, match: int, mismatch: int):
        n = len(alphabet)
        self.size = n + 1
        self.m = np.zeros((n + 1, n + 1), dtype=np.int32)
        self.m[:n, :n] = mismatch
        self.m[np.arange(n), np.arange(n)] = match
        self.mapper = np.full(256, n, dtype=np.int64)
        for i, c in enumerate(alphabet):
            self.mapper[ord(c.upper())] = i
            self.mapper[ord(c.lower())] = i

    def __setitem__(self, key, value):
        self.m[key] = value

    def __getitem__(self, key):
        return self.m[key]

    def encode(self, s: str) -> np.ndarray:
        return self.mapper[np.frombuffer(s.encode("ascii"), dtype=np.uint8)]


class _Profile:
    def __init__(self, s1: str, matrix: _Matrix):
        self.s1 = s1
        self.matrix = matrix


class _Result:
    def __init__(self, score, end_query, end_ref):
        self.score, self.end_query, self.end_ref = int(score), int(end_query), int(end_ref)


def _sg(profile: _Profile, s2: str, open_: int, ext: int, s1_beg, s1_end, s2_beg, s2_end) -> _Result:
    assert open_ == ext, "stub only restates the linear-gap case the reference uses (5, 5)"
    g = open_
    mat = profile.matrix
    a = mat.encode(profile.s1)
    b = mat.encode(s2)
    n1, n2 = len(a), len(b)
    assert n1 > 0 and n2 > 0
    jj = np.arange(n2 + 1, dtype=np.int64)
    prev = np.zeros(n2 + 1, dtype=np.int64) if s2_beg else -g * jj
    best, bq, br = None, n1 - 1, n2 - 1
    for i in range(1, n1 + 1):
        sub = mat.m[a[i - 1]][b].astype(np.int64)
        t = np.empty(n2 + 1, dtype=np.int64)
        t[0] = 0 if s1_beg else -g * i
        t[1:] = np.maximum(prev[:-1] + sub, prev[1:] - g)
        cur = np.maximum.accumulate(t + g * jj) - g * jj  # horizontal gaps
        prev = cur
        if s1_end and (best is None or cur[n2] > best):
            best, bq, br = cur[n2], i - 1, n2 - 1
    if s2_end:
        for j in range(1, n2 + 1):
            if best is None or prev[j] > best:
                best, bq, br = prev[j], n1 - 1, j - 1
    if best is None or prev[n2] > best or (not s1_end and not s2_end):
        best, bq, br = prev[n2], n1 - 1, n2 - 1
    return _Result(best, bq, br)


parasail = types.ModuleType("parasail")
parasail.matrix_create = _Matrix
parasail.Profile = _Profile
parasail.profile_create_sat = _Profile
parasail.sg_qe_scan_profile_sat = lambda p, s2, o, e: _sg(p, s2, o, e, False, True, False, False)
parasail.sg_scan_profile_sat = lambda p, s2, o, e: _sg(p, s2, o, e, True, True, True, True)
parasail.sg_flags = _sg
sys.modules["parasail"] = parasail


# ------------------------------------------------------------------ strkit_rust_ext stub (RESTATED)
def _rust_get_repeat_count(start_count, tr_seq, fl, fr, motif, max_iters, local_search_range, step_size,
                           use_shortcuts=False):
    from strkit.call.align_matrix import dna_matrix, indel_penalty

    prof = parasail.profile_create_sat(f"{fl}{tr_seq}{fr}", dna_matrix)
    to_explore = [(start_count - step_size, -1), (start_count + step_size, 1), (start_count, 0)]
    sizes_and_scores: dict[int, int] = {}
    n_explored = 0
    while to_explore and n_explored < max_iters:
        size, direction = to_explore.pop()
        if size < 0:
            continue
        szs = []
        start_size = max(size - (local_search_range if (direction < 1 or step_size > local_search_range) else 0), 0)
        end_size = size + (local_search_range if (direction > -1 or step_size > local_search_range) else 0)
        for i in range(start_size, end_size + 1):
            if i not in sizes_and_scores:
                r = parasail.sg_scan_profile_sat(prof, f"{fl}{motif * i}{fr}", indel_penalty, indel_penalty)
                sizes_and_scores[i] = r.score
                n_explored += 1
            szs.append((i, sizes_and_scores[i]))
        mv = max(szs, key=lambda x: x[1])
        if mv[0] > size and (new_rc := mv[0] + step_size) not in sizes_and_scores and new_rc >= 0:
            to_explore.append((new_rc, 1))
        if mv[0] < size and (new_rc := mv[0] - step_size) not in sizes_and_scores and new_rc >= 0:
            to_explore.append((new_rc, -1))
    res = max(sizes_and_scores.items(), key=lambda x: x[1])
    return res, n_explored, res[0] - start_count


rust = types.ModuleType("strkit_rust_ext")
rust.get_repeat_count = _rust_get_repeat_count
rust.get_repeat_count_compostr = lambda tr_seq, motif: (0, 0)
sys.modules["strkit_rust_ext"] = rust

_orig_version = importlib.metadata.version
importlib.metadata.version = lambda name: "0.25.0a5" if name == "strkit" else _orig_version(name)
sys.path.insert(0, REF)

from strkit.call import repeats as ref_repeats  # noqa: E402  (REAL reference code)
from strkit.call.align_matrix import dna_bases_str, dna_matrix  # noqa: E402
from strkit.call.repeat_count_params import RepeatCountParams, get_reference_rc_params  # noqa: E402


# ------------------------------------------------------------------ case generators
def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n))


def mutate(rng, s, sub=0.01, ins=0.01, dele=0.01, alphabet="ACGT"):
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


IUPAC_MOTIFS = ["RAAAT", "AARRG", "GCN", "CNG", "YTC", "GGCCTS", "WCA", "AAAG", "CAG", "AT", "TTTCA", "GAAGGA",
                "CCCCGGCCCCGG", "A", "NGC", "BDHV"]


def gen_ref_case(rng, idx):
    motif = IUPAC_MOTIFS[idx % len(IUPAC_MOTIFS)] if idx % 3 == 0 else rand_seq(rng, int(rng.integers(1, 7)))
    m = len(motif)
    k = int(rng.integers(3, 40))
    concrete = "".join(rng.choice(list("ACGT")) if c not in "ACGT" else c for c in motif)
    tr = mutate(rng, concrete * k, 0.02, 0.01, 0.01)
    flank = int(rng.choice([20, 30, 70]))
    fl, fr = rand_seq(rng, flank), rand_seq(rng, flank)
    # let the repeat spill into the flanks in some cases -> non-zero offsets
    if idx % 2 == 0:
        el = int(rng.integers(0, 3 * m + 1))
        er = int(rng.integers(0, 3 * m + 1))
        rep = concrete * 6
        if el:
            fl = fl[: flank - el] + rep[len(rep) - el:]
        if er:
            fr = rep[:er] + fr[er:]
    if idx % 5 == 0:
        fl, tr, fr = fl.lower(), tr.lower(), fr.lower()  # soft-masked reference
    if idx % 7 == 0 and len(tr) > 4:
        p = int(rng.integers(0, len(tr) - 2))
        tr = tr[:p] + "NN" + tr[p + 2:]
    ref_size = len(tr)
    est = round(ref_size / m)
    start = max(0, est + int(rng.integers(-4, 5))) if idx % 4 == 0 else est
    if idx % 11 == 0:
        params = RepeatCountParams("repalign", int(rng.integers(3, 12)), 3, 1)  # tight iteration budget
    elif idx % 13 == 0:
        params = RepeatCountParams("repalign", 50, 1, 4)  # step > range
    elif idx % 17 == 0:
        params = RepeatCountParams("repalign", 200, 3, 3)
    else:
        params = get_reference_rc_params("repalign", est, 250)
    return dict(start_count=start, tr_seq=tr, flank_left_seq=fl, flank_right_seq=fr, motif=motif, ref_size=ref_size,
                vcf_anchor_size=5, max_iters=params.max_iters, local_search_range=params.initial_local_search_range,
                step_size=params.initial_step_size, respect_coords=bool(idx % 19 == 0))


def main():
    rng = np.random.default_rng(20261018)
    golden = {"alphabet": dna_bases_str,
              "matrix": [[int(dna_matrix[i, j]) for j in range(17)] for i in range(17)],
              "ref": [], "boundaries": [], "read_restated": [], "sg": []}

    for idx in range(160):
        c = gen_ref_case(rng, idx)
        params = RepeatCountParams("repalign", c["max_iters"], c["local_search_range"], c["step_size"])
        ref_repeats.get_repeat_count.cache_clear()
        res = ref_repeats.get_ref_repeat_count(
            c["start_count"], c["tr_seq"], c["flank_left_seq"], c["flank_right_seq"], c["motif"], c["ref_size"],
            c["vcf_anchor_size"], params, c["respect_coords"])
        (cn, score), l_off, r_off, (n_off, n_fin), (fl2, tr2, fr2) = res
        c["expect"] = dict(cn=int(cn), score=int(score), l_offset=int(l_off), r_offset=int(r_off),
                           n_offset_scores=int(n_off), n_iters_final=int(n_fin), fl=fl2, tr=tr2, fr=fr2)
        golden["ref"].append(c)

    for idx in range(60):
        c = gen_ref_case(rng, idx)
        db = c["flank_left_seq"] + c["tr_seq"] + c["flank_right_seq"]
        prof = parasail.profile_create_sat(db, dna_matrix)
        prof_r = parasail.profile_create_sat(db[::-1], dna_matrix)
        n = max(0, c["start_count"] + int(rng.integers(-3, 4)))
        if n == 0 and not c["flank_left_seq"]:
            n = 1
        (fs, ra), (rs, la) = ref_repeats.score_ref_boundaries(
            prof, prof_r, c["motif"] * n, c["flank_left_seq"], c["flank_right_seq"], c["ref_size"])
        golden["boundaries"].append(dict(tr_seq=c["tr_seq"], flank_left_seq=c["flank_left_seq"],
                                         flank_right_seq=c["flank_right_seq"], motif=c["motif"], n=n,
                                         ref_size=c["ref_size"], expect=[int(fs), int(ra), int(rs), int(la)]))

    # read path: dispatcher is the reference's, the search body is the restated stub
    for idx in range(120):
        c = gen_ref_case(rng, idx + 1000)
        tr = mutate(rng, c["tr_seq"].upper(), 0.01, 0.02, 0.02)
        if idx % 3 == 0:  # low-quality wildcards (align_matrix.py:23-24)
            tr = "".join("X" if rng.random() < 0.03 else ch for ch in tr)
        start = max(0, round(len(tr) / len(c["motif"])) + int(rng.integers(-6, 7)) * (idx % 2))
        params = RepeatCountParams("repalign", 50 if idx % 9 else 7, 3, 1 if idx % 10 else 2)
        ref_repeats.get_repeat_count.cache_clear()
        (n, s), n_exp, delta = ref_repeats.get_repeat_count(
            start, tr, c["flank_left_seq"].upper(), c["flank_right_seq"].upper(), c["motif"], params)
        golden["read_restated"].append(dict(start_count=start, tr_seq=tr, flank_left_seq=c["flank_left_seq"].upper(),
                                            flank_right_seq=c["flank_right_seq"].upper(), motif=c["motif"],
                                            max_iters=params.max_iters, local_search_range=3,
                                            step_size=params.initial_step_size,
                                            expect=[int(n), int(s), int(n_exp), int(delta)]))

    # raw alignments in all 16 free-end modes (numpy restatement vs. the C oracle)
    alpha = "ACGTRYSWKMBDHVNXacgtn-"
    for idx in range(200):
        n1, n2 = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        base = rand_seq(rng, max(n1, n2) + 5, "ACGT")
        s1 = mutate(rng, base, 0.1, 0.1, 0.1)[:n1] or "A"
        s2 = mutate(rng, base, 0.1, 0.1, 0.1)[:n2] or "C"
        if idx % 4 == 0:
            s1 = "".join(rng.choice(list(alpha)) if rng.random() < 0.2 else ch for ch in s1)
            s2 = "".join(rng.choice(list(alpha)) if rng.random() < 0.2 else ch for ch in s2)
        flags = idx % 16
        r = _sg(parasail.profile_create_sat(s1, dna_matrix), s2, 5, 5, bool(flags & 1), bool(flags & 2),
                bool(flags & 4), bool(flags & 8))
        golden["sg"].append(dict(s1=s1, s2=s2, flags=flags, expect=[r.score, r.end_query, r.end_ref]))

    with open(OUT / "repeats_golden.json", "w") as fh:
        json.dump(golden, fh, indent=0, sort_keys=True)
    print({k: len(v) for k, v in golden.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
