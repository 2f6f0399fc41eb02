"""The CPU oracle against (i) vectors produced by executing the reference's own repeats.py
(tests/golden/gen_golden.py) and (ii) hand-derivable known answers."""
import numpy as np
import pytest

from tests.oracle_lib import MODE_SG, MODE_SG_QE


def test_matrix_matches_reference_construction(oracle, golden):
    # golden["matrix"] was built by the reference's align_matrix._create_dna_matrix (align_matrix.py:33-44)
    assert golden["alphabet"] == "ACGTRYSWKMBDHVNX"
    assert oracle.matrix.reshape(17, 17).tolist() == golden["matrix"]


def test_matrix_known_entries(oracle):
    m = oracle.matrix.reshape(17, 17)
    s = oracle.symbol
    assert m[s("A"), s("A")] == 2 and m[s("A"), s("C")] == -7
    assert m[s("X"), s("A")] == 0 and m[s("T"), s("X")] == 0 and m[s("X"), s("X")] == 2
    assert m[s("N"), s("G")] == 2 and m[s("N"), s("N")] == 2
    assert m[s("R"), s("A")] == 2 and m[s("R"), s("G")] == 2 and m[s("R"), s("C")] == -7
    assert m[s("R"), s("N")] == -7  # code-vs-code pairs are not overridden
    assert m[s("D"), s("G")] == -7 and m[s("D"), s("C")] == 2  # iupac.py:17 quirk: D == H
    assert m[s("a"), s("A")] == 2  # case-insensitive mapper
    assert s("-") == 16 and (m[16, :] == 0).all() and (m[:, 16] == 0).all()


def test_sg_align_all_modes(oracle, golden):
    for c in golden["sg"]:
        assert list(oracle.sg_align(c["s1"], c["s2"], c["flags"])) == c["expect"], c


def test_sg_affine_equals_linear_when_open_eq_extend(oracle):
    # the only gap setting the reference uses is (5, 5) (repeats.py:33,40)
    assert oracle.sg_align("ACGTACGT", "ACGACGT", 0)[0] == 2 * 7 - 5
    assert oracle.sg_align("ACGTACGT", "ACGT", 0)[0] == 8 - 20
    assert oracle.sg_align("ACGTACGT", "ACGT", MODE_SG_QE) == (8, 3, 3)
    # affine: one gap of 2 with open 7 / extend 1 costs 8
    assert oracle.sg_align("AACCGGTT", "AACCTT", 0, 7, 1)[0] == 12 - 8


def test_score_ref_boundaries_golden(oracle, golden):
    for c in golden["boundaries"]:
        (fs, ra), (rs, la) = oracle.score_ref_boundaries(c["tr_seq"], c["flank_left_seq"], c["flank_right_seq"],
                                                         c["motif"], c["n"], c["ref_size"])
        assert [fs, ra, rs, la] == c["expect"], c


def test_get_ref_repeat_count_golden(oracle, golden):
    for c in golden["ref"]:
        e = c["expect"]
        res = oracle.get_ref_repeat_count(c["start_count"], c["tr_seq"], c["flank_left_seq"], c["flank_right_seq"],
                                          c["motif"], c["ref_size"], c["vcf_anchor_size"], c["max_iters"],
                                          c["local_search_range"], c["step_size"], c["respect_coords"])
        (cn, score), lo, ro, (n_off, n_fin), (fl2, tr2, fr2) = res
        assert (cn, score, lo, ro, n_off, n_fin) == (e["cn"], e["score"], e["l_offset"], e["r_offset"],
                                                     e["n_offset_scores"], e["n_iters_final"]), c
        # the reference upper-cases only what it passes on, not what it returns (repeats.py:183,190-192)
        assert (fl2, tr2, fr2) == (e["fl"], e["tr"], e["fr"])


def test_get_repeat_count_restated_golden(oracle, golden):
    for c in golden["read_restated"]:
        (n, s), n_exp, delta = oracle.get_repeat_count(c["start_count"], c["tr_seq"], c["flank_left_seq"],
                                                       c["flank_right_seq"], c["motif"], c["max_iters"],
                                                       c["local_search_range"], c["step_size"])
        assert [n, s, n_exp, delta] == c["expect"], c


@pytest.mark.parametrize("flags", [0, MODE_SG_QE, 5, 10, MODE_SG])
def test_known_answers_exact_tract(oracle, flags):
    """db = fl + motif*k + fr with random non-periodic flanks:
    score(k) = 2L, score(k+1) = 2L - 5m, score(k-1) = 2L - 7m, in every free-end mode."""
    rng = np.random.default_rng(7 + flags)
    for _ in range(20):
        m = int(rng.integers(2, 7))
        while True:
            motif = "".join(rng.choice(list("ACGT"), size=m))
            if len(set(motif)) > 1 and all(motif != motif[p:] + motif[:p] for p in range(1, m)):
                break
        k = int(rng.integers(8, 30))
        while True:
            fl = "".join(rng.choice(list("ACGT"), size=70))
            fr = "".join(rng.choice(list("ACGT"), size=70))
            if fl[-m:] != motif and fr[:m] != motif:
                break
        tr = motif * k
        L = 140 + m * k
        assert oracle.score_candidate(tr, fl, fr, motif, k, flags) == 2 * L
        assert oracle.score_candidate(tr, fl, fr, motif, k + 1, flags) == 2 * L - 5 * m
        if flags == 0:
            assert oracle.score_candidate(tr, fl, fr, motif, k - 1, flags) == 2 * L - 7 * m
        (n, s), n_explored, delta = oracle.get_repeat_count(k, tr, fl, fr, motif, 50, 3, 1, flags)
        assert (n, s, n_explored, delta) == (k, 2 * L, 9, 0)


def test_hill_climb_budget_and_start_dependence(oracle):
    """The result is not an arg-max: it depends on start_count and max_iters (SURVEY 8a dry run)."""
    rng = np.random.default_rng(3)
    motif = "CAG"
    fl = "".join(rng.choice(list("ACGT"), size=70))
    fr = "".join(rng.choice(list("ACGT"), size=70))
    k = 20
    tr = motif * k
    (n, _), n_exp, delta = oracle.get_repeat_count(k + 10, tr, fl, fr, motif, 50, 3, 1)
    assert (n, n_exp, delta) == (k, 16, -10)
    (n, _), n_exp, _ = oracle.get_repeat_count(k + 60, tr, fl, fr, motif, 50, 3, 1)
    assert n_exp == 51 and n == k + 13  # budget exhausted before reaching the optimum (sizes 83..33 scored)
    (n, _), n_exp, _ = oracle.get_repeat_count(2, tr, fl, fr, motif, 50, 3, 1)
    assert n == k


def test_read_loop_carries_offset(oracle):
    """call_locus.py:1129-1161: the start guess of read k uses the offset fraction of reads < k."""
    rng = np.random.default_rng(11)
    motif = b"CAG"
    fl = bytes(rng.choice(list(b"ACGT"), size=70).astype(np.uint8))
    fr = bytes(rng.choice(list(b"ACGT"), size=70).astype(np.uint8))
    reads = [fl + motif * k + fr for k in (20, 20, 21, 20)]
    arena = np.frombuffer(b"".join(reads) + motif, dtype=np.uint8)
    lens = np.array([[70, len(r) - 140, 70] for r in reads], dtype=np.int32)
    seq_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])[:-1]]).astype(np.uint64)
    est = np.array([22, 22, 23, 22], dtype=np.int32)  # estimates biased by +2
    out, cells = oracle.count_loci(arena, seq_off, lens, est, np.array([0, 4]), np.array([sum(map(len, reads))]),
                                   np.array([3]))
    assert out[:, 0].tolist() == [20, 20, 21, 20]
    # read 0 starts at 22; the -2/20 fraction moves later starts: round(-0.1*22) = -2
    assert out[:, 3].tolist() == [22, 20, 21, 20]
    assert cells > 0


def test_locus_memo_equals_per_call_search(oracle):
    """The batch loop answers identical calls of a locus from a memo (the reference's lru_cache, repeats.py:47);
    every row must equal the per-call search replayed here with the carried offset of call_locus.py:1129-1161."""
    from strkit_b200 import synth

    b = synth.generate(synth.CONFIGS[2], 60, seed=17).to_host()   # HiFi-like: a third of the reads are duplicates
    out, _ = oracle.count_loci(b.arena, b.seq_off, b.lens, b.est_cn, b.read_begin, b.motif_off, b.motif_len, n_threads=4)
    arena = b.arena.tobytes()
    dup = 0
    for l in range(b.n_loci):
        motif = arena[int(b.motif_off[l]):int(b.motif_off[l]) + int(b.motif_len[l])].decode()
        frac, seen = 0.0, set()
        for r in range(int(b.read_begin[l]), int(b.read_begin[l + 1])):
            o, (nfl, ntr, nfr) = int(b.seq_off[r]), (int(v) for v in b.lens[r])
            fl, tr, fr = (arena[o:o + nfl].decode(), arena[o + nfl:o + nfl + ntr].decode(),
                          arena[o + nfl + ntr:o + nfl + ntr + nfr].decode())
            start = int(b.est_cn[r])
            off = round(frac * start)
            if off < -start:
                frac = 0.0
            else:
                start += off
            dup += (start, fl, tr, fr) in seen
            seen.add((start, fl, tr, fr))
            (n, score), n_explored, delta = oracle.get_repeat_count(start, tr, fl, fr, motif, 50, 3, 1)
            assert out[r].tolist() == [n, score, n_explored, start], (l, r)
            frac += delta / max(n, 1)
    assert dup > 0.2 * b.n_reads  # the memo was exercised


def _py_search(score, start, max_iters, rng_, step, policy):
    """Independent pure-Python statement of the single-score search with the range-narrowing hypotheses
    (flag 4: range -> 1 after the first window; flag 8: range halves after every window); policy 0 is
    repeats.py:100-156 with one score per size."""
    to_explore = [(start - step, -1), (start + step, 1), (start, 0)]
    seen: dict[int, int] = {}
    n = 0
    while to_explore and n < max_iters:
        size, d = to_explore.pop()
        if size < 0:
            continue
        skip = step > rng_
        lo = max(size - (rng_ if (d < 1 or skip) else 0), 0)
        hi = size + (rng_ if (d > -1 or skip) else 0)
        if policy & 4 and rng_ > 1:
            rng_ = 1
        if policy & 8 and rng_ > 1:
            rng_ //= 2
        szs = []
        for i in range(lo, hi + 1):
            if i not in seen:
                seen[i] = score(i)
                n += 1
            szs.append((i, seen[i]))
        mv = max(szs, key=lambda x: x[1])
        if mv[0] > size and (new := mv[0] + step) not in seen and new >= 0:
            to_explore.append((new, 1))
        if mv[0] < size and (new := mv[0] - step) not in seen and new >= 0:
            to_explore.append((new, -1))
    best = max(seen.items(), key=lambda x: x[1])
    return best, n, best[0] - start


@pytest.mark.parametrize("policy", [0, 4, 8])
def test_search_policy_switches_match_python_statement(oracle, policy):
    """The narrowing switches of the oracle equal an independent Python statement; a perfect start scores 9 sizes with the
    in-tree search and 7 with either narrowing hypothesis (range 3, step 1)."""
    rng = np.random.default_rng(5 + policy)
    motif, k = "CAG", 20
    fl = "".join(rng.choice(list("ACGT"), size=70))
    fr = "".join(rng.choice(list("ACGT"), size=70))
    tr = motif * k
    (n, _), n_exp, _ = oracle.get_repeat_count(k, tr, fl, fr, motif, 50, 3, 1, tie_flags=policy)
    assert (n, n_exp) == (k, 9 if policy == 0 else 7)
    for start, mi, r_, st in ((k + 10, 50, 3, 1), (k - 9, 50, 3, 1), (k + 40, 30, 7, 2), (k + 5, 50, 8, 1), (k + 20, 50, 1, 4),
                              (3, 12, 2, 3), (k + 60, 50, 3, 1)):
        tr2 = tr[:31] + "X" + tr[32:]
        want = _py_search(lambda i: oracle.score_candidate(tr2, fl, fr, motif, i), start, mi, r_, st, policy)
        got = oracle.get_repeat_count(start, tr2, fl, fr, motif, mi, r_, st, tie_flags=policy)
        assert got == want, (start, mi, r_, st)


def test_simd_baseline_kernel_equals_scalar_restatement(oracle, golden):
    """The AVX2 scan kernel that the timed CPU baseline uses (16-bit lanes, striped profile, the layout of parasail's
    *_scan_profile kernels) returns exactly what the scalar restatement returns: raw alignments in all 16 free-end
    modes (score, end_query, end_ref), the golden get_ref_repeat_count vectors, and a batch of HiFi-like loci."""
    from strkit_b200 import synth

    assert oracle.have_simd(), "oracle built without AVX2: the CPU baseline would silently be the scalar port"
    rng = np.random.default_rng(12)
    for it in range(1500):
        n1, n2 = int(rng.integers(1, 420)), int(rng.integers(1, 420))
        alpha = "ACGT" if it % 3 else "ACGTRYSWKMBDHVNXacgtn-"
        s1 = "".join(rng.choice(list(alpha), size=n1))
        s2 = "".join(rng.choice(list(alpha), size=n2))
        if it % 2:  # related sequences: long diagonal runs, gaps
            s2 = "".join(c for c in s1[int(rng.integers(0, n1)):] if rng.random() > 0.04) or "A"
        f = int(rng.integers(0, 16))
        assert oracle.sg_align(s1, s2, f, simd=True) == oracle.sg_align(s1, s2, f), (it, n1, n2, f)
    long1, long2 = "ACGTT" * 1500, "ACGTT" * 1400 + "GG"          # beyond 16-bit lanes: falls back, same answer
    assert oracle.sg_align(long1, long2, 15, simd=True) == oracle.sg_align(long1, long2, 15)
    prev = oracle.set_simd(True)
    try:
        for c in golden["ref"]:
            got = oracle.get_ref_repeat_count(c["start_count"], c["tr_seq"], c["flank_left_seq"], c["flank_right_seq"],
                                              c["motif"], c["ref_size"], c["vcf_anchor_size"], c["max_iters"],
                                              c["local_search_range"], c["step_size"], c["respect_coords"])
            e = c["expect"]
            assert (got[0][0], got[0][1], got[1], got[2], got[3][0], got[3][1]) == (
                e["cn"], e["score"], e["l_offset"], e["r_offset"], e["n_offset_scores"], e["n_iters_final"]), c
        b = synth.generate(synth.CONFIGS[3], 40, seed=5).to_host()
        fast, cells_fast = oracle.count_loci(b.arena, b.seq_off, b.lens, b.est_cn, b.read_begin, b.motif_off, b.motif_len,
                                             n_threads=4)
        oracle.set_simd(False)
        slow, cells_slow = oracle.count_loci(b.arena, b.seq_off, b.lens, b.est_cn, b.read_begin, b.motif_off, b.motif_len,
                                             n_threads=4)
        assert np.array_equal(fast, slow) and cells_fast == cells_slow
    finally:
        oracle.set_simd(prev)
