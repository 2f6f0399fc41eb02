"""Soft-clip realignment (strkit/call/realign.py:34-72).  CPU: the checker against an independent
pure-Python three-state DP and against properties every CIGAR must have.  GPU (-m gpu): strk_realign against the
checker, bit for bit (score, end_ref, CIGAR) under every traceback switch."""
import numpy as np
import pytest

from tests.helpers import mutate, rand_seq


def _py_affine_sg_dx(s1, s2, M, sym, o, e):
    """Gotoh, s1 global / s2 ends free: best score only (independent statement: full matrices, no rolling rows)."""
    NEG = -10 ** 9
    n1, n2 = len(s1), len(s2)
    H = [[0] * (n2 + 1) for _ in range(n1 + 1)]
    E = [[NEG] * (n2 + 1) for _ in range(n1 + 1)]
    F = [[NEG] * (n2 + 1) for _ in range(n1 + 1)]
    for i in range(1, n1 + 1):
        H[i][0] = -o - (i - 1) * e
        for j in range(1, n2 + 1):
            E[i][j] = max(H[i][j - 1] - o, E[i][j - 1] - e)
            F[i][j] = max(H[i - 1][j] - o, F[i - 1][j] - e)
            H[i][j] = max(H[i - 1][j - 1] + M[sym(s1[i - 1])][sym(s2[j - 1])], E[i][j], F[i][j])
    return max(H[n1])


def _score_of_cigar(cigar, s1, s2, M, sym, o, e):
    """Re-score a CIGAR that starts at (0, 0): leading read bases are free, every other gap run costs o + (k-1)e."""
    i = j = 0
    total = 0
    for k, c in enumerate(cigar):
        n, op = int(c) >> 4, int(c) & 15
        if op in (7, 8):
            for _ in range(n):
                sc = M[sym(s1[i])][sym(s2[j])]
                assert (sc > 0) == (op == 7)
                total += sc
                i += 1
                j += 1
        elif op == 1:
            total -= o + (n - 1) * e
            i += n
        else:
            assert op == 2
            if k > 0 or i > 0:
                total -= o + (n - 1) * e
            j += n
    return total, i, j


@pytest.mark.parametrize("flags", [0, 1, 2, 4, 7])
def test_realign_checker_properties(oracle, flags):
    rng = np.random.default_rng(50 + flags)
    M = oracle.matrix.reshape(17, 17).tolist()
    for it in range(60):
        n1 = int(rng.integers(1, 90))
        ref = rand_seq(rng, n1, "ACGT" if it % 4 else "ACGTNXR")
        inner = mutate(rng, ref, 0.04, 0.04, 0.04) or "A"
        if it % 5 == 0:   # an expansion inside the window: one long insertion in the read
            k = int(rng.integers(0, len(inner)))
            inner = inner[:k] + "CAG" * int(rng.integers(3, 30)) + inner[k:]
        read = rand_seq(rng, int(rng.integers(0, 60))) + inner + rand_seq(rng, int(rng.integers(0, 60)))
        o = 7 if it % 3 else int(rng.integers(1, 9))
        e = 0 if it % 3 else int(rng.integers(0, min(o, 3) + 1))   # extend <= open (a dearer extension would be re-opened)
        score, end_ref, cigar = oracle.realign(ref, read, o, e, flags)
        assert score == _py_affine_sg_dx(ref, read, M, oracle.symbol, o, e), (ref, read, o, e)
        total, used1, used2 = _score_of_cigar(cigar, ref, read, M, oracle.symbol, o, e)
        assert (total, used1, used2) == (score, n1, end_ref + 1), (ref, read, cigar)
        ops = [int(c) & 15 for c in cigar]
        assert all(a != b for a, b in zip(ops, ops[1:]))          # runs are merged
    # known answer: the window sits verbatim inside the read
    ref = "ACGTTGCATGCAGCAGCAGCAGTTGACCATGA"
    read = "GGGTTTAAACCC" + ref + "TTTGGGAAAC"
    score, end_ref, cigar = oracle.realign(ref, read)
    assert (score, end_ref) == (2 * len(ref), 12 + len(ref) - 1) and [int(c) for c in cigar] == [(12 << 4) | 2, (len(ref) << 4) | 7]
    # a 30-base expansion in the read costs the open penalty once (extend = 0)
    read2 = "GGGTTTAAACCC" + ref[:16] + "CAG" * 10 + ref[16:] + "TTTGGGAAAC"
    assert oracle.realign(ref, read2)[0] == 2 * len(ref) - 7


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 1, 2, 4, 7])
def test_realign_gpu_equals_checker(oracle, flags):
    import strkit_b200 as sb
    from strkit_b200 import realign

    rng = np.random.default_rng(70 + flags)
    pairs = []
    for it in range(48):
        n1 = int(rng.choice([1, 5, 33, 100, 128, 129, 257, 300, 384, 400, 513, 700, 1100]))
        ref = rand_seq(rng, n1, "ACGT" if it % 4 else "ACGTNXRacgt")
        inner = mutate(rng, ref, 0.03, 0.02, 0.02) or "A"
        if it % 3 == 0:
            k = int(rng.integers(0, len(inner)))
            inner = inner[:k] + "CAG" * int(rng.integers(3, 200)) + inner[k:]
        pairs.append((ref, rand_seq(rng, int(rng.integers(0, 2500))) + inner + rand_seq(rng, int(rng.integers(0, 2500)))))
    pairs.append(("ACGT", "A"))                      # read shorter than the window
    pairs.append(("A" * 40, "C" * 70))               # nothing matches
    got = realign.realign_batch(pairs, trace_flags=flags)
    for (ref, read), (score, end_ref, cigar) in zip(pairs, got):
        w_score, w_end, w_cigar = oracle.realign(ref, read, 7, 0, flags)
        assert (score, end_ref) == (w_score, w_end), (len(ref), len(read))
        assert np.array_equal(cigar, w_cigar), (len(ref), len(read), realign.cigar_to_string(cigar)[:80],
                                                realign.cigar_to_string(w_cigar)[:80])
    # other gap models, several groups (tiny trace budget), the drop-in call with the reference's threshold
    import os

    os.environ["STRK_REALIGN_TRACE_MB"] = "1"
    try:
        got2 = realign.realign_batch(pairs[:12], gap_open=5, gap_extend=2, trace_flags=flags)
    finally:
        del os.environ["STRK_REALIGN_TRACE_MB"]
    for (ref, read), (score, end_ref, cigar) in zip(pairs[:12], got2):
        w = oracle.realign(ref, read, 5, 2, flags)
        assert (score, end_ref) == w[:2] and np.array_equal(cigar, w[2])
    fl = rand_seq(rng, 70)
    fr = rand_seq(rng, 70)
    ref = fl + "CAG" * 20 + fr
    read = rand_seq(rng, 900) + fl + "CAG" * 55 + fr + rand_seq(rng, 1200)
    res = realign.realign_read(ref, read, 1_000_000, 70)
    assert res is not None
    read_coords, ref_coords = res
    assert read_coords.shape == ref_coords.shape and ref_coords[0] == 1_000_000 and read_coords[0] == 900
    assert ref_coords[-1] == 1_000_000 + len(ref) - 1 and (np.diff(read_coords) > 0).all()
    assert realign.realign_read(ref, rand_seq(rng, 3000), 1_000_000, 70) is None      # unrelated read: below the threshold
    assert sb.default_engine().stats()["executed_cells"] > 0
