#!/usr/bin/env python
"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Packed kernel (both lane widths, reference mode), general kernel (multi-strip pipeline, steady-state loop, paired
sweeps), dedupe / planning / nibble expansion, replay and widening passes, realignment fill + traceback, allele calls."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strkit_b200 as sb  # noqa: E402
from strkit_b200 import realign, synth  # noqa: E402
from strkit_b200.batcher import LocusReads, pack_loci  # noqa: E402


def main():
    p = sb.RepeatCountParams("repalign", 50, 3, 1)
    eng = sb.Engine()
    b = synth.generate(synth.CONFIGS[2], 600, seed=3).to_host()          # > 512 reads: device-side planning + dedupe
    out = eng.count_reads(b.to_nibble(), p)
    assert np.array_equal(out, eng.count_reads(b, p))
    b3 = synth.generate(synth.CONFIGS[3], 40, seed=4).to_host()          # noisy: widening passes
    b3.est_cn[::7] += 9
    eng.count_reads(b3, p)
    eng.count_reads(b3, p, kernel=sb.KERNEL_GENERAL)
    bx, _ = synth.generate_expansions(n_loci=3, reads_per_locus=3, max_tract=2600, big_lo=300, big_hi=800,
                                      motifs=["CAG", "RAAAT", "GCN"])     # multi-strip general kernel, IUPAC motifs
    eng.count_reads(bx, p)
    first = b.read_begin[:-1][:200]
    loci = []
    arena = b.arena.tobytes().decode()
    for l, r in enumerate(first):
        o, (fl, tr, fr) = int(b.seq_off[r]), (int(v) for v in b.lens[r])
        motif = arena[int(b.motif_off[l]):int(b.motif_off[l]) + int(b.motif_len[l])]
        loci.append(LocusReads(motif, [int(b.est_cn[r])], [arena[o + fl:o + fl + tr]], [arena[o:o + fl]],
                               [arena[o + fl + tr:o + fl + tr + fr]]))
    loci.append(LocusReads("CAG", [220], ["CAG" * 220], ["ACGTTGCATGCATTGACCATGACTGAATCG"], ["TTGACGATCGGATCGATTAGCTAGCTAAGC"]))
    rb = pack_loci(loci)
    rc = np.tile(np.array([250, 3, 1], dtype=np.int32), (rb.n_loci, 1))
    rc[-1] = [200, 3, 3]
    eng.ref_counts(rb, rb.est_cn.copy(), rb.lens[:, 1].copy(), rc, 5)
    for _ in eng.count_reads_stream([b.slice_loci(0, 300), b.slice_loci(300, 600)], p):
        pass
    rng = np.random.default_rng(1)
    ref = "".join(rng.choice(list("ACGT"), size=600))
    read = "".join(rng.choice(list("ACGT"), size=900)) + ref[:300] + "CAG" * 40 + ref[300:] + "".join(rng.choice(list("ACGT"), size=500))
    realign.realign_batch([(ref, read), (ref[:90], read[800:1400])], engine=eng)
    cn = np.array([10, 10, 11, 14, 14, 14, 15, 10, 11, 14] * 4, dtype=np.int32)
    sb.call_alleles_batch(cn, np.full(cn.shape[0], 0.1), np.arange(0, 41, 10, dtype=np.int64), 2, num_bootstrap=20, seed=1,
                          engine=eng)
    eng.close()
    print("sanitize smoke: done")


if __name__ == "__main__":
    main()
