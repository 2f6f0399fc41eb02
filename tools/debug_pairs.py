import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import strkit_b200
from strkit_b200 import synth
from tests import oracle_lib
orc = oracle_lib.load()
eng = strkit_b200.Engine()
params = strkit_b200.RepeatCountParams("repalign", 50, 3, 1)
for n_loci, seed in ((1, 3), (2, 4), (40, 11)):
    b = synth.generate(synth.CONFIGS[1], n_loci, seed=seed).to_host()
    got = eng.count_reads(b, params)
    want, _ = orc.count_loci(b.arena, b.seq_off, b.lens, b.est_cn, b.read_begin, b.motif_off, b.motif_len, n_threads=4)
    bad = np.flatnonzero((got != want).any(axis=1))
    n1 = b.lens.sum(axis=1)
    print(n_loci, "reads", b.n_reads, "bad", len(bad), eng.stats()["reads_packed_kernel"], flush=True)
    for r in bad[:12]:
        l = np.searchsorted(b.read_begin, r, side="right") - 1
        print("  read", r, "pos", r, "n1", n1[r], "cls", (n1[r] + 1 + 31) // 32, "m", b.motif_len[l], "lens", b.lens[r], "got", got[r], "want", want[r], flush=True)
