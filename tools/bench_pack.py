#!/usr/bin/env python
"""Host-side packing rate of a block of loci (strkit_b200.batcher.pack_loci -> csrc/fastpack.c): per-read Python str
objects, as call_locus holds them (call_locus.py:1144-1155), into the flat arrays of a ReadBatch.  CPU only.

    python tools/bench_pack.py [n_loci]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from strkit_b200.batcher import LocusReads, pack_loci  # noqa: E402


def main():
    n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    rng = np.random.default_rng(1)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)

    def seq(n):
        return letters[rng.integers(0, 4, n)].tobytes().decode()

    loci = []
    for _ in range(n_loci):  # config 2 shapes: motif 2-6 bp x 10-60 copies, 30 reads, 70-base flanks
        m = seq(int(rng.integers(2, 7)))
        trs = [m * int(rng.integers(10, 61)) for _ in range(30)]
        loci.append(LocusReads(m, [len(t) // len(m) for t in trs], trs, [seq(70) for _ in range(30)], [seq(70) for _ in range(30)]))
    n_reads = 30 * n_loci
    out = {"n_loci": n_loci, "n_reads": n_reads, "host_cores": os.cpu_count(), "rates_M_reads_per_s": {}}
    for nibble in (False, True):
        for threads in (1, 2, 4, 8, 16, 0):
            best = 1e9
            for _ in range(4):
                t = time.perf_counter()
                b = pack_loci(loci, nibble=nibble, threads=threads)
                best = min(best, time.perf_counter() - t)
            out["rates_M_reads_per_s"][f"{'nibble' if nibble else 'ascii'}_threads_{threads or 'auto'}"] = round(n_reads / best / 1e6, 2)
            out[f"arena_bytes_{'nibble' if nibble else 'ascii'}"] = int(b.arena.nbytes)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
