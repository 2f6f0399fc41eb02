#!/usr/bin/env python
"""Throughput + parity of the BASELINE.json configs other than the headline one (bench.py covers config 2).

    python tools/bench_configs.py --configs 1,3,4 > gpurun_out/configs.json

Every config goes through the host-buffer call (Engine.count_reads: H2D + kernels + D2H timed) and is
checked bit-exact against the CPU oracle on a bounded sample of loci.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def run_config(eng, params, name, batches, oracle, sample_loci, threads):
    import strkit_b200  # noqa: F401

    reads = loci = 0
    t_total = 0.0
    agg = dict(executed_cells=0.0, reference_cells=0.0, dp_ms=0.0, replay_ms=0.0, widening_passes=0.0,
               reads_packed_kernel=0.0, reads_general_kernel=0.0)
    parity = None
    cpu = None
    for bi, batch in enumerate(batches):
        eng.count_reads(batch.slice_loci(0, min(8, batch.n_loci)), params)  # warm buffers
        t0 = time.perf_counter()
        out = eng.count_reads(batch, params)
        t_total += time.perf_counter() - t0
        st = eng.stats()
        for k in agg:
            agg[k] += st[k]
        reads += batch.n_reads
        loci += batch.n_loci
        if bi == 0:
            sub = batch.slice_loci(0, min(sample_loci, batch.n_loci))
            t0 = time.perf_counter()
            want, cells = oracle.count_loci(sub.arena, sub.seq_off, sub.lens, sub.est_cn, sub.read_begin,
                                            sub.motif_off, sub.motif_len, max_iters=params.max_iters,
                                            n_threads=threads)
            dt = time.perf_counter() - t0
            parity = bool(np.array_equal(out[:sub.n_reads], want))
            cpu = {"value": sub.n_reads / dt, "unit": "reads*loci/s", "cores": threads, "kind": "port",
                   "gcups": cells / dt / 1e9, "sample": f"{sub.n_loci} loci, {sub.n_reads} reads, {dt:.1f} s"}
    return {"config": name, "loci": loci, "reads": reads, "e2e_reads_per_s": reads / t_total, "e2e_s": t_total,
            "gcups_executed_e2e": agg["executed_cells"] / t_total / 1e9,
            "gcups_reference_equivalent_e2e": agg["reference_cells"] / t_total / 1e9,
            "dp_kernel_ms": agg["dp_ms"], "gcups_executed_kernel": agg["executed_cells"] / max(agg["dp_ms"], 1e-9) / 1e6,
            "widening_passes": int(agg["widening_passes"]), "reads_packed_kernel": int(agg["reads_packed_kernel"]),
            "reads_general_kernel": int(agg["reads_general_kernel"]), "parity_sample_bit_exact": parity,
            "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4")
    ap.add_argument("--loci3", type=int, default=100_000)
    ap.add_argument("--sample-loci", type=int, default=256)
    args = ap.parse_args()
    import torch

    import strkit_b200
    from strkit_b200 import synth
    from tests import oracle_lib

    oracle = oracle_lib.load()
    threads = os.cpu_count() or 1
    eng = strkit_b200.Engine()
    params = strkit_b200.RepeatCountParams("repalign", 50, 3, 1)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    results = []
    for c in [int(x) for x in args.configs.split(",")]:
        if c == 1:
            batches = [synth.generate(synth.CONFIGS[1], 1000, device=dev).to_host()]
            results.append(run_config(eng, params, "cfg1: 1k loci x 30 HiFi reads", batches, oracle, 1000, threads))
        elif c == 3:
            batches = []
            left = args.loci3
            i = 0
            while left > 0:
                n = min(16384, left)
                batches.append(synth.generate(synth.CONFIGS[3], n, seed=20261018 + 3000 + i, device=dev).to_host())
                left -= n
                i += 1
            results.append(run_config(eng, params, f"cfg3: {args.loci3} loci x 40 ONT-like reads (~5% errors)", batches,
                                      oracle, args.sample_loci, threads))
        elif c == 4:
            batch, _ = synth.generate_expansions(60, 50)
            results.append(run_config(eng, params, "cfg4: 60 pathogenic-style loci x 50 reads, expansions to 6 kb",
                                      [batch], oracle, 2, threads))
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
