#!/usr/bin/env python
"""Throughput + parity of the BASELINE.json configs other than the headline one (bench.py covers config 2).

    python tools/bench_configs.py --configs 1,3,4 > gpurun_out/configs.json

Every config goes through the host-buffer calls (Engine.count_reads_stream over its locus blocks, pinned host
arrays: H2D + kernels + D2H timed) and is checked bit-exact against the CPU oracle on a bounded sample of loci.
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

    reads = sum(b.n_reads for b in batches)
    loci = sum(b.n_loci for b in batches)
    agg = dict(executed_cells=0.0, reference_cells=0.0, dp_ms=0.0, replay_ms=0.0, widening_passes=0.0,
               reads_packed_kernel=0.0, reads_general_kernel=0.0)
    parity = None
    cpu = None
    # untimed pass: warms the recycled device buffers, collects the per-batch counters
    outs = []
    for batch in batches:
        outs.append(eng.count_reads(batch, params))
        st = eng.stats()
        for k in agg:
            agg[k] += st[k]
    for _ in eng.count_reads_stream(batches[:2], params):  # second context
        pass
    # timed pass: the blocks streamed through the host-buffer calls (H2D + kernels + D2H)
    t0 = time.perf_counter()
    streamed = list(eng.count_reads_stream(batches, params))
    t_total = time.perf_counter() - t0
    assert all(np.array_equal(a, b) for a, b in zip(outs, streamed))
    out = outs[0]
    batch = batches[0]
    sub = batch.slice_loci(0, min(sample_loci, batch.n_loci))
    t0 = time.perf_counter()
    want, cells = oracle.count_loci(sub.arena, sub.seq_off, sub.lens, sub.est_cn, sub.read_begin, sub.motif_off,
                                    sub.motif_len, max_iters=params.max_iters, n_threads=threads)
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


def run_full_pipeline(eng, params, n_loci, num_bootstrap, dev, sample_loci, threads, oracle):
    """Config 5: repeat counting (streamed locus blocks) + bootstrap / GMM allele calls with `num_bootstrap`
    replicates, end to end from host arrays to per-locus calls and confidence intervals on the host.  Read weights:
    uniform per locus (get_read_weight belongs to the read-extraction step, SURVEY 8f N1, not built).
    CPU side: the oracle port for the counts + the reference algorithm for the allele calls (numpy + scikit-learn,
    oracle/alleles_oracle.py) on a bounded sample of loci, one process per core like the reference's worker pool."""
    import multiprocessing as mp

    from strkit_b200 import alleles, synth

    blocks, left, i = [], n_loci, 0
    while left > 0:
        n = min(16384, left)
        blocks.append(synth.generate(synth.CONFIGS[5], n, seed=20261018 + 5000 + i, device=dev).to_host(pin=True))
        left -= n
        i += 1
    for _ in eng.count_reads_stream(blocks[:2], params):  # warm both contexts
        pass
    def call_block(block, counts_block, seed):
        rb_b = block.read_begin
        w_b = (1.0 / np.repeat(np.diff(rb_b), np.diff(rb_b))).astype(np.float64)
        return alleles.call_alleles_batch(counts_block[:, 0], w_b, rb_b, 2, num_bootstrap=num_bootstrap, seed=seed, engine=eng)

    call_block(blocks[0], eng.count_reads(blocks[0], params), 1)  # grows the recycled device buffers to block size
    t0 = time.perf_counter()
    counts, parts, t_call = [], [], 0.0
    for bi, c in enumerate(eng.count_reads_stream(blocks, params)):  # counts of block i+1 overlap the calls of block i
        t1 = time.perf_counter()
        parts.append(call_block(blocks[bi], c, 1234 + bi))
        t_call += time.perf_counter() - t1
        counts.append(c)
    t_total = time.perf_counter() - t0
    t_count = t_total - t_call
    rb = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(b.read_begin) for b in blocks]))]).astype(np.int64)
    calls = alleles.AlleleCalls(*(np.concatenate([getattr(p, f) for p in parts]) for f in
                                  ("call", "call_95_cis", "call_99_cis", "means", "weights", "stdevs", "modal_n", "status")),
                                kernel_ms=sum(p.kernel_ms for p in parts))
    reads = int(rb[-1])
    # ---- CPU side on a bounded sample
    sub = blocks[0].slice_loci(0, min(sample_loci, blocks[0].n_loci))
    tc0 = time.perf_counter()
    want, _ = oracle.count_loci(sub.arena, sub.seq_off, sub.lens, sub.est_cn, sub.read_begin, sub.motif_off,
                                sub.motif_len, max_iters=params.max_iters, n_threads=threads)
    t_cpu_count = time.perf_counter() - tc0
    parity_counts = bool(np.array_equal(counts[0][:sub.n_reads], want))
    jobs = [(want[sub.read_begin[l]:sub.read_begin[l + 1], 0].copy(), num_bootstrap, 1234 + l) for l in range(sub.n_loci)]
    tc1 = time.perf_counter()
    with mp.get_context("fork").Pool(threads) as pool:
        cpu_calls = pool.map(_cpu_call_one, jobs)
    t_cpu_call = time.perf_counter() - tc1
    same = sum(cpu_calls[l] is not None and calls.call[l].tolist() == cpu_calls[l][0] for l in range(sub.n_loci))
    ci1 = sum(cpu_calls[l] is not None and int(np.abs(calls.call_95_cis[l] - np.array(cpu_calls[l][1])).max()) <= 1
              for l in range(sub.n_loci))
    return {"config": f"cfg5: full pipeline, {n_loci} loci x 30 HiFi reads, counts + {num_bootstrap}-replicate bootstrap GMM",
            "loci": n_loci, "reads": reads, "e2e_reads_per_s": reads / t_total, "e2e_s": t_total,
            "count_s": t_count, "allele_call_s": t_call, "allele_kernel_ms": calls.kernel_ms,
            "loci_per_s_allele_calls": n_loci / t_call,
            "status_counts": {str(k): int(v) for k, v in zip(*np.unique(calls.status, return_counts=True))},
            "parity": {"counts_bit_exact_sample": parity_counts, "sample_loci": sub.n_loci,
                       "calls_identical_to_cpu": same / sub.n_loci, "ci95_within_1_of_cpu": ci1 / sub.n_loci,
                       "note": "different random streams: statistical agreement (tests/test_gpu_alleles.py states the tolerance)"},
            "cpu_baseline": {"kind": "port (counts) + reference algorithm on numpy / scikit-learn (allele calls)",
                             "cores": threads, "sample": f"{sub.n_loci} loci, {sub.n_reads} reads",
                             "count_s": t_cpu_count, "allele_call_s": t_cpu_call,
                             "value": sub.n_reads / (t_cpu_count + t_cpu_call), "unit": "reads*loci/s",
                             "loci_per_s_allele_calls": sub.n_loci / t_cpu_call}}


def _cpu_call_one(job):
    from oracle import alleles_oracle as ao

    cn, num_bootstrap, seed = job
    r = ao.call_alleles(cn, np.full(len(cn), 1.0 / len(cn)), 2, 4, seed, ao.OracleParams(num_bootstrap=num_bootstrap))
    return None if r is None else (r["call"].tolist(), r["call_95_cis"].tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4")
    ap.add_argument("--loci5", type=int, default=100_000)
    ap.add_argument("--bootstrap5", type=int, default=1000)
    ap.add_argument("--cpu-sample-loci5", type=int, default=64)
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
            batches = [synth.generate(synth.CONFIGS[1], 1000, device=dev).to_host(pin=True)]
            results.append(run_config(eng, params, "cfg1: 1k loci x 30 HiFi reads", batches, oracle, 1000, threads))
        elif c == 3:
            batches = []
            left = args.loci3
            i = 0
            while left > 0:
                n = min(16384, left)
                batches.append(synth.generate(synth.CONFIGS[3], n, seed=20261018 + 3000 + i, device=dev).to_host(pin=True))
                left -= n
                i += 1
            results.append(run_config(eng, params, f"cfg3: {args.loci3} loci x 40 ONT-like reads (~5% errors)", batches,
                                      oracle, args.sample_loci, threads))
        elif c == 5:
            results.append(run_full_pipeline(eng, params, args.loci5, args.bootstrap5, dev, args.cpu_sample_loci5, threads,
                                             oracle))
        elif c == 4:
            batch, _ = synth.generate_expansions(60, 50)
            results.append(run_config(eng, params, "cfg4: 60 pathogenic-style loci x 50 reads, expansions to 6 kb",
                                      [batch], oracle, 2, threads))
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
