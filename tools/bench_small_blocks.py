#!/usr/bin/env python
"""Blocking-call latency of small locus blocks (the reference hands <= 200 loci to a worker at a time,
call_sample.py:103-157; config 1 is 1000 loci): mean over repeated Engine.count_reads calls, host arrays pinned.

    python tools/bench_small_blocks.py            # config 1 loci (short reads)
    python tools/bench_small_blocks.py --config 4 # expansion loci (50 reads, half of them 1-6 kb): 1, 2, 5, 20 loci per call
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from strkit_b200 import synth, Engine, RepeatCountParams

cfg4 = "--config" in sys.argv and sys.argv[sys.argv.index("--config") + 1] == "4"
eng = Engine(0)
params = RepeatCountParams("repalign", 50, 3, 1)
res = []
for n_loci in ((1, 2, 5, 20) if cfg4 else (1, 20, 200, 1000, 4000)):
    if cfg4:
        b = synth.generate_expansions(n_loci, 50, seed=20261018 + 4000 + n_loci)[0]
    else:
        b = synth.generate(synth.CONFIGS[1], n_loci, seed=20261018 + 1000, device="cuda").to_host(pin=True)
    for _ in range(5):
        eng.count_reads(b, params)
    reps = 40
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.count_reads(b, params)
    dt = (time.perf_counter() - t0) / reps
    st = eng.stats()
    res.append({"loci": n_loci, "reads": int(b.n_reads), "ms_per_call": dt * 1e3, "reads_per_s": b.n_reads / dt,
                "dp_ms": st["dp_ms"], "replay_ms": st["replay_ms"]})
    print(json.dumps(res[-1]), flush=True)
