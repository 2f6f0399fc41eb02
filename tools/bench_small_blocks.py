#!/usr/bin/env python
"""Blocking-call latency of small locus blocks (the reference hands <= 200 loci to a worker at a time,
call_sample.py:103-157; config 1 is 1000 loci): mean over repeated Engine.count_reads calls, host arrays pinned."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from strkit_b200 import synth, Engine, RepeatCountParams

eng = Engine(0)
params = RepeatCountParams("repalign", 50, 3, 1)
res = []
for n_loci in (1, 20, 200, 1000, 4000):
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
