#!/usr/bin/env python
"""Per-call timings of config-3 blocks (ONT-like reads): first window policy and widening passes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import strkit_b200
from strkit_b200 import synth, Engine, RepeatCountParams

eng = Engine(0)
params = RepeatCountParams("repalign", 50, 3, 1)
blocks = [synth.generate(synth.CONFIGS[3], 16384, seed=20261018 + 3000 + i, device="cuda").to_host(pin=True) for i in range(3)]
for rep in range(3):
    for i, b in enumerate(blocks):
        t0 = time.perf_counter()
        out = eng.count_reads(b, params)
        dt = time.perf_counter() - t0
        st = eng.stats()
        print(f"rep {rep} block {i}: {dt*1e3:.2f} ms, dp {st['dp_ms']:.2f} ms, replay {st['replay_ms']:.2f}, widening {st['widening_passes']}, "
              f"packed {st['reads_packed_kernel']:.0f} general {st['reads_general_kernel']:.0f} launches {st.get('launches')}", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    n = sum(len(o) for o in eng.count_reads_stream(blocks, params))
    dt = time.perf_counter() - t0
    print(f"stream rep {rep}: {dt*1e3:.2f} ms, {n/dt/1e6:.2f} M reads/s", flush=True)
