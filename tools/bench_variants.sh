#!/bin/bash
# Kernel-variant experiments on the GPU box: tools/bench_variants.sh "NAME:ENV=VAL ENV2=VAL2" ...
# Each entry runs a short resident-batch bench (no CPU baseline, no e2e) and prints one summary line.
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/var_{n}.json"))
    print(f"{n:24s} value {d['value']/1e6:7.2f} M/s  dp_ms/step {d['roofline']['kernel_ms_per_step']:7.3f}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/var_{n}.err").read()[-400:])
PY
done
