#!/usr/bin/env python
"""How many results change between the hypotheses about strkit_rust_ext.get_repeat_count that the tree cannot pin
(DESIGN 2): search policy (range narrowing, repeat_count_params.py:13), tie-breaks, free-end mode.
Runs BASELINE configs 2 and 3 on the GPU under every switch and counts the reads whose n / score / n_explored /
carried start differ from the default (in-tree) semantics.  Output: one JSON object (committed under profiles/)."""
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strkit_b200 as sb  # noqa: E402
from strkit_b200 import synth  # noqa: E402


def main():
    n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    params = sb.RepeatCountParams("repalign", 50, 3, 1)
    report = {"loci_per_config": n_loci, "params": "max_iters 50, range 3, step 1", "configs": {}}
    for cfg in (2, 3):
        batch = synth.generate(synth.CONFIGS[cfg], n_loci, seed=5150 + cfg, device="cuda").to_host()
        base_eng = sb.Engine()
        base = base_eng.count_reads(batch, params)
        base_eng.close()
        rows = {}
        variants = [("search_narrow_first (flag 4)", dict(tie_flags=4)), ("search_narrow_halve (flag 8)", dict(tie_flags=8)),
                    ("tie_window_last (flag 1)", dict(tie_flags=1)), ("tie_final_last (flag 2)", dict(tie_flags=2))]
        variants += [(f"end_flags {f} ({name})", dict(end_flags=f)) for f, name in
                     ((0, "global / nw"), (2, "sg_qe"), (10, "sg_qx: db ends free"), (5, "sg_dx-like: begins free"),
                      (12, "candidate ends free"), (3, "db begin+end free"))]
        for name, kw in variants:
            eng = sb.Engine(**kw)
            got = eng.count_reads(batch, params)
            eng.close()
            d = got != base
            rows[name] = {"reads": int(batch.n_reads), "n_differs": int(d[:, 0].sum()), "score_differs": int(d[:, 1].sum()),
                          "n_explored_differs": int(d[:, 2].sum()), "start_differs": int(d[:, 3].sum()),
                          "n_differs_frac": float(d[:, 0].mean()), "mean_n_explored": float(got[:, 2].mean()),
                          "mean_n_explored_default": float(base[:, 2].mean())}
        report["configs"][f"config {cfg}"] = rows
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
