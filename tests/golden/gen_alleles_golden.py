#!/usr/bin/env python
"""Generate golden vectors for the allele-calling step by EXECUTING the reference's own
strkit/call/allele.py and strkit/call/gmm.py (unmodified, imported from /root/reference).

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/gen_alleles_golden.py

Real: call_alleles, fit_gmm, get_resampled_bootstrapped_reads, GMMParams, make_single_gaussian, get_new_seed,
numpy's Generator, scikit-learn's GaussianMixture (1.9.0, the locked version).
Stubbed: strkit_rust_ext.CallData (a Rust class that only stores its constructor arguments) and
importlib.metadata.version("strkit") (strkit/__init__.py:7).
"""
from __future__ import annotations

import importlib.metadata
import importlib.util
import json
import logging
import sys
import types
from pathlib import Path

import numpy as np

REF = "/root/reference"
OUT = Path(__file__).parent


class CallData:  # stand-in for the PyO3 class: keeps the keyword arguments
    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_reference():
    _orig = importlib.metadata.version
    importlib.metadata.version = lambda name: "0.0.0-golden" if name == "strkit" else _orig(name)
    ext = types.ModuleType("strkit_rust_ext")
    ext.CallData = CallData
    sys.modules["strkit_rust_ext"] = ext
    sys.path.insert(0, REF)
    # strkit/call/__init__.py pulls in the whole caller (pysam, ...): load the two modules of this path directly
    pkg = types.ModuleType("strkit")
    pkg.__path__ = [f"{REF}/strkit"]
    sys.modules["strkit"] = pkg
    call_pkg = types.ModuleType("strkit.call")
    call_pkg.__path__ = [f"{REF}/strkit/call"]
    sys.modules["strkit.call"] = call_pkg

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("strkit.utils", f"{REF}/strkit/utils.py")
    load("strkit.call.constants", f"{REF}/strkit/call/constants.py")
    load("strkit.call.utils", f"{REF}/strkit/call/utils.py")
    gmm = load("strkit.call.gmm", f"{REF}/strkit/call/gmm.py")
    allele = load("strkit.call.allele", f"{REF}/strkit/call/allele.py")
    return allele, gmm


def make_cases(rng: np.random.Generator):
    cases = []

    def add(cn, n_alleles=2, num_bootstrap=100, min_reads=4, min_allele_reads=2, force=False, uniform_w=False, tag=""):
        cn = np.asarray(cn, dtype=np.int32)
        w = np.ones(len(cn)) if uniform_w else rng.uniform(0.5, 1.5, size=len(cn))
        w = w / w.sum()
        cases.append({"tag": tag, "cn": cn.tolist(), "w": w.tolist(), "n_alleles": n_alleles,
                      "num_bootstrap": num_bootstrap, "min_reads": min_reads, "min_allele_reads": min_allele_reads,
                      "force_gm_filter": force, "seed": int(rng.integers(0, 4096))})

    def reads(a1, a2, n, stutter=0.08):
        al = rng.choice([a1, a2], size=n)
        return al + rng.choice([-1, 0, 1], size=n, p=[stutter / 2, 1 - stutter, stutter / 2])

    for i in range(14):  # heterozygous, close alleles
        a1 = int(rng.integers(8, 50))
        add(reads(a1, a1 + int(rng.integers(1, 4)), int(rng.integers(12, 45))), tag="het_close")
    for i in range(8):  # homozygous with stutter
        a1 = int(rng.integers(8, 50))
        add(reads(a1, a1, int(rng.integers(10, 40)), stutter=0.15), tag="hom_stutter")
    for i in range(6):  # expansions: the large allele far above expansion_ratio * small
        a1 = int(rng.integers(10, 30))
        a2 = a1 * int(rng.integers(6, 30))
        n = int(rng.integers(20, 50))
        cn = np.where(rng.random(n) < 0.15, a2 + rng.integers(-20, 21, size=n), a1 + rng.choice([-1, 0, 1], size=n, p=[.05, .9, .05]))
        add(cn, tag="expansion")
    for i in range(4):
        add(reads(20, 23, 30), force=True, tag="force_gm_filter")
    for i in range(4):
        add(reads(15, 16, 25), n_alleles=1, min_reads=2, tag="haploid")
    add([12] * 20, tag="single_value")
    add([12, 13, 12], tag="too_few")
    add([12, 13, 12, 13], tag="min_reads_exact", uniform_w=True)
    add([30, 31] * 6 + [45], min_allele_reads=3, tag="min_allele_reads_3")
    add(reads(20, 26, 40), num_bootstrap=2, tag="two_replicates")  # num_bootstrap = 1 raises IndexError (allele.py:258)
    add(reads(20, 21, 30), num_bootstrap=37, tag="odd_bootstrap")
    add(rng.integers(10, 60, size=50), tag="noisy")
    add(reads(9, 40, 64, stutter=0.3), tag="far_apart")
    return cases


def main():
    allele, gmm = load_reference()
    rng = np.random.default_rng(20261018)
    cases = make_cases(rng)
    logger = logging.getLogger("golden")
    out = []
    for c in cases:
        params = types.SimpleNamespace(
            num_bootstrap=c["num_bootstrap"], min_allele_reads=c["min_allele_reads"], force_gm_filter=c["force_gm_filter"],
            gmm_params=gmm.GMMParams(init_params_method="k-means++", n_init=3, pre_filter_factor=5, expansion_ratio=5.0,
                                     filter_factor=3))
        cd = allele.call_alleles(np.array(c["cn"], dtype=np.int32), np.array([], dtype=np.int32),
                                 np.array(c["w"], dtype=np.float64), np.array([], dtype=np.float64), params=params,
                                 min_reads=c["min_reads"], n_alleles=c["n_alleles"], separate_strands=False,
                                 read_bias_corr_min=0, seed=c["seed"], logger_=logger, debug_str="golden")
        if cd is None:
            c["expect"] = None
        else:
            c["expect"] = {"call": np.asarray(cd.call).tolist(), "call_95_cis": np.asarray(cd.call_95_cis).tolist(),
                           "call_99_cis": np.asarray(cd.call_99_cis).tolist(),
                           "means": [float.hex(float(x)) for x in np.asarray(cd.means).ravel()],
                           "weights": [float.hex(float(x)) for x in np.asarray(cd.weights).ravel()],
                           "stdevs": [float.hex(float(x)) for x in np.asarray(cd.stdevs).ravel()],
                           "modal_n": int(cd.modal_n)}
        out.append(c)
    import sklearn

    doc = {"generator": "tests/golden/gen_alleles_golden.py", "numpy": np.__version__, "sklearn": sklearn.__version__,
           "cases": out}
    (OUT / "alleles_golden.json").write_text(json.dumps(doc))
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
