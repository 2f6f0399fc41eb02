"""CPU-side tests (no GPU): the C-ABI library loads and exports what include/strkit_b200.h declares,
fails loudly without a device, and the host batcher / generators behave."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from strkit_b200 import _native

    header = open(os.path.join(ROOT, "include", "strkit_b200.h")).read()
    declared = set(re.findall(r"\b(strk_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.EXPORTED)
    assert _native.lib.strk_version().decode().startswith("strkit_b200")


def test_no_cpu_fallback_without_device():
    import strkit_b200
    from strkit_b200._native import StrkError

    if strkit_b200.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(StrkError, match="no CPU fallback"):
        strkit_b200.Engine()
    with pytest.raises(StrkError):
        strkit_b200.get_repeat_count(5, "CAGCAGCAGCAGCAG", "ACGT", "TTGA", "CAG",
                                     strkit_b200.RepeatCountParams("repalign", 50, 3, 1))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "strkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_matrix_matches_reference_golden(golden):
    from strkit_b200.align_matrix import dna_bases_str, dna_matrix

    assert dna_bases_str == golden["alphabet"]
    assert dna_matrix.tolist() == golden["matrix"]


def test_reference_rc_param_tiers():
    from strkit_b200 import get_reference_rc_params as g

    assert (g("repalign", 10, 250).max_iters, g("repalign", 10, 250).initial_step_size) == (250, 1)
    p = g("repalign", 200, 250)
    assert (p.max_iters, p.initial_step_size, p.initial_local_search_range) == (200, 3, 3)
    p = g("repalign", 1999, 250)
    assert (p.max_iters, p.initial_step_size, p.initial_local_search_range) == (150, 5, 3)
    p = g("repalign", 2000, 250)
    assert (p.max_iters, p.initial_step_size, p.initial_local_search_range) == (50, 15, 1)


def test_pack_loci_layout():
    from strkit_b200 import LocusReads, pack_loci

    b = pack_loci([LocusReads("CAG", [2, 3], ["CAGCAG", "CAGCAGCAG"], ["AA", "AC"], ["TT", "T"]),
                   LocusReads("AT", [1], ["AT"], [""], ["GG"])])
    b.validate()
    assert b.n_reads == 3 and b.n_loci == 2
    assert b.read_begin.tolist() == [0, 2, 3]
    assert b.lens.tolist() == [[2, 6, 2], [2, 9, 1], [0, 2, 2]]
    arena = bytes(b.arena)
    assert arena[int(b.seq_off[1]):int(b.seq_off[1]) + 12] == b"ACCAGCAGCAGT"
    assert arena[int(b.motif_off[1]):int(b.motif_off[1]) + 2] == b"AT"
    s = b.slice_loci(1, 2)
    assert s.n_reads == 1 and s.read_begin.tolist() == [0, 1] and s.lens.tolist() == [[0, 2, 2]]


def test_pack_loci_c_helper_equals_python_body():
    """csrc/fastpack.c (built by __graft_entry__.build()) and the pure-Python body of pack_loci produce the same
    arrays: ragged loci, empty strings, a locus without reads, tuples instead of lists, numpy estimates."""
    import importlib

    import __graft_entry__
    from strkit_b200 import LocusReads, batcher
    from tests.helpers import random_families

    if batcher._fastpack is None:  # first run in a fresh tree
        __graft_entry__.build()
        importlib.reload(batcher)
    if batcher._fastpack is None:
        pytest.skip("no Python.h / C compiler on this box: _fastpack.so cannot be built")
    pack_loci = batcher.pack_loci
    rng = np.random.default_rng(9)
    fams = random_families(rng, 300)
    loci, at = [], 0
    while at < len(fams):
        n = int(rng.integers(0, 6))
        grp = fams[at:at + n]
        at += n
        motif = grp[0][0] if grp else "ACG"
        loci.append(LocusReads(motif, np.array([len(t) // len(motif) for _, t, _, _ in grp], dtype=np.int64),
                               tuple(t for _, t, _, _ in grp), [fl for _, _, fl, _ in grp], [fr for _, _, _, fr in grp]))
    for subset in (loci, loci[:1], []):
        a, b = pack_loci(subset, use_helper=True), pack_loci(subset, use_helper=False)
        for f in ("arena", "seq_off", "lens", "est_cn", "read_begin", "motif_off", "motif_len"):
            x, y = getattr(a, f), getattr(b, f)
            assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), f
        a.validate()
    with pytest.raises(ValueError):
        pack_loci([LocusReads("AC", [1], ["AC\u00e9"], ["A"], ["C"])])
    with pytest.raises(ValueError):
        pack_loci([LocusReads("AC", [1, 2], ["AC"], ["A"], ["C"])])


def test_synth_generator_is_deterministic_and_sane(oracle):
    from strkit_b200 import synth

    a = synth.generate(synth.CONFIGS[1], 50, seed=1).to_host()
    b = synth.generate(synth.CONFIGS[1], 50, seed=1).to_host()
    assert np.array_equal(a.arena, b.arena) and np.array_equal(a.lens, b.lens)
    assert a.n_reads == 1500 and (a.lens[:, 0] <= 70).all() and (a.lens[:, 2] <= 70).all()
    assert set(np.unique(a.arena)) <= set(b"ACGTX")
    # the generator's copies are recovered by the reference search on almost every HiFi read
    sb = synth.generate(synth.CONFIGS[1], 50, seed=1)
    out, _ = oracle.count_loci(a.arena, a.seq_off, a.lens, a.est_cn, a.read_begin, a.motif_off, a.motif_len,
                               n_threads=4)
    assert (out[:, 0] == sb.true_cn.numpy()).mean() > 0.97


def test_install_rebinds_strkit_names(monkeypatch):
    """install() must patch both strkit.call.repeats and the from-imported names in strkit.call.call_locus
    (call_locus.py:32).  Uses stand-in modules: the real strkit is not installable in the test box."""
    import sys
    import types

    import strkit_b200
    from strkit_b200 import repeats as ours

    pkg, call = types.ModuleType("strkit"), types.ModuleType("strkit.call")
    rep, loc = types.ModuleType("strkit.call.repeats"), types.ModuleType("strkit.call.call_locus")
    rep.get_repeat_count = rep.get_ref_repeat_count = lambda *a, **k: "reference"
    loc.get_repeat_count, loc.get_ref_repeat_count = rep.get_repeat_count, rep.get_ref_repeat_count
    for name, mod in (("strkit", pkg), ("strkit.call", call), ("strkit.call.repeats", rep),
                      ("strkit.call.call_locus", loc)):
        monkeypatch.setitem(sys.modules, name, mod)
    patched = strkit_b200.install()
    assert len(patched) == 4
    assert loc.get_repeat_count is ours.get_repeat_count and rep.get_ref_repeat_count is ours.get_ref_repeat_count
    strkit_b200.uninstall()
    assert loc.get_repeat_count() == "reference" and rep.get_ref_repeat_count() == "reference"
    # opt-in: call_alleles (allele.py:176, from-imported at call_locus.py:28)
    al = types.ModuleType("strkit.call.allele")
    al.call_alleles = loc.call_alleles = lambda *a, **k: "reference"
    monkeypatch.setitem(sys.modules, "strkit.call.allele", al)
    assert len(strkit_b200.install()) == 4 and loc.call_alleles() == "reference"
    strkit_b200.uninstall()
    patched = strkit_b200.install(alleles=True)
    assert len(patched) == 6 and loc.call_alleles is strkit_b200.call_alleles and al.call_alleles is strkit_b200.call_alleles
    strkit_b200.uninstall()
    assert loc.call_alleles() == "reference" and al.call_alleles() == "reference"


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line on stdout with the
    contract's keys, exactly K timed steps; ranks other than 0 print nothing."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-sample-loci", "48"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    for key in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "vs_baseline", "dtype", "data"):
        assert key in d, key
    assert "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"),
                         timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""
