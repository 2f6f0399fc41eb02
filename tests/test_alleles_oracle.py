"""CPU tests of the allele-calling checker: the restatement (oracle/alleles_oracle.py) reproduces, bit for bit,
the vectors produced by executing the reference's own allele.py / gmm.py (tests/golden/gen_alleles_golden.py)."""
import numpy as np

from oracle import alleles_oracle as ao
from tests.alleles_helpers import load_golden


def _hex(v):
    return [float.hex(float(x)) for x in np.asarray(v).ravel()]


def test_oracle_reproduces_reference_vectors_exactly():
    doc = load_golden()
    import sklearn

    assert sklearn.__version__ == doc["sklearn"], "golden vectors were generated with another scikit-learn"
    assert len(doc["cases"]) >= 40
    for c in doc["cases"]:
        p = ao.OracleParams(num_bootstrap=c["num_bootstrap"], min_allele_reads=c["min_allele_reads"],
                            force_gm_filter=c["force_gm_filter"])
        got = ao.call_alleles(c["cn"], c["w"], c["n_alleles"], c["min_reads"], c["seed"], p)
        exp = c["expect"]
        if exp is None:
            assert got is None, c["tag"]
            continue
        assert got["call"].tolist() == exp["call"], c["tag"]
        assert got["call_95_cis"].tolist() == exp["call_95_cis"], c["tag"]
        assert got["call_99_cis"].tolist() == exp["call_99_cis"], c["tag"]
        assert _hex(got["means"]) == exp["means"] and _hex(got["weights"]) == exp["weights"], c["tag"]
        assert _hex(got["stdevs"]) == exp["stdevs"] and got["modal_n"] == exp["modal_n"], c["tag"]


def test_oracle_known_answers():
    p = ao.OracleParams()
    assert ao.call_alleles([12, 13, 12], [1 / 3] * 3, 2, 4, 1, p) is None            # fewer than min_reads
    r = ao.call_alleles([12] * 9, [1 / 9] * 9, 2, 4, 1, p)                           # one value: no bootstrap
    assert r["call"].tolist() == [12, 12] and r["modal_n"] == 1 and r["stdevs"].tolist() == [0.0, 0.0]
    # well separated, balanced alleles: the call is the pair, whatever the seed
    cn = [20] * 15 + [30] * 15
    for seed in (1, 2, 3):
        r = ao.call_alleles(cn, [1 / 30] * 30, 2, 4, seed, p)
        assert r["call"].tolist() == [20, 30] and r["modal_n"] == 2
        assert r["call_95_cis"].tolist() == [[20, 20], [30, 30]]
