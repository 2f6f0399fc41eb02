"""CPU restatement of the reference's bootstrap / GMM allele caller -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
nothing under strkit_b200/ does (tests/test_host_cpu.py enforces it).

Restates strkit/call/allele.py with numpy and scikit-learn -- the third-party dependency the reference itself
calls (sklearn.mixture.GaussianMixture, pinned 1.9.0 = uv.lock:939-940, present in this image):

    call_alleles                      allele.py:176-336
    get_resampled_bootstrapped_reads  allele.py:126-173   (separate_strands = False branch: both call sites,
                                                            call_locus.py:201-214,255-268)
    fit_gmm                           allele.py:56-123
    make_single_gaussian              gmm.py:72-80
    GMMParams.make_fitted_gmm         gmm.py:59-69         get_new_seed: call/utils.py:29-30

PINNED: tests/golden/alleles_golden.json was produced by executing the reference's own allele.py / gmm.py
(tests/golden/gen_alleles_golden.py, run where /root/reference exists); with the same seeds this restatement
consumes numpy's Generator identically and reproduces those vectors exactly (tests/test_alleles_oracle.py).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np
from sklearn.exceptions import ConvergenceWarning
from sklearn.mixture import GaussianMixture

SMALL_ALLELE_MIN = 8                        # allele.py:47
F32_EPS = float(np.finfo(np.float32).eps)   # call/constants.py:11


@dataclass
class OracleParams:
    """The CallParams fields the path reads (params.py:39-56,166-172), reference defaults."""
    num_bootstrap: int = 100
    min_allele_reads: int = 2
    force_gm_filter: bool = False
    n_init: int = 3
    expansion_ratio: float = 5.0
    filter_factor: int = 3
    init_params_method: str = "k-means++"


class _Single:
    """make_single_gaussian (gmm.py:72-80)."""

    def __init__(self, sample_rs: np.ndarray):
        self.means_ = np.array([[np.mean(sample_rs)]])
        self.weights_ = np.array([[1.0]])
        self.covariances_ = np.array([[np.var(sample_rs)]])


def new_seed(rng: np.random.Generator) -> int:
    return rng.integers(0, 4096, dtype=int)  # call/utils.py:29-30


def count_useless(means: np.ndarray, weights: np.ndarray, n_components: int, allele_filter: float, p: OracleParams) -> int:
    """The peak filters of fit_gmm (allele.py:88-121): how many components of a fitted mixture are dropped."""
    keep_1 = weights > allele_filter
    order = np.argsort(means)
    strict = n_components > 2 or (n_components == 2 and (
        p.force_gm_filter or means[order[-1]] < p.expansion_ratio * max(float(means[order[0]]), SMALL_ALLELE_MIN)))
    keep_2 = weights > (1 / (p.filter_factor * n_components)) if strict else weights > F32_EPS
    return int(np.size(keep_1) - np.count_nonzero(keep_1 & keep_2))


def fit_gmm(rng: np.random.Generator, sample: np.ndarray, n_alleles: int, allele_filter: float, p: OracleParams):
    """allele.py:56-123.  `sample` is one sorted bootstrap replicate."""
    sample_rs = sample.reshape(-1, 1)
    if np.unique(sample).shape[0] == 1:  # Counter(sample).most_common(2) has one entry (:66-76)
        return _Single(sample_rs)
    n_components = n_alleles
    g = None
    while n_components > 0:
        if n_components == 1:
            return _Single(sample_rs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=ConvergenceWarning)
            g = GaussianMixture(n_components=n_components, init_params=p.init_params_method,
                                covariance_type="spherical", n_init=p.n_init, random_state=new_seed(rng)).fit(sample_rs)
        n_useless = count_useless(g.means_[:, 0], g.weights_, n_components, allele_filter, p)
        if not n_useless:
            return g
        n_components -= n_useless
    return g


def percentile_cis(samples: np.ndarray, is_99: bool) -> np.ndarray:
    """_calculate_cis (allele.py:50-54)."""
    q = (0.5, 99.5) if is_99 else (2.5, 97.5)
    return np.rint(np.percentile(samples, q, axis=1, method="interpolated_inverted_cdf").transpose()).astype(np.int32)


def aggregate(allele_samples: np.ndarray, weight_samples: np.ndarray, stdev_samples: np.ndarray, peaks: np.ndarray):
    """allele.py:295-336 on arrays [n_alleles, num_bootstrap] (+ peaks [num_bootstrap])."""
    order = allele_samples.argsort(axis=1, kind="stable")
    a = np.take_along_axis(allele_samples, order, axis=1)
    w = np.take_along_axis(weight_samples, order, axis=1)
    s = np.take_along_axis(stdev_samples, order, axis=1)
    mid = a.shape[1] // 2
    pw = w[:, mid].flatten()
    pw = pw / pw.sum()
    vals, counts = np.unique(np.sort(peaks, kind="stable"), return_counts=True)
    modal = int(vals[np.argmax(counts)])  # statistics.mode of the sorted list: the smallest of the most common
    return {"call": np.rint(a[:, mid]).astype(np.int32), "call_95_cis": percentile_cis(a, False),
            "call_99_cis": percentile_cis(a, True), "means": a[:, mid].flatten(), "weights": pw,
            "stdevs": s[:, mid].flatten(), "modal_n": modal}


def replicate_estimates(rng: np.random.Generator, cn: np.ndarray, w: np.ndarray, n_alleles: int, p: OracleParams):
    """Bootstrap + per-replicate GMM of call_alleles (allele.py:218-293): arrays [n_alleles, B] and peaks [B]."""
    n = cn.shape[0]
    if p.num_bootstrap < 2:  # the reference indexes a 1-D array with [i, :] in that case (allele.py:163-167,258)
        raise IndexError("num_bootstrap = 1 is not usable in the reference")
    reps = np.sort(rng.choice(cn, size=(p.num_bootstrap, n), replace=True, p=w), kind="stable")
    allele_filter = (p.min_allele_reads - 0.1) / reps.shape[0]  # sic: the number of replicates (:243)
    cache: dict[bytes, object] = {}
    m_all, w_all, s_all, peaks = [], [], [], []
    for i in range(p.num_bootstrap):
        rep = reps[i, :]
        key = rep.tobytes()
        if key not in cache:
            cache[key] = fit_gmm(rng, rep, n_alleles, allele_filter, p)
        g = cache[key]
        means = np.asarray(g.means_).reshape(-1)
        weights = np.asarray(g.weights_).reshape(-1)
        stdevs = np.sqrt(np.asarray(g.covariances_)).reshape(-1)
        peaks.append(means.shape[0])
        missing = n_alleles - means.shape[0]
        if missing:
            pick = rng.choice(np.arange(len(means)), size=missing, p=weights / np.abs(weights).sum())
            means, weights, stdevs = (np.append(v, v[pick]) for v in (means, weights, stdevs))
        o = np.argsort(means, kind="stable")
        m_all.append(means[o]), w_all.append(weights[o]), s_all.append(stdevs[o])
    return (np.array(m_all).T, np.array(w_all).T, np.array(s_all).T, np.array(peaks, dtype=np.int32))


def call_alleles(cn, w, n_alleles: int, min_reads: int, seed, p: OracleParams):
    """call_alleles (allele.py:176-336) for one locus; None as the reference returns it."""
    cn = np.asarray(cn, dtype=np.int32)
    w = np.asarray(w, dtype=np.float64)
    if cn.shape[0] < min_reads:
        return None
    if np.unique(cn).shape[0] == 1:
        c = np.full(n_alleles, cn[0], dtype=np.int32)
        ci = np.full((n_alleles, 2), cn[0], dtype=np.int32)
        return {"call": c, "call_95_cis": ci, "call_99_cis": ci, "means": c.astype(np.float64),
                "weights": np.full(n_alleles, 1.0 / n_alleles), "stdevs": np.full(n_alleles, 0.0), "modal_n": 1}
    rng = np.random.default_rng(seed=seed)
    return aggregate(*replicate_estimates(rng, cn, w, n_alleles, p))
