"""Helpers for the allele-calling tests: synthetic loci and a k-means++ seeding hook for scikit-learn."""
from __future__ import annotations

import contextlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "alleles_golden.json")) as fh:
        return json.load(fh)


def synthetic_loci(rng, n_loci, n_reads=(8, 45), expansions=0.0, single=0.0):
    """Per-read copy numbers of diploid loci (two alleles, +-1 stutter), weights normalised per locus."""
    cn, w, rb = [], [], [0]
    for _ in range(n_loci):
        n = int(rng.integers(n_reads[0], n_reads[1] + 1))
        a1 = int(rng.integers(8, 60))
        u = rng.random()
        if u < single:
            x = np.full(n, a1)
        elif u < single + expansions:
            a2 = a1 * int(rng.integers(6, 25))
            x = np.where(rng.random(n) < 0.2, a2 + rng.integers(-15, 16, size=n),
                         a1 + rng.choice([-1, 0, 1], size=n, p=[.05, .9, .05]))
        else:
            a2 = a1 + int(rng.choice([0, 0, 1, 1, 2, 3, 5]))
            x = rng.choice([a1, a2], size=n) + rng.choice([-1, 0, 1], size=n, p=[.04, .92, .04])
        ww = rng.uniform(0.6, 1.4, size=n)
        cn.append(x.astype(np.int32))
        w.append(ww / ww.sum())
        rb.append(rb[-1] + n)
    return np.concatenate(cn), np.concatenate(w), np.asarray(rb, dtype=np.int64)


@contextlib.contextmanager
def forced_kmeanspp(point_index_pairs):
    """Make sklearn's GaussianMixture(init_params="k-means++") use the given seed points, one pair per restart
    (sklearn.mixture._base.BaseMixture._initialize_parameters calls kmeans_plusplus once per restart)."""
    import sklearn.mixture._base as base

    it = iter(point_index_pairs)
    orig = base.kmeans_plusplus

    def fake(X, n_clusters, *, random_state=None, **kw):
        idx = np.asarray(next(it), dtype=np.int64)
        return X[idx], idx

    base.kmeans_plusplus = fake
    try:
        yield
    finally:
        base.kmeans_plusplus = orig
