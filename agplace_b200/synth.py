"""Seeded synthetic inputs shaped like BASELINE.json's configs (no datasets are reachable).

Descriptor distributions follow SURVEY.md section 8(d): database rows are unit-norm (the reference's
database model L2-normalises: models_baseline/dbvanilla2d.py:79-84), query rows have norm
U(0.8, 1.2) (``final_l2=False``, tools/options.py:118).  ``clustered`` makes each query a noisy
copy of a database row -- the cancellation regime used by the adversarial parity tests.
Positives are the database items within ``radius`` metres of the query in UTM space, the same
rule as the reference's ``NearestNeighbors.radius_neighbors`` (datasets_ws_kitti360.py:613-618).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (db rows, queries, dim, k, seed, area side in metres)
    "cfg1": dict(n=10_000, nq=2_000, d=256, k=20, seed=0, side=1_000.0,
                 desc="KITTI-360-AG-shaped: 10k aerial DB x 2k ground queries x 256-d, top-20 + recall@1/5/10"),
    "cfg2": dict(n=100_000, nq=20_000, d=512, k=50, seed=1, side=3_162.0,
                 desc="nuScenes-AG-shaped: 100k DB x 20k queries x 512-d, top-50, UTM-radius positives"),
    "cfg3": dict(n=100_000, nq=1_000, d=512, k=10, seed=2, side=3_162.0,
                 desc="partial hard-negative mining: 1000 cached queries x 100k-negative cache x 512-d, top-10"),
    "cfg4": dict(n=10_000_000, nq=100_000, d=512, k=100, seed=3, side=31_620.0,
                 desc="large-scale DB 10M x 512-d sharded across 1/2/4/8 B200, 100k queries top-100"),
    "cfg5": dict(n=1_000_000, nq=10_000, d=4096, k=100, seed=4, side=10_000.0,
                 desc="high-dim aggregation descriptors: 1M DB x 4096-d, 10k queries (harness choice), top-100"),
}


def descriptors(n, d, seed, kind="db", dtype=np.float32, chunk=65536):
    """iid N(0,1) rows, L2-normalised (db) or scaled to norm U(0.8,1.2) (queries)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=dtype)
    for i0 in range(0, n, chunk):
        x = rng.standard_normal((min(chunk, n - i0), d), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
        if kind != "db":
            x *= rng.uniform(0.8, 1.2, size=(x.shape[0], 1)).astype(np.float32)
        out[i0:i0 + x.shape[0]] = x
    return out


def clustered_queries(xb, nq, sigma, seed):
    """Each query = a random database row + sigma * N(0, I/d), renormalised (near-duplicate regime)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, xb.shape[0], size=nq)
    q = xb[src] + (sigma / np.sqrt(xb.shape[1])) * rng.standard_normal((nq, xb.shape[1])).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True) + 1e-12
    return q.astype(np.float32), src


def utm_positions(n, nq, side, seed, sigma_m=5.0):
    """Database UTM ~ uniform on a side x side square; query UTM = a database point + N(0, sigma_m)."""
    rng = np.random.default_rng(seed + 1000)
    db = rng.uniform(0.0, side, size=(n, 2))
    src = rng.integers(0, n, size=nq)
    q = db[src] + rng.normal(0.0, sigma_m, size=(nq, 2))
    return db, q


def radius_positives(db_utm, q_utm, radius):
    """Object array of int64 index arrays, like sklearn ``radius_neighbors(..., return_distance=False)``."""
    try:
        from sklearn.neighbors import NearestNeighbors
        knn = NearestNeighbors(n_jobs=-1)
        knn.fit(db_utm)
        return knn.radius_neighbors(q_utm, radius=radius, return_distance=False)
    except Exception:
        out = np.empty(len(q_utm), dtype=object)
        for i, p in enumerate(q_utm):
            out[i] = np.nonzero(((db_utm - p) ** 2).sum(1) <= radius * radius)[0].astype(np.int64)
        return out


@dataclass
class SyntheticEvalSet:
    """Stand-in for the reference's ``*BaseDataset`` as far as ``compute_recall`` needs it."""
    database_features: np.ndarray
    queries_features: np.ndarray
    positives_per_query: np.ndarray
    database_num: int
    queries_num: int

    def get_positives(self):
        return self.positives_per_query


def make_eval_set(name_or_cfg, radius=25.0, correlated=0.0, scale=1.0):
    """Build descriptors + UTM positives for a config.  ``correlated`` > 0 pulls each query's
    descriptor towards the descriptor of its UTM source row so that recall is not trivially 0."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    n, nq = max(1, int(cfg["n"] * scale)), max(1, int(cfg["nq"] * scale))
    xb = descriptors(n, cfg["d"], cfg["seed"], "db")
    xq = descriptors(nq, cfg["d"], cfg["seed"] + 7, "q")
    db_utm, q_utm = utm_positions(n, nq, cfg["side"] * np.sqrt(scale), cfg["seed"])
    positives = radius_positives(db_utm, q_utm, radius)
    if correlated > 0:
        rng = np.random.default_rng(cfg["seed"] + 99)
        for i in range(nq):
            if len(positives[i]):
                j = positives[i][rng.integers(0, len(positives[i]))]
                xq[i] = (1 - correlated) * xq[i] + correlated * xb[j] * np.linalg.norm(xq[i])
    return SyntheticEvalSet(xb, xq, positives, n, nq)
