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


# ------------------------------------------------------------------------------------------ counter-based rows
# Databases too large to ship from the host (cfg4: 10 M x 512 = 20.5 GB; cfg5: 1 M x 4096) are generated on the GPU that
# owns the shard.  So that ANY row can be regenerated bit-for-bit on the CPU (sampled verification of a 10 M-row search
# against the oracle), element (row, col) is a pure function of (seed, row * d + col):
#   z = splitmix64(seed * GOLDEN + row * d + col);  a = sum of z's four 16-bit fields - 131070   (Irwin-Hall, n = 4)
#   x = float32(a) * float32(1 / (sigma_a * sqrt(d)))      -- one correctly rounded fp32 multiply of an exact integer
# i.e. integer arithmetic plus a single IEEE operation: identical bits from numpy and from torch on CUDA.  Rows are
# approximately unit-norm (|x| = 1 +- 1.3/sqrt(d)) with a bell-shaped, bounded element distribution.
_SM_GOLDEN, _SM_M1, _SM_M2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
_IH_SIGMA = float(np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))


def _counter_scale(d):
    return np.float32(1.0 / (_IH_SIGMA * np.sqrt(float(d))))


def counter_rows(row_ids, d, seed, chunk=8192):
    """Rows ``row_ids`` (any int64 array) of the counter-based database ``seed``, on the CPU (numpy; chunks run on a
    small thread pool -- numpy releases the GIL inside the integer kernels)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    row_ids = np.asarray(row_ids, dtype=np.int64).reshape(-1)
    out = np.empty((len(row_ids), d), dtype=np.float32)
    cols = np.arange(d, dtype=np.uint64)[None, :]
    base = np.uint64((seed * _SM_GOLDEN) & 0xFFFFFFFFFFFFFFFF)
    c = _counter_scale(d)
    m = np.uint64(0xFFFF)

    def fill(i0):
        with np.errstate(over="ignore"):
            r = row_ids[i0:i0 + chunk].astype(np.uint64)[:, None]
            z = r * np.uint64(d) + cols + base
            z = (z ^ (z >> np.uint64(30))) * np.uint64(_SM_M1)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(_SM_M2)
            z = z ^ (z >> np.uint64(31))
            a = ((z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + (z >> np.uint64(48))).astype(np.int64) - 131070
            out[i0:i0 + chunk] = a.astype(np.float32) * c

    starts = range(0, len(row_ids), chunk)
    workers = min(len(starts), max(1, (os.cpu_count() or 2) // 2), 16)
    if workers <= 1:
        for i0 in starts:
            fill(i0)
    else:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(fill, starts))
    return out


def counter_rows_device(row0, row1, d, seed, device):
    """Rows [row0, row1) of the same database as a CUDA tensor (torch integer ops: bit-identical to counter_rows)."""
    import torch

    def s64(v):          # two's-complement int64 view of a 64-bit constant
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(z, s):       # logical shift right of an int64 tensor
        return (z >> s) & ((1 << (64 - s)) - 1)

    r = torch.arange(row0, row1, dtype=torch.int64, device=device)[:, None]
    z = r * d + torch.arange(d, dtype=torch.int64, device=device)[None, :] + s64(seed * _SM_GOLDEN)
    z = (z ^ lsr(z, 30)) * s64(_SM_M1)
    z = (z ^ lsr(z, 27)) * s64(_SM_M2)
    z = z ^ lsr(z, 31)
    a = (z & 0xFFFF) + ((z >> 16) & 0xFFFF) + ((z >> 32) & 0xFFFF) + ((z >> 48) & 0xFFFF) - 131070
    return a.to(torch.float32) * float(_counter_scale(d))


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
