"""Shared test helpers (CPU reference implementations used as checkers only)."""
from types import SimpleNamespace

import numpy as np


def reference_recall_loop(predictions, positives_per_query, recall_values):
    """Verbatim semantics of reference test.py:75-82."""
    recalls = np.zeros(len(recall_values))
    for query_index, pred in enumerate(predictions):
        for i, n in enumerate(recall_values):
            if np.any(np.isin(pred[:n], positives_per_query[query_index])):
                recalls[i:] += 1
                break
    return recalls / len(predictions) * 100


def numpy_merge(D_lists, d_stride, I_lists, i_stride, nq, k, n_lists, id_bound, metric=1):
    """CPU checker for the cross-shard merge (same argument convention as sharded._cuda_merge); metric 0 = inner
    product (largest first, padding -FLT_MAX), 1 = L2."""
    import torch
    Dl = D_lists.cpu().numpy().reshape(-1)
    Il = I_lists.cpu().numpy().reshape(-1)
    sign = -1.0 if metric == 0 else 1.0
    D = np.full((nq, k), np.float32(sign * 3.4028234663852886e38), np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        ds = np.concatenate([Dl[g * d_stride + q * k: g * d_stride + (q + 1) * k] for g in range(n_lists)])
        ids = np.concatenate([Il[g * i_stride + q * k: g * i_stride + (q + 1) * k] for g in range(n_lists)])
        keep = ids >= 0
        ds, ids = ds[keep], ids[keep]
        order = np.lexsort((ids, sign * ds))[:k]
        D[q, :len(order)] = ds[order]
        I[q, :len(order)] = ids[order]
    return torch.from_numpy(D), torch.from_numpy(I)


def make_mining_problem(seed, database_num=600, queries_num=80, d=32):
    """Descriptors + UTM-derived hard/soft positives shaped like the reference's TripletsDataset state."""
    from agplace_b200 import synth
    rng = np.random.default_rng(seed)
    xb = synth.descriptors(database_num, d, seed, "db")
    xq = synth.descriptors(queries_num, d, seed + 1, "q")
    db_utm, q_utm = synth.utm_positions(database_num, queries_num, 300.0, seed)
    hard = synth.radius_positives(db_utm, q_utm, 10.0)
    soft = synth.radius_positives(db_utm, q_utm, 25.0)
    # the reference drops queries without hard positives (datasets_ws_kitti360.py:750-760): emulate by
    # giving such queries their nearest database point
    for i in range(queries_num):
        if len(hard[i]) == 0:
            j = int(np.argmin(((db_utm - q_utm[i]) ** 2).sum(1)))
            hard[i] = np.array([j], dtype=np.int64)
            soft[i] = np.union1d(soft[i], hard[i])
    cache = np.concatenate([xb, xq]).astype(np.float32)
    return SimpleNamespace(xb=xb, xq=xq, hard=hard, soft=soft, cache=cache, database_num=database_num,
                           queries_num=queries_num, d=d, rng=rng)


def faiss_tutorial_data():
    """The data of faiss's own tutorial (tutorial/python/1-Flat.py): legacy np.random.seed(1234) stream, which is
    stable across numpy versions.  Returns (xb, xq, published) with the published outputs of real faiss."""
    import json
    from pathlib import Path
    pub = json.loads((Path(__file__).parent / "golden" / "faiss_tutorial_1flat.json").read_text())
    d, nb, nq = 64, 100000, 10000
    state = np.random.get_state()
    try:
        np.random.seed(1234)
        xb = np.random.random((nb, d)).astype("float32")
        xb[:, 0] += np.arange(nb) / 1000.0
        xq = np.random.random((nq, d)).astype("float32")
        xq[:, 0] += np.arange(nq) / 1000.0
    finally:
        np.random.set_state(state)
    return xb, xq, pub
