"""Host-side mirror of the reference's recall@N evaluation (call site A).

``compute_recall`` follows reference test.py:24-84 for every test_method that takes the plain
search + recall path ('hard_resize', 'single_query', 'central_crop', 'five_crops'; the
``nearest_crop`` / ``maj_voting`` branches are dead code in the reference: test.py:160
uses an undefined ``images``), with the index class injectable so the same function runs against
the CUDA engine (default) or, in tests, against the CPU oracle.

``compute_recall_device`` is the "next" row N1 of SURVEY.md section 8(f): descriptors that are
already CUDA tensors are searched and scored (K5) without leaving the GPU; only the
``len(recall_values)`` hit counters come back.
"""
from __future__ import annotations

import numpy as np

from .index import IndexFlatIP, IndexFlatL2, recall_hits


def compute_recall(args, queries_features, database_features, test_ds, test_method="hard_resize", index_cls=None,
                   on_device_recall=False):
    """Returns ``(recalls, recalls_str)`` exactly like reference test.py:24-84.

    ``args`` needs ``features_dim`` and ``recall_values``; ``test_ds`` needs ``queries_num`` and
    ``get_positives()`` (object array of per-query positive database ids)."""
    # test.py:24-84 special-cases only 'nearest_crop' and 'maj_voting' (five descriptors per query, dead code in the
    # reference: test.py:160 uses an undefined ``images``); every other test_method ('hard_resize', 'single_query',
    # 'central_crop', 'five_crops' after its mean over crops) goes through the plain search + recall below
    assert test_method in ["hard_resize", "single_query", "central_crop", "five_crops", "nearest_crop", "maj_voting"], \
        f"test_method can't be {test_method}"                       # test.py:92-93
    if test_method in ("nearest_crop", "maj_voting"):
        raise NotImplementedError(f"test_method={test_method!r} needs five descriptors per query; that branch of the "
                                  "reference is dead code (test.py:160) and is not mirrored")
    index_cls = index_cls or IndexFlatL2
    faiss_index = index_cls(args.features_dim)                      # test.py:27
    faiss_index.add(database_features)                              # test.py:28
    distances, predictions = faiss_index.search(queries_features, max(args.recall_values))   # test.py:32

    positives_per_query = test_ds.get_positives()                   # test.py:73
    if on_device_recall:
        hits = recall_hits(predictions, positives_per_query, args.recall_values)
        recalls = hits.astype(np.float64)
    else:
        # test.py:75-80, verbatim semantics (np.in1d is deprecated in numpy 2: np.isin is the same test)
        recalls = np.zeros(len(args.recall_values))
        for query_index, pred in enumerate(np.asarray(predictions)):
            for i, n in enumerate(args.recall_values):
                if np.any(np.isin(pred[:n], positives_per_query[query_index])):
                    recalls[i:] += 1
                    break
    recalls = recalls / test_ds.queries_num * 100                   # test.py:82
    recalls_str = ", ".join([f"R@{val}: {rec:.1f}" for val, rec in zip(args.recall_values, recalls)])
    return recalls, recalls_str


def compute_recall_device(features_dim, recall_values, queries_features, database_features, positives_per_query, index=None):
    """GPU-resident variant: CUDA tensors in, recall percentages out (fp64 numpy array)."""
    index = index or IndexFlatL2(features_dim, device=database_features.device.index)
    if index.ntotal == 0:
        index.add(database_features)
    _, predictions = index.search(queries_features, max(recall_values))
    hits = recall_hits(predictions, positives_per_query, recall_values)
    return hits.astype(np.float64) / queries_features.shape[0] * 100


def get_top_k_recall(top_k, db, qu, gt_pos, method="cosine", norm_descs=True, use_gpu=True, use_percentage=True,
                     sub_sample_db=1, sub_sample_qu=1):
    """Mirror of reference anyloc/utilities.py:396-475: ``IndexFlatIP`` for ``method='cosine'``,
    ``IndexFlatL2`` for ``'l2'``; descriptors are torch tensors (CPU or CUDA) or arrays.  The reference's
    ``use_gpu=True`` branch needs faiss-gpu; here every index is a GPU index, so the flag changes nothing.
    Returns ``(distances, indices, recalls)`` with ``recalls`` a dict keyed by the ``top_k`` values."""
    import torch
    import torch.nn.functional as F
    db = torch.as_tensor(db)
    qu = torch.as_tensor(qu)
    if len(qu.shape) == 1:
        qu = qu.unsqueeze(0)
    if norm_descs:
        db = F.normalize(db)
        qu = F.normalize(qu)
    D = db.shape[1]
    if method == "cosine":
        index = IndexFlatIP(D)
    elif method == "l2":
        index = IndexFlatL2(D)
    else:
        raise NotImplementedError(f"Method: {method}")
    index.add(db)
    distances, indices = index.search(qu, max(top_k))
    recalls = dict(zip(top_k, [0] * len(top_k)))
    ind_host = indices.cpu().numpy() if hasattr(indices, "cpu") else np.asarray(indices)
    for i_qu, qu_retr in enumerate(ind_host):
        for i_rec in top_k:
            correct_retr = gt_pos[i_qu * sub_sample_qu]
            if np.any(np.isin(qu_retr[:i_rec] * sub_sample_db, correct_retr)):
                recalls[i_rec] += 1
    if use_percentage:
        for k in recalls:
            recalls[k] /= len(ind_host)
    return distances, indices, recalls
