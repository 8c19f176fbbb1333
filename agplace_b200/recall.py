"""Host-side mirror of the reference's recall@N evaluation (call site A).

``compute_recall`` follows reference test.py:24-84 for the default ``test_method='hard_resize'``
path (the ``nearest_crop`` / ``maj_voting`` branches are dead code in the reference: test.py:160
uses an undefined ``images``), with the index class injectable so the same function runs against
the CUDA engine (default) or, in tests, against the CPU oracle.

``compute_recall_device`` is the "next" row N1 of SURVEY.md section 8(f): descriptors that are
already CUDA tensors are searched and scored (K5) without leaving the GPU; only the
``len(recall_values)`` hit counters come back.
"""
from __future__ import annotations

import numpy as np

from .index import IndexFlatL2, recall_hits


def compute_recall(args, queries_features, database_features, test_ds, test_method="hard_resize", index_cls=None,
                   on_device_recall=False):
    """Returns ``(recalls, recalls_str)`` exactly like reference test.py:24-84.

    ``args`` needs ``features_dim`` and ``recall_values``; ``test_ds`` needs ``queries_num`` and
    ``get_positives()`` (object array of per-query positive database ids)."""
    if test_method != "hard_resize":
        raise NotImplementedError("only the reference's live 'hard_resize' path is mirrored")
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
