"""Host-side mirror of the reference's hard-negative mining helpers (call sites B and C).

Restates, with an injectable index class, the faiss-backed pieces of
``datasets/datasets_ws_kitti360.py`` (identical copies live in ``datasets_ws_nuscenes.py:1241-1258``
and ``datasets_ws.py:689-706``):

* ``get_best_positive_index``         kitti360:976-983   (fresh index, k = 1 over the hard positives)
* ``get_hardest_negatives_indexes``   kitti360:985-993   (fresh index, k = negs_num_per_query)
* the per-query loops of ``compute_triplets_partial[_sep]`` (kitti360:1056-1137) and
  ``compute_triplets_full`` (kitti360:1022-1049), including the order of ``np.random`` draws, so
  that the mined ``triplets_global_indexes`` are identical for identical descriptors.

``compute_triplets_partial_batched`` / ``compute_triplets_full_batched`` produce the SAME
``triplets_global_indexes`` -- same RNG draws, same arithmetic (exact fp32 difference form), same tie order -- with
two GPU calls per refresh instead of two index constructions + two searches per query: one
``best_of_lists`` over every query's own hard positives and one ``IndexFlatL2.search_masked`` over the sampled
database rows with the per-query soft positives masked out.

Feature extraction (``compute_cache*``) is the caller's business: ``cache`` is any object whose
``cache[i]`` / ``cache[index_array]`` returns fp32 rows, e.g. :class:`RAMEfficient2DMatrix`
(kitti360:1147-1167) or a plain ndarray.
"""
from __future__ import annotations

import numpy as np

from .index import IndexFlatL2, best_of_lists


class RAMEfficient2DMatrix:
    """List-of-rows matrix with numpy-style row gathers (reference kitti360:1147-1167)."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = shape
        self.dtype = dtype
        self.matrix = [None] * shape[0]

    def __setitem__(self, indexes, vals):
        assert vals.shape[1] == self.shape[1], f"{vals.shape[1]} {self.shape[1]}"
        for i, val in zip(indexes, vals):
            self.matrix[i] = val.astype(self.dtype, copy=False)

    def __getitem__(self, index):
        if hasattr(index, "__len__"):
            return np.array([self.matrix[i] for i in index])
        return self.matrix[index]


class TripletMiner:
    """The mining state the reference keeps on its ``*TripletsDataset`` objects."""

    def __init__(self, features_dim, database_num, queries_num, hard_positives_per_query, soft_positives_per_query,
                 negs_num_per_query=10, neg_samples_num=1000, index_cls=None):
        self.features_dim = int(features_dim)
        self.database_num = int(database_num)
        self.queries_num = int(queries_num)
        self.hard_positives_per_query = hard_positives_per_query
        self.soft_positives_per_query = soft_positives_per_query
        self.negs_num_per_query = int(negs_num_per_query)
        self.neg_samples_num = int(neg_samples_num)
        self.index_cls = index_cls or IndexFlatL2
        # compute_triplets_full keeps the previous hardest negatives per query (reference neg_cache)
        self.neg_cache = [np.empty((0,), dtype=np.int32) for _ in range(self.queries_num)]
        self.triplets_global_indexes = None

    # ---- kitti360:965-974
    def get_query_features(self, query_index, cache):
        query_features = cache[query_index + self.database_num]
        if query_features is None:
            raise RuntimeError(f"For query with index {query_index} features have not been computed!\n"
                               "There might be some bug with caching")
        return query_features

    # ---- kitti360:976-983
    def get_best_positive_index(self, query_index, cache, query_features):
        positives_features = cache[self.hard_positives_per_query[query_index]]
        faiss_index = self.index_cls(self.features_dim)
        faiss_index.add(positives_features)
        # Search the best positive (within 10 meters AND nearest in features space)
        _, best_positive_num = faiss_index.search(query_features.reshape(1, -1), 1)
        best_positive_index = self.hard_positives_per_query[query_index][best_positive_num[0]].item()
        return best_positive_index

    # ---- kitti360:985-993
    def get_hardest_negatives_indexes(self, cache, query_features, neg_samples):
        neg_features = cache[neg_samples]
        faiss_index = self.index_cls(self.features_dim)
        faiss_index.add(neg_features)
        # Search the 10 nearest negatives (further than 25 meters and nearest in features space)
        _, neg_nums = faiss_index.search(query_features.reshape(1, -1), self.negs_num_per_query)
        neg_nums = neg_nums.reshape(-1)
        neg_indexes = neg_samples[neg_nums].astype(np.int32)
        return neg_indexes

    # ---- kitti360:1056-1093 / 1099-1137 (the loop after the cache has been computed)
    def compute_triplets_partial(self, cache, cache_refresh_rate):
        """``cache`` must hold every database row in ``sampled_database_indexes`` / the hard positives
        and every sampled query; computing it is the caller's job.  Returns int64 [refresh, 2 + negs]."""
        triplets = []
        sampled_queries_indexes = np.random.choice(self.queries_num, cache_refresh_rate, replace=False)
        sampled_database_indexes = np.random.choice(self.database_num, self.neg_samples_num, replace=False)
        for query_index in sampled_queries_indexes:
            query_features = self.get_query_features(query_index, cache)
            best_positive_index = self.get_best_positive_index(query_index, cache, query_features)
            soft_positives = self.soft_positives_per_query[query_index]
            neg_indexes = np.setdiff1d(sampled_database_indexes, soft_positives, assume_unique=True)
            neg_indexes = self.get_hardest_negatives_indexes(cache, query_features, neg_indexes)
            triplets.append((query_index, best_positive_index, *neg_indexes))
        self.triplets_global_indexes = np.asarray(triplets, dtype=np.int64)
        return self.triplets_global_indexes

    # ---- kitti360:1022-1049
    def compute_triplets_full(self, cache, cache_refresh_rate):
        triplets = []
        sampled_queries_indexes = np.random.choice(self.queries_num, cache_refresh_rate, replace=False)
        for query_index in sampled_queries_indexes:
            query_features = self.get_query_features(query_index, cache)
            best_positive_index = self.get_best_positive_index(query_index, cache, query_features)
            neg_indexes = np.random.choice(self.database_num, self.neg_samples_num, replace=False)
            soft_positives = self.soft_positives_per_query[query_index]
            neg_indexes = np.setdiff1d(neg_indexes, soft_positives, assume_unique=True)
            neg_indexes = np.unique(np.concatenate([self.neg_cache[query_index], neg_indexes]))
            neg_indexes = self.get_hardest_negatives_indexes(cache, query_features, neg_indexes)
            self.neg_cache[query_index] = neg_indexes
            triplets.append((query_index, best_positive_index, *neg_indexes))
        self.triplets_global_indexes = np.asarray(triplets, dtype=np.int64)
        return self.triplets_global_indexes

    # ------------------------------------------------------------------ batched (N2)
    def _best_positives_batched(self, sampled_queries_indexes, cache, query_features):
        """get_best_positive_index for every sampled query in one kernel launch."""
        lens = np.array([len(self.hard_positives_per_query[q]) for q in sampled_queries_indexes], dtype=np.int64)
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        all_pos = np.concatenate([np.asarray(self.hard_positives_per_query[q]).reshape(-1) for q in sampled_queries_indexes])
        rows = np.asarray(cache[all_pos], dtype=np.float32)
        _, best = best_of_lists(query_features, rows, offsets)
        return np.array([self.hard_positives_per_query[q][b] for q, b in zip(sampled_queries_indexes, best)], dtype=np.int64)

    def compute_triplets_partial_batched(self, cache, cache_refresh_rate):
        """Same result as :meth:`compute_triplets_partial` (kitti360:1056-1137): the sampled database rows are
        indexed ONCE, in the order of the random draw -- ``np.setdiff1d(..., assume_unique=True)`` does not sort, so
        that is the order (and the tie-break order) of every per-query subset in the reference -- each query's soft
        positives inside the sample become its exclusion list, and positions map back through the sample."""
        sampled_queries_indexes = np.random.choice(self.queries_num, cache_refresh_rate, replace=False)
        sampled_database_indexes = np.random.choice(self.database_num, self.neg_samples_num, replace=False)
        query_features = np.stack([self.get_query_features(q, cache) for q in sampled_queries_indexes]).astype(np.float32)
        best_pos = self._best_positives_batched(sampled_queries_indexes, cache, query_features)
        sample = np.asarray(sampled_database_indexes)
        index = self.index_cls(self.features_dim)
        index.add(np.asarray(cache[sample], dtype=np.float32))
        exclude = []
        for q in sampled_queries_indexes:
            soft = np.asarray(self.soft_positives_per_query[q]).reshape(-1)
            exclude.append(np.flatnonzero(np.isin(sample, soft)).astype(np.int64))
        # the engine answers k + longest exclusion list <= 512 candidates per query (agp_index_search_masked); a query
        # with more soft positives inside the sample than that takes the reference's own per-query route instead
        k = self.negs_num_per_query
        max_k = getattr(index, "MAX_MASKED", 512)
        fits = np.array([k + len(e) <= max_k or len(sample) <= max_k for e in exclude], dtype=bool)
        I = np.full((len(exclude), k), -1, dtype=np.int64)
        if fits.any():
            rows = np.flatnonzero(fits)
            _, I[rows] = index.search_masked(query_features[rows], k, [exclude[r] for r in rows])
        for r in np.flatnonzero(~fits):
            keep = np.ones(len(sample), dtype=bool)
            keep[exclude[r]] = False
            sub = self.index_cls(self.features_dim)
            sub.add(np.asarray(cache[sample[keep]], dtype=np.float32))
            _, pos = sub.search(query_features[r].reshape(1, -1), k)
            I[r] = np.where(pos[0] >= 0, np.flatnonzero(keep)[np.maximum(pos[0], 0)], -1)
        negs = sample[I].astype(np.int32)
        # I == -1 (fewer than negs_num_per_query candidates): the reference's numpy indexing neg_samples[neg_nums] wraps
        # to the LAST element of that query's own subset, i.e. the last sampled row that is not excluded
        short = np.flatnonzero((I < 0).any(axis=1))
        for r in short:
            keep = np.ones(len(sample), dtype=bool)
            keep[exclude[r]] = False
            subset = sample[keep]
            if len(subset) == 0:
                raise IndexError("index -1 is out of bounds for axis 0 with size 0")      # what the reference raises
            negs[r, I[r] < 0] = subset[-1]
        self.triplets_global_indexes = np.concatenate(
            [sampled_queries_indexes.reshape(-1, 1).astype(np.int64), best_pos.reshape(-1, 1), negs.astype(np.int64)], axis=1)
        return self.triplets_global_indexes

    def compute_triplets_full_batched(self, cache, cache_refresh_rate):
        """Same result as :meth:`compute_triplets_full` (kitti360:1022-1049), including the ``neg_cache`` it leaves
        behind.  The global ``np.random`` stream is consumed in the reference's order (the query draw, then one
        database draw per query inside the loop -- neither search consumes random numbers), the whole database is
        indexed ONCE, and every query's own sorted-unique candidate set (``np.unique(concatenate([neg_cache[q],
        setdiff1d(draw, soft_positives)]))``) goes through one ``IndexFlatL2.search_subset`` call: exact fp32
        difference form, ties by position in the sorted set = by database id, like a fresh index over the set."""
        sampled_queries_indexes = np.random.choice(self.queries_num, cache_refresh_rate, replace=False)
        candidates = []
        for query_index in sampled_queries_indexes:
            neg_indexes = np.random.choice(self.database_num, self.neg_samples_num, replace=False)
            soft_positives = self.soft_positives_per_query[query_index]
            neg_indexes = np.setdiff1d(neg_indexes, soft_positives, assume_unique=True)
            candidates.append(np.unique(np.concatenate([self.neg_cache[query_index], neg_indexes])))
        query_features = np.stack([self.get_query_features(q, cache) for q in sampled_queries_indexes]).astype(np.float32)
        best_pos = self._best_positives_batched(sampled_queries_indexes, cache, query_features)
        index = self.index_cls(self.features_dim)
        index.add(np.asarray(cache[np.arange(self.database_num)], dtype=np.float32))
        _, I = index.search_subset(query_features, self.negs_num_per_query, candidates)
        negs = np.empty((len(sampled_queries_indexes), self.negs_num_per_query), dtype=np.int32)
        for r, (query_index, cand) in enumerate(zip(sampled_queries_indexes, candidates)):
            negs[r] = cand[I[r]].astype(np.int32)        # -1 wraps to the last candidate, as in the reference
            self.neg_cache[query_index] = negs[r].copy()
        self.triplets_global_indexes = np.concatenate(
            [sampled_queries_indexes.reshape(-1, 1).astype(np.int64), best_pos.reshape(-1, 1), negs.astype(np.int64)], axis=1)
        return self.triplets_global_indexes
