"""``faiss.Kmeans`` on the engine (reference model/aggregation.py:148-173, the NetVLAD initialisation:
``kmeans = faiss.Kmeans(descs_dim, clusters_num, niter=100, verbose=False); kmeans.train(descriptors);
kmeans.centroids``).

Lloyd iterations exactly as faiss's ``Clustering::train`` runs them -- optional subsampling to
``max_points_per_centroid`` points per centroid, centroids initialised with distinct random training points, then per
iteration: assignment = exact L2 top-1 search of every point against the centroids (the hot part: the same
``IndexFlatL2`` kernels as the retrieval path, device-resident), centroid = mean of its points, empty clusters re-seeded
by splitting a large one with a +-1/1024 perturbation.  The random choices come from ``numpy.random.RandomState(seed)``,
not from faiss's own generator, so centroids agree with faiss statistically (same objective up to the usual k-means
variance), not bit for bit; everything after the draw is deterministic.
"""
from __future__ import annotations

import numpy as np

from .index import IndexFlatL2


class Kmeans:
    def __init__(self, d, k, niter=25, nredo=1, verbose=False, seed=1234, max_points_per_centroid=256,
                 min_points_per_centroid=39, spherical=False, device=None, gpu=True):
        self.d = int(d)
        self.k = int(k)
        self.niter = int(niter)
        self.nredo = int(nredo)
        self.verbose = bool(verbose)
        self.seed = int(seed)
        self.max_points_per_centroid = int(max_points_per_centroid)
        self.min_points_per_centroid = int(min_points_per_centroid)
        self.spherical = bool(spherical)
        self.device = device
        self.centroids = None
        self.obj = np.empty(0, dtype=np.float32)
        self.iteration_stats = []
        self.index = None

    # faiss Clustering::train
    def train(self, x, weights=None, init_centroids=None):
        import torch
        if weights is not None:
            raise NotImplementedError("weighted k-means is not part of the reference's use")
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, d = x.shape
        assert d == self.d
        if n < self.k:
            raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({self.k})")
        rs = np.random.RandomState(self.seed)
        if n > self.k * self.max_points_per_centroid:           # subsample_training_set
            keep = rs.permutation(n)[: self.k * self.max_points_per_centroid]
            x = x[keep]
            n = len(x)
        probe = IndexFlatL2(d, device=self.device)
        dev = torch.device("cuda", probe.device)
        xd = torch.from_numpy(x).to(dev)
        best = None
        for redo in range(max(self.nredo, 1)):
            if init_centroids is not None and redo == 0:
                cent = torch.from_numpy(np.ascontiguousarray(init_centroids, dtype=np.float32)).to(dev).clone()
                assert tuple(cent.shape) == (self.k, d)
            else:
                cent = xd[torch.from_numpy(rs.permutation(n)[: self.k]).to(dev)].clone()
            objs, stats = [], []
            for it in range(self.niter):
                index = IndexFlatL2(d, device=probe.device)
                index.add(cent)
                D, I = index.search(xd, 1)                        # assignment: exact L2 top-1 on the engine
                assign = I[:, 0]
                obj = float(D[:, 0].sum().item())
                counts = torch.bincount(assign, minlength=self.k).to(torch.float32)
                sums = torch.zeros((self.k, d), dtype=torch.float32, device=dev).index_add_(0, assign, xd)
                nonempty = counts > 0
                cent = torch.where(nonempty[:, None], sums / counts.clamp(min=1.0)[:, None], cent)
                nsplit = self._split_empty(cent, counts, rs)
                if self.spherical:
                    cent = cent / cent.norm(dim=1, keepdim=True).clamp(min=1e-30)
                objs.append(obj)
                stats.append({"obj": obj, "nsplit": nsplit})
                if self.verbose:
                    print(f"  Iteration {it} objective={obj:g} nsplit={nsplit}")
            if best is None or objs[-1] < best[0][-1]:
                best = (objs, stats, cent.clone())
        objs, stats, cent = best
        self.obj = np.asarray(objs, dtype=np.float32)
        self.iteration_stats = stats
        self.centroids = cent.cpu().numpy()
        self.index = IndexFlatL2(d, device=probe.device)
        self.index.add(self.centroids)
        return float(self.obj[-1]) if len(self.obj) else 0.0

    @staticmethod
    def _split_empty(cent, counts, rs, eps=1.0 / 1024.0):
        """faiss ``split_clusters``: every empty cluster takes over half of a populated one (chosen with probability
        proportional to its size), the two copies pushed apart by a symmetric +-eps perturbation."""
        import torch
        empty = torch.nonzero(counts == 0).flatten().tolist()
        if not empty:
            return 0
        c = counts.cpu().numpy().astype(np.float64)
        sign = torch.ones(cent.shape[1], device=cent.device)
        sign[1::2] = -1.0
        for ci in empty:
            p = np.maximum(c - 1.0, 0.0)
            if p.sum() <= 0:
                break
            cj = int(rs.choice(len(c), p=p / p.sum()))
            cent[ci] = cent[cj] * (1.0 + eps * sign)
            cent[cj] = cent[cj] * (1.0 - eps * sign)
            c[ci] = c[cj] / 2.0
            c[cj] -= c[ci]
        return len(empty)

    def assign(self, x):
        """``(D, I)`` of every row's nearest centroid (faiss ``Kmeans.assign``)."""
        assert self.index is not None, "should train first"
        D, I = self.index.search(np.ascontiguousarray(x, dtype=np.float32), 1)
        return D.ravel(), I.ravel()
