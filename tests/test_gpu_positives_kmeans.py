"""SURVEY 8f N4 on the GPU: UTM radius positives against sklearn (the library the reference calls:
datasets/datasets_ws_kitti360.py:613-618, 740-745) and faiss.Kmeans' algorithm on the engine
(reference model/aggregation.py:170-171) -- needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def agp():
    import agplace_b200
    return agplace_b200


def sklearn_radius(db, q, r):
    from sklearn.neighbors import NearestNeighbors
    knn = NearestNeighbors(n_jobs=1)
    knn.fit(db)
    return knn.radius_neighbors(q, radius=r, return_distance=False)


@pytest.mark.parametrize("n,nq,side,r", [(5000, 700, 1000.0, 25.0), (20000, 3000, 3000.0, 10.0), (300, 40, 50.0, 25.0), (1, 3, 1.0, 5.0)])
def test_radius_positives_equal_sklearn(n, nq, side, r):
    rng = np.random.default_rng(n + nq)
    # UTM-sized coordinates (easting ~5e5, northing ~4e6): the differences need fp64
    db = rng.uniform(0, side, (n, 2)) + np.array([5.0e5, 4.0e6])
    q = db[rng.integers(0, n, nq)] + rng.normal(0, 5.0, (nq, 2))
    got = agp().radius_neighbors(db, q, r)
    want = sklearn_radius(db, q, r)
    assert got.dtype == object and len(got) == nq
    for i in range(nq):
        assert got[i].dtype == np.int64
        assert np.all(np.diff(got[i]) > 0), "ids ascending"
        np.testing.assert_array_equal(got[i], np.sort(want[i]))
    off, ids = agp().radius_neighbors(db, q, r, return_csr=True)
    assert off[0] == 0 and off[-1] == len(ids) == sum(len(w) for w in want)


def test_radius_is_inclusive_and_handles_empty_results():
    xs, ys = np.meshgrid(np.arange(20.0), np.arange(20.0))
    db = np.stack([xs.ravel(), ys.ravel()], 1)                       # integer lattice: squared distances are exact
    q = np.array([[10.0, 10.0], [0.0, 0.0], [100.0, 100.0]])
    got = agp().radius_neighbors(db, q, 5.0)                         # (3,4,5) triangles sit exactly on the boundary
    want = sklearn_radius(db, q, 5.0)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, np.sort(w))
    assert len(got[0]) == 81 and len(got[2]) == 0
    with pytest.raises(RuntimeError):
        agp().radius_neighbors(np.zeros((3, 9)), np.zeros((1, 9)), 1.0)     # dim > 8


def test_radius_positives_feed_the_recall_kernel():
    from agplace_b200 import synth
    from tests.helpers import reference_recall_loop
    ev = synth.make_eval_set(dict(n=3000, nq=400, d=64, k=20, seed=3, side=500.0), correlated=0.5)
    db_utm, q_utm = synth.utm_positions(3000, 400, 500.0, 3)
    pos = agp().radius_neighbors(db_utm, q_utm, 25.0)
    ix = agp().IndexFlatL2(64); ix.add(ev.database_features)
    _, I = ix.search(ev.queries_features, 20)
    hits = agp().recall_hits(I, pos, [1, 5, 10, 20])
    want = reference_recall_loop(I, sklearn_radius(db_utm, q_utm, 25.0), [1, 5, 10, 20])
    np.testing.assert_allclose(hits / 400 * 100, want)


def test_kmeans_recovers_separated_blobs_and_lowers_the_objective():
    rng = np.random.default_rng(11)
    k, d, per = 16, 32, 400
    centers = rng.standard_normal((k, d)).astype(np.float32) * 10.0
    x = (centers[:, None, :] + rng.standard_normal((k, per, d)).astype(np.float32)).reshape(-1, d)
    rng.shuffle(x)
    km = agp().Kmeans(d, k, niter=20, seed=7)
    final = km.train(x)
    assert km.centroids.shape == (k, d) and km.centroids.dtype == np.float32
    assert len(km.obj) == 20 and final == pytest.approx(float(km.obj[-1]))
    assert np.all(np.diff(km.obj[1:]) <= 1e-3 * km.obj[1]), "Lloyd iterations do not increase the objective"
    # every point's assigned centroid is its nearest one (fp64 check), and the objective matches
    D, I = km.assign(x)
    d2 = ((x[:, None, :].astype(np.float64) - km.centroids[None].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_array_equal(I, d2.argmin(1))
    np.testing.assert_allclose(D, d2.min(1), rtol=1e-4, atol=1e-4)
    # random-point initialisation can merge blobs (a local minimum of Lloyd's algorithm, in faiss too); started near
    # the truth, the iterations must land on the blob means
    km3 = agp().Kmeans(d, k, niter=10, max_points_per_centroid=1000)       # no subsampling: the objective covers all of x
    km3.train(x, init_centroids=centers + 0.5 * rng.standard_normal((k, d)).astype(np.float32))
    near = ((centers[:, None, :] - km3.centroids[None]) ** 2).sum(-1).min(1)
    assert np.all(near < 0.5)
    assert km3.obj[-1] / len(x) <= km.obj[-1] / min(len(x), k * 256) * 1.0001      # per point: km trained on a subsample
    assert km3.obj[-1] == pytest.approx(len(x) * d, rel=0.05)          # unit-variance blobs: E|x - mean|^2 = d per point
    # subsampling path (n > k * max_points_per_centroid) and the faiss-style surface
    km2 = agp().Kmeans(d, 4, niter=5, max_points_per_centroid=64)
    km2.train(x)
    assert km2.centroids.shape == (4, d) and km2.index.ntotal == 4
    with pytest.raises(RuntimeError):
        agp().Kmeans(d, 50).train(x[:10])
