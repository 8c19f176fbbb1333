"""Pins the CPU oracle (the faiss IndexFlatL2 restatement) -- CPU only.

The reference has no tests or golden vectors at this boundary and faiss cannot run here, so the
oracle is pinned against (i) hand-constructed known answers, (ii) the committed golden fixtures
(fp64 brute force on exactly-representable inputs) and (iii) two independent restatements
(C handlers vs pure numpy) agreeing with the fp64 ground truth."""
from pathlib import Path

import numpy as np
import pytest

from oracle import flatl2_oracle as orc

GOLDEN = sorted((Path(__file__).parent / "golden").glob("*.npz"))
FLT_MAX = np.float32(3.4028234663852886e38)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
@pytest.mark.parametrize("branch", ["seq", "blas", "blas_c", "numpy", None])
def test_oracle_matches_golden(path, branch):
    g = np.load(path)
    xb, xq, k = g["xb"], g["xq"], int(g["k"])
    if branch == "numpy":
        D, I = orc.knn_fp32_numpy(xq, xb, k)
    else:
        D, I = orc.knn_fp32(xq, xb, k, path=branch)
    if path.stem.startswith(("gauss", "pad")):   # gaussian inputs: fp64 truth, tolerance compare
        ok, msg = orc.compare_knn(D, I, g["D"], g["I"], rel_d=1e-5)
        assert ok, msg
    else:   # exactly representable inputs: bit-exact distances and canonical tie order
        np.testing.assert_array_equal(I, g["I"])
        np.testing.assert_array_equal(D, g["D"])


def test_known_answers_by_hand():
    xb = np.array([[0, 0], [3, 4], [1, 0], [0, 1], [3, 4]], dtype=np.float32)
    xq = np.array([[0, 0]], dtype=np.float32)
    D, I = orc.knn_fp32(xq, xb, 5)
    np.testing.assert_array_equal(I, [[0, 2, 3, 1, 4]])          # ties (1,1) and (25,25) -> ascending id
    np.testing.assert_array_equal(D, [[0, 1, 1, 25, 25]])
    D, I = orc.knn_fp32(xq, xb, 7)                                # k > ntotal
    assert I[0, 5:].tolist() == [-1, -1] and (D[0, 5:] == FLT_MAX).all()
    D, I = orc.knn_fp32(xq, xb, 1)                                # Top1 handler: strict minimum
    assert I.tolist() == [[0]] and D.tolist() == [[0.0]]
    D, I = orc.knn_fp32(xq, np.empty((0, 2), np.float32), 3)      # empty index
    assert (I == -1).all() and (D == FLT_MAX).all()


@pytest.mark.parametrize("nq,n,d,k", [(1, 50, 7, 4), (19, 400, 33, 10), (20, 400, 33, 10), (64, 3000, 128, 99),
                                      (64, 3000, 128, 100), (25, 5000, 64, 256), (300, 10, 16, 32)])
def test_fp32_restatements_agree_with_fp64(nq, n, d, k):
    rng = np.random.default_rng(nq * 1000 + n + d + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    D64, I64 = orc.knn_fp64(xq, xb, k)
    for path in (None, "seq", "blas", "blas_c"):
        D, I = orc.knn_fp32(xq, xb, k, path=path)
        ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64)
        assert ok, f"{path}: {msg}"
        real = I >= 0
        assert np.all(np.diff(D, axis=1)[real[:, 1:]] >= 0), "distances must ascend"
    Dn, In = orc.knn_fp32_numpy(xq, xb, k)
    ok, msg = orc.compare_knn(Dn, In, D64.astype(np.float32), I64)
    assert ok, msg


def test_blas_branch_switch_and_clamp():
    # nq >= 20 takes the expansion form (clamped at 0), nq < 20 the exact difference form
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((100, 64)).astype(np.float32) * 100
    xq = xb[:30].copy()
    D, I = orc.knn_fp32(xq, xb, 1)
    assert (D >= 0).all() and (I[:, 0] == np.arange(30)).all()
    D2, I2 = orc.knn_fp32(xq[:5], xb, 1)
    assert (D2 == 0).all() and (I2[:, 0] == np.arange(5)).all()     # difference form is exactly 0 on duplicates


def test_index_surface_matches_faiss_wrapper():
    rng = np.random.default_rng(9)
    xb = rng.standard_normal((200, 12))                              # float64, coerced like faiss
    ix = orc.IndexFlatL2(12)
    assert ix.ntotal == 0 and ix.is_trained and ix.d == 12
    ix.add(xb[:120]); ix.add(np.asfortranarray(xb[120:].astype(np.float32)))
    assert ix.ntotal == 200
    xq = rng.standard_normal((22, 12)).astype(np.float32)
    D, I = ix.search(xq, 6)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (22, 6)
    one = orc.IndexFlatL2(12); one.add(xb.astype(np.float32))
    D1, I1 = one.search(xq, 6)
    np.testing.assert_array_equal(I, I1); np.testing.assert_array_equal(D, D1)
    Dp, Ip = np.empty((22, 6), np.float32), np.empty((22, 6), np.int64)
    ix.search(xq, 6, D=Dp, I=Ip)
    np.testing.assert_array_equal(Ip, I)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((3, 11), np.float32))
    with pytest.raises(AssertionError):
        ix.search(xq, 0)
    ix.reset()
    assert ix.ntotal == 0
    D, I = ix.search(xq, 2)
    assert (I == -1).all()


def test_compare_knn_rejects_wrong_answers():
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((500, 16)).astype(np.float32)
    xq = rng.standard_normal((20, 16)).astype(np.float32)
    D, I = orc.knn_fp32(xq, xb, 8)
    assert orc.compare_knn(D, I, D, I)[0]
    Ibad = I.copy(); Ibad[3, 2] = (Ibad[3, 2] + 1) % 500
    assert not orc.compare_knn(D, Ibad, D, I)[0]
    Dbad = D.copy(); Dbad[4, 1] *= 1.001
    assert not orc.compare_knn(Dbad, I, D, I)[0]


def test_inner_product_oracle_matches_fp64_and_orders_ties_by_id():
    """knn_ip_fp32 (IndexFlatIP restatement, reference anyloc/utilities.py:446-457) against fp64 brute force."""
    rng = np.random.default_rng(21)
    xb = rng.standard_normal((500, 24)).astype(np.float32)
    xb[100:110] = xb[5:15]                      # exact ties
    for nq in (3, 40):                          # seq and blas branches
        xq = rng.standard_normal((nq, 24)).astype(np.float32)
        D, I = orc.knn_ip_fp32(xq, xb, 12)
        D64, I64 = orc.knn_ip_fp64(xq, xb, 12)
        ok, msg = orc.compare_knn(D, I, D64, I64, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
        assert ok, msg
        assert np.all(np.diff(D, axis=1) <= 0)
    D, I = orc.knn_ip_fp32(xq[:2], xb[:4], 6)   # k > n: (-FLT_MAX, -1) padding
    assert np.all(I[:, 4:] == -1) and np.all(D[:, 4:] == -orc.FLT_MAX)
    D, I = orc.knn_ip_fp32(xb[5:6], xb, 3)      # xb[5] == xb[100]: equal products, lower id first
    assert list(I[0, :2]) == [5, 100]


def test_oracle_reproduces_the_published_output_of_the_faiss_tutorial():
    """The one known-answer vector that real faiss has published for IndexFlatL2: its first tutorial
    (tutorial/python/1-Flat.py, faiss wiki 'Getting started').  The restatement must reproduce every id and the
    distances to the printed precision, on both of faiss's code paths (5 queries: nq < 20, difference form;
    10000 queries: sgemm blocks)."""
    from tests.helpers import faiss_tutorial_data
    xb, xq, pub = faiss_tutorial_data()
    D, I = orc.knn_fp32(xb[:5], xb, 4)                     # "sanity check" of the tutorial
    np.testing.assert_array_equal(I, np.array(pub["sanity_I"]))
    np.testing.assert_allclose(D, np.array(pub["sanity_D"]), rtol=2e-7, atol=1e-6)
    D, I = orc.knn_fp32(xq, xb, 4)                         # the actual search
    np.testing.assert_array_equal(I[:5], np.array(pub["search_I_first5"]))
    np.testing.assert_array_equal(I[-5:], np.array(pub["search_I_last5"]))
    D5, I5 = orc.knn_fp32(xb[:5], xb, 4, path="blas")      # same answers through the expansion form
    np.testing.assert_array_equal(I5, np.array(pub["sanity_I"]))


@pytest.mark.parametrize("nq,n,d,k", [(7, 900, 48, 5), (40, 2500, 64, 20), (64, 4000, 256, 100), (33, 700, 512, 50)])
def test_oracle_agrees_with_scikit_learn_brute_force(nq, n, d, k):
    """A third-party cross-check (not faiss, but an independently written exact k-NN that IS installed here and that the
    reference also depends on -- datasets_ws_kitti360.py:613): sklearn's brute-force NearestNeighbors with the squared
    euclidean metric returns the same neighbours as both of the oracle's code paths (ids equal except ties within 1e-5
    relative, distances within 1e-4 relative: BASELINE.json's tolerances)."""
    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(nq + n + d + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(xb.astype(np.float64))
    Ds, Is = nn.kneighbors(xq.astype(np.float64))
    for path in ("seq", "blas"):
        D, I = orc.knn_fp32(xq, xb, k, path=path)
        ok, msg = orc.compare_knn(D, I, Ds.astype(np.float32), Is.astype(np.int64), xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
        assert ok, f"{path}: {msg}"
    # inner product: the same library's dot products, descending
    Dip, Iip = orc.knn_ip_fp32(xq, xb, k)
    P = xq.astype(np.float64) @ xb.astype(np.float64).T
    order = np.argsort(-P, axis=1, kind="stable")[:, :k]
    top = np.take_along_axis(P, order, axis=1)
    assert np.allclose(Dip, top, rtol=1e-4, atol=1e-4)
    assert (Iip == order).mean() > 0.999
