"""IndexFlatIP on the engine (reference anyloc/utilities.py:396-475) against the CPU oracle -- needs a B200.

Tolerances: BASELINE.json's 1e-4 relative on the products plus an absolute floor of a few fp32 ulps of |q||y|
(an fp32 inner product of d terms is not more accurate than that in faiss either); indices identical except ties
within 1e-5 relative (+ the same floor)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatl2_oracle as orc

NEG_MAX = np.float32(-3.4028234663852886e38)


def agp():
    import agplace_b200
    return agplace_b200


def check(xb, xq, k, precision="auto", chunks=1):
    ix = agp().IndexFlatIP(xb.shape[1], precision=precision)
    assert ix.metric_type == agp().METRIC_INNER_PRODUCT == 0
    for part in np.array_split(xb, chunks):
        ix.add(part)
    D, I = ix.search(xq, k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (len(xq), k)
    Dr, Ir = orc.knn_ip_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=16 * 2.0 ** -24)
    assert ok, f"{precision} nq={len(xq)} n={len(xb)} d={xb.shape[1]} k={k}: {msg}"
    # the products must be non-increasing and the ids must be rows whose exact product matches
    valid = I >= 0
    assert np.all(np.diff(D, axis=1)[valid[:, 1:]] <= 0)
    D64, I64 = orc.knn_ip_fp64(xq, xb, k)
    got = np.einsum("qd,qkd->qk", xq.astype(np.float64), xb[np.clip(I, 0, len(xb) - 1)].astype(np.float64))
    scale = np.linalg.norm(xq, axis=1)[:, None] * np.linalg.norm(xb, axis=1).max()
    assert np.all(np.abs(got - D)[valid] <= (1e-4 * np.abs(got) + 32 * 2.0 ** -24 * scale)[valid])
    # recall of the true top-k set (fp64), allowing boundary ties
    kth = D64[:, -1:][:, 0]
    assert np.all((got >= kth[:, None] - 1e-5 * np.abs(kth[:, None]) - 32 * 2.0 ** -24 * scale)[valid])
    return D, I


SHAPES = [(1, 1, 1, 1), (1, 1000, 256, 10), (3, 5, 3, 8), (19, 999, 64, 20), (20, 999, 64, 20), (129, 257, 33, 33),
          (300, 5000, 255, 50), (257, 4097, 512, 100), (64, 1500, 100, 256), (50, 3000, 16, 257), (1000, 130, 48, 20)]


@pytest.mark.parametrize("precision", ["auto", "fp16_screen", "fp32_simt", "exact_diff"])
@pytest.mark.parametrize("nq,n,d,k", SHAPES)
def test_random_shapes_match_oracle(nq, n, d, k, precision):
    rng = np.random.default_rng(nq * 11 + n * 5 + d + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    check(xb, xq, k, precision, chunks=2 if n > 10 else 1)


@pytest.mark.parametrize("sigma", [3e-2, 3e-3])
def test_unit_norm_clustered_cosine(sigma):
    """The anyloc call pattern: F.normalize'd descriptors, queries next to database rows (near-ties everywhere)."""
    rng = np.random.default_rng(5)
    n, nq, d, k = 20000, 700, 256, 20
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = xb[rng.integers(0, n, nq)] + sigma * rng.standard_normal((nq, d)).astype(np.float32) / np.sqrt(d)
    xq = (xq / np.linalg.norm(xq, axis=1, keepdims=True)).astype(np.float32)
    check(xb, xq, k)


def test_mixed_signs_scales_and_negative_best_products():
    rng = np.random.default_rng(6)
    xb = (rng.standard_normal((3000, 64)) * rng.uniform(0.1, 30.0, (3000, 1))).astype(np.float32)
    xq = (rng.standard_normal((200, 64)) * rng.uniform(0.01, 5.0, (200, 1))).astype(np.float32)
    check(xb, xq, 25)
    # every product negative: the "largest" are the least negative ones
    xb2 = np.abs(xb)
    xq2 = -np.abs(xq)
    D, _ = check(xb2, xq2, 10)
    assert np.all(D < 0)


def test_padding_empty_reset_and_duplicates():
    rng = np.random.default_rng(7)
    xb = rng.standard_normal((30, 16)).astype(np.float32)
    xq = rng.standard_normal((25, 16)).astype(np.float32)
    ix = agp().IndexFlatIP(16)
    D, I = ix.search(xq, 4)                                   # empty index: all padding
    assert np.all(I == -1) and np.all(D == NEG_MAX)
    ix.add(xb)
    D, I = ix.search(xq, 40)                                  # k > ntotal
    assert np.all(I[:, 30:] == -1) and np.all(D[:, 30:] == NEG_MAX) and np.all(I[:, :30] >= 0)
    D1, I1 = ix.search(xq[:2], 40)                            # small-batch path pads the same way
    np.testing.assert_array_equal(I1[:, 30:], -1)
    ix.reset()
    assert ix.ntotal == 0
    # heavy duplication: exact ties resolve by id, and the certified band overflows into the exact fallback
    base = rng.standard_normal((3, 32)).astype(np.float32)
    xb = base[rng.integers(0, 3, 4000)]
    xq = rng.standard_normal((64, 32)).astype(np.float32)
    ix = agp().IndexFlatIP(32); ix.add(xb)
    D, I = ix.search(xq, 20)
    Dr, Ir = orc.knn_ip_fp32(xq, xb, 20)
    np.testing.assert_allclose(D, Dr, rtol=1e-4, atol=1e-5)
    for q in range(len(xq)):                                  # among equal products the lowest ids win, ascending
        for v in np.unique(D[q]):
            ids = I[q][D[q] == v]
            assert np.all(np.diff(ids) > 0)


def test_torch_cuda_tensors_and_the_anyloc_recall_helper():
    import torch
    from agplace_b200.recall import get_top_k_recall
    rng = np.random.default_rng(8)
    n, nq, d = 5000, 300, 128
    db = rng.standard_normal((n, d)).astype(np.float32)
    qu = db[rng.integers(0, n, nq)] + 0.3 * rng.standard_normal((nq, d)).astype(np.float32)
    dbn = db / np.linalg.norm(db, axis=1, keepdims=True)
    qun = qu / np.linalg.norm(qu, axis=1, keepdims=True)
    Dr, Ir = orc.knn_ip_fp32(qun.astype(np.float32), dbn.astype(np.float32), 10)
    gt = np.empty(nq, dtype=object)
    for i in range(nq):
        gt[i] = rng.choice(n, size=5, replace=False) if i % 3 else Ir[i, :2].copy()
    top_k = [1, 5, 10]
    want = {k: float(np.mean([np.any(np.isin(Ir[i, :k], gt[i])) for i in range(nq)])) for k in top_k}
    for dev in ("cpu", "cuda"):
        dist, ind, rec = get_top_k_recall(top_k, torch.from_numpy(db).to(dev), torch.from_numpy(qu).to(dev), gt, method="cosine")
        assert str(ind.device).startswith(dev)
        ok, msg = orc.compare_knn(dist.cpu().numpy(), ind.cpu().numpy(), Dr, Ir, xq=qun, xb=dbn, abs_floor_eps=16 * 2.0 ** -24)
        assert ok, msg
        assert rec == pytest.approx(want)
    dist, ind, rec = get_top_k_recall(top_k, torch.from_numpy(db), torch.from_numpy(qu), gt, method="l2")
    Dl, Il = orc.knn_fp32(qun.astype(np.float32), dbn.astype(np.float32), 10)
    ok, msg = orc.compare_knn(dist.numpy(), ind.numpy(), Dl, Il, xq=qun, xb=dbn, abs_floor_eps=16 * 2.0 ** -24)
    assert ok, msg


def test_faiss_style_helpers_and_l2_only_entry_points():
    a = agp()
    ix = a.IndexFlat(8, a.METRIC_INNER_PRODUCT)
    assert isinstance(ix, a.IndexFlatIP)
    assert a.index_cpu_to_gpu(a.StandardGpuResources(), 0, ix) is ix
    ix.add(np.eye(8, dtype=np.float32))
    D, I = ix.search(np.eye(8, dtype=np.float32)[:3] * 2.0, 1)
    np.testing.assert_array_equal(I[:, 0], [0, 1, 2])
    np.testing.assert_array_equal(D[:, 0], [2.0, 2.0, 2.0])
    with pytest.raises(RuntimeError):
        ix.search_masked(np.zeros((1, 8), np.float32), 1, [np.array([0])])
    with pytest.raises(RuntimeError):
        ix.search_subset(np.zeros((1, 8), np.float32), 1, [np.array([0])])
    with pytest.raises(RuntimeError):
        a.IndexFlatIP(8, precision="3xtf32")


def test_virtual_shards_merge_equals_single_inner_product_index():
    """agp_merge_topk_metric over per-shard IndexFlatIP results (what ShardedIndexFlatIP exchanges) = one index."""
    import ctypes
    import torch
    from agplace_b200 import _lib
    from agplace_b200.sharded import shard_bounds
    rng = np.random.default_rng(9)
    n, nq, d, k, G = 3001, 130, 24, 40, 3
    xb = rng.integers(-4, 5, size=(n, d)).astype(np.float32)          # lattice: exact ties across shards
    xq = rng.integers(-4, 5, size=(nq, d)).astype(np.float32)
    single = agp().IndexFlatIP(d); single.add(xb)
    Ds, Is = single.search(xq, k)
    xq_dev = torch.from_numpy(xq).cuda()
    Dl = torch.empty((G, nq, k), dtype=torch.float32, device="cuda")
    Il = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
    keep = []
    for g, (a, b) in enumerate(shard_bounds(n, G)):
        sh = agp().IndexFlatIP(d); sh.add(xb[a:b]); sh.set_id_base(a)
        sh.search(xq_dev, k, D=Dl[g], I=Il[g])
        keep.append(sh)
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.agp_merge_topk_metric(0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), nq, k, G,
                                         ctypes.c_void_p(Dl.data_ptr()), nq * k, ctypes.c_void_p(Il.data_ptr()), nq * k, n, 0,
                                         ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr())), "agp_merge_topk_metric")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(D.cpu().numpy(), Ds)
    np.testing.assert_array_equal(I.cpu().numpy(), Is)
    Dr, Ir = orc.knn_ip_fp32(xq, xb, k)
    np.testing.assert_array_equal(Is, Ir)
    np.testing.assert_array_equal(Ds, Dr)
