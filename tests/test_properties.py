"""Property tests (hypothesis; SURVEY.md section 4's plan): random shapes and value regimes.

CPU half (runs everywhere): both code paths of the oracle -- the nq < 20 difference form and the sgemm expansion
form -- against fp64 truth and against each other, plus the structural invariants of a faiss result (sorted,
unique ids, (FLT_MAX, -1) padding exactly when k > ntotal, idempotence, permutation equivariance).
GPU half (``-m gpu``): the engine against the oracle on the same draws."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import flatl2_oracle as orc

FLT_MAX = np.float32(3.4028234663852886e38)
DIMS = [1, 3, 8, 64, 255, 256, 512, 513]


@st.composite
def problems(draw, max_n=5000, max_nq=300):
    d = draw(st.sampled_from(DIMS))
    n = draw(st.one_of(st.integers(1, 40), st.integers(1, max_n)))
    nq = draw(st.one_of(st.integers(1, 25), st.integers(1, max_nq)))
    k = draw(st.one_of(st.integers(1, 12), st.integers(1, 300), st.just(n), st.just(n + 7)))
    k = min(k, 512)
    regime = draw(st.sampled_from(["gauss", "unit", "lattice", "clustered", "scaled"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    if regime == "lattice":           # exactly representable values: exact ties, bit-exact distances on every path
        xb = rng.integers(-4, 5, size=(n, d)).astype(np.float32)
        xq = rng.integers(-4, 5, size=(nq, d)).astype(np.float32)
    else:
        xb = rng.standard_normal((n, d)).astype(np.float32)
        xq = rng.standard_normal((nq, d)).astype(np.float32)
        if regime == "unit":
            xb /= np.linalg.norm(xb, axis=1, keepdims=True) + 1e-12
            xq /= np.linalg.norm(xq, axis=1, keepdims=True) + 1e-12
        elif regime == "clustered":   # queries are noisy copies of rows
            xq = (xb[rng.integers(0, n, nq)] + np.float32(1e-2) * rng.standard_normal((nq, d)).astype(np.float32)).astype(np.float32)
        elif regime == "scaled":
            s = np.float32(10.0 ** rng.uniform(-3, 3))
            xb *= s
            xq *= s
    return xb, xq, k, regime


def check_structure(D, I, n, k):
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == I.shape
    real = I >= 0
    m = min(n, k)
    assert real[:, :m].all() and not real[:, m:].any(), "padding exactly beyond min(k, ntotal)"
    assert (D[~real] == FLT_MAX).all()
    assert ((I < n) & (I >= -1)).all()
    assert (np.diff(D[:, :m], axis=1) >= 0).all(), "distances ascending"
    for row in I[:, :m]:
        assert len(np.unique(row)) == m, "ids unique"
    # exact ties are ordered by id
    same = np.diff(D[:, :m], axis=1) == 0
    assert (np.diff(I[:, :m], axis=1)[same] > 0).all()


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(problems(max_n=1500, max_nq=120))
def test_oracle_paths_agree_with_fp64_truth(p):
    xb, xq, k, regime = p
    n = len(xb)
    D64, I64 = orc.knn_fp64(xq, xb, k)
    for path in ("seq", "blas"):
        D, I = orc.knn_fp32(xq, xb, k, path=path)
        check_structure(D, I, n, k)
        floor = 0.0 if regime == "lattice" else 8 * 2.0 ** -24
        ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64, xq=xq, xb=xb, abs_floor_eps=floor)
        assert ok, f"{path} {regime}: {msg}"
        if regime == "lattice":
            np.testing.assert_array_equal(I, I64)
            np.testing.assert_array_equal(D[I >= 0], D64.astype(np.float32)[I64 >= 0])
    # idempotence and query-permutation equivariance of the restatement
    D1, I1 = orc.knn_fp32(xq, xb, k)
    D2, I2 = orc.knn_fp32(xq, xb, k)
    np.testing.assert_array_equal(I1, I2)
    perm = np.random.default_rng(1).permutation(len(xq))
    if len(xq) < 20 or regime == "lattice":          # (the sgemm path's rounding depends on the block a query lands in)
        Dp, Ip = orc.knn_fp32(xq[perm], xb, k, path="seq")
        Ds, Is = orc.knn_fp32(xq, xb, k, path="seq")
        np.testing.assert_array_equal(Ip, Is[perm])
        np.testing.assert_array_equal(Dp, Ds[perm])


@pytest.mark.gpu
@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(problems())
def test_engine_agrees_with_oracle_on_random_problems(p):
    import agplace_b200 as agp
    xb, xq, k, regime = p
    n, d = xb.shape
    ix = agp.IndexFlatL2(d)
    cut = n // 3
    if cut:
        ix.add(xb[:cut])
    ix.add(xb[cut:])
    D, I = ix.search(xq, k)
    check_structure(D, I, n, k)
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    if regime == "lattice":
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D, Dr)
    else:
        ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=32 * 2.0 ** -24)
        assert ok, f"{regime} nq={len(xq)} n={n} d={d} k={k}: {msg}"
        D64, I64 = orc.knn_fp64(xq, xb, k)
        ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64, xq=xq, xb=xb, abs_floor_eps=32 * 2.0 ** -24)
        assert ok, f"vs fp64 {regime} nq={len(xq)} n={n} d={d} k={k}: {msg}"
    # idempotence: the same call returns the same bits (no dependence on scheduling / atomics order)
    D2, I2 = ix.search(xq, k)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)
