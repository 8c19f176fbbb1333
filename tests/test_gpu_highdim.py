"""Parity of the d_pad > 512 path (BASELINE.json configs[4]: 4096-d aggregation descriptors; reference producer
model/aggregation.py:148-173 NetVLAD, clusters x dim) and a direct test of the certification constant.

For d_pad > 512 the screen kernel cannot keep the query tile in shared memory: queries stream through the operand
ring together with the database chunks (knn_screen.cuh: q_resident = 0, 2 x 16 KB stages) -- a different shared
memory layout, ring depth and barrier protocol from the d <= 512 path, so it gets its own parity matrix here:
every mode against the fp32 restatement of faiss AND fp64 truth, ragged d, near-duplicates, large k, a reduced
cfg5, and a cfg5-shaped full-width check against planted neighbours.

``test_screened_distance_stays_inside_the_certified_band`` reads the screened distance of EVERY (query, row) pair
back from the tensor-core kernel (agp_index_screen_probe) and asserts |dis~ - dis_fp64| <= screen_band(q): the
selection is a superset of the true top-k iff that holds, and it is the only place the tensor core's fp32
accumulation behaviour (truncation per K = 16 step, launch.h:screen_band c_acc) enters the argument."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatl2_oracle as orc

FLT_MAX = np.float32(3.4028234663852886e38)


def agp():
    import agplace_b200
    return agplace_b200


def search(xb, xq, k, precision="auto", chunks=1):
    ix = agp().IndexFlatL2(xb.shape[1], precision=precision)
    for part in np.array_split(xb, chunks):
        ix.add(part)
    out = ix.search(xq, k)
    return out, ix


HIGH_D = [  # nq, n, d, k
    (300, 3000, 513, 10), (257, 2500, 576, 50), (300, 4000, 1000, 100), (513, 3000, 1024, 20), (40, 1500, 1024, 256),
    (300, 2000, 2048, 100), (300, 2600, 4096, 10), (64, 1300, 4096, 256), (20, 700, 4097, 5), (600, 257, 640, 300),
]


@pytest.mark.parametrize("precision", ["auto", "fp16_screen", "fp32_simt"])
@pytest.mark.parametrize("nq,n,d,k", HIGH_D)
def test_high_dimensional_shapes_match_oracle(nq, n, d, k, precision):
    rng = np.random.default_rng(nq * 7 + n * 3 + d + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    (D, I), ix = search(xb, xq, k, precision, chunks=2)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (nq, k) and I.shape == (nq, k)
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    # the expansion form the restatement (and faiss) uses for nq >= 20 carries ~1e-7 (|q|^2 + |x|^2) of cancellation
    # error; at d = 4096 that is a few 1e-4 absolute on distances of ~8000, inside the 1e-4 relative tolerance
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    D64, I64 = orc.knn_fp64(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, "vs fp64 truth: " + msg
    real = I >= 0
    assert np.all(np.diff(D, axis=1)[real[:, 1:]] >= 0)
    if k > n:
        assert (I[:, n:] == -1).all() and (D[:, n:] == FLT_MAX).all()
    if precision == "fp16_screen":
        assert ix.get_stats() == (nq, 0), "well-conditioned data must be answered by the tensor-core screen itself"


@pytest.mark.parametrize("d", [1024, 4096])
@pytest.mark.parametrize("sigma", [3e-2, 3e-3])
@pytest.mark.parametrize("k", [10, 100, 256])
def test_high_dimensional_near_duplicates(d, sigma, k):
    """Unit-norm descriptors, every query a noisy copy of a database row (the cancellation regime)."""
    from agplace_b200 import synth
    xb = synth.descriptors(3000, d, 5, "db")
    xq, src = synth.clustered_queries(xb, 200, sigma, 6)
    (D, I), ix = search(xb, xq, k, "fp16_screen")
    assert (I[:, 0] == src).all(), "the perturbed source row must be the nearest neighbour"
    D64, I64 = orc.knn_fp64(xq, xb, k)
    # distances come from the fp32 difference form: accurate relative to the distance itself
    ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64, rel_d=2e-5)
    assert ok, msg
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=32 * 2.0 ** -24)
    assert ok, msg


def test_reduced_cfg5_matches_oracle():
    """cfg5 at 1/20 of its rows: 50k x 4096 database, 2k queries, top-100 -- whole waves of pair tiles (8 pair
    tiles, split remainder) on the streamed-query path, against the restatement (all queries) and fp64 (a sample)."""
    from agplace_b200 import synth
    n, nq, d, k = 50_000, 2_000, 4096, 100
    xb = synth.descriptors(n, d, 4, "db")
    xq = synth.descriptors(nq, d, 11, "q")
    (D, I), ix = search(xb, xq, k, "auto", chunks=3)
    assert ix.get_stats() == (nq, 0)
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    sample = np.arange(0, nq, 31)
    D64, I64 = orc.knn_fp64(xq[sample], xb, k)
    ok, msg = orc.compare_knn(D[sample], I[sample], D64.astype(np.float32), I64, rel_d=2e-5)
    assert ok, "vs fp64 truth: " + msg


def test_cfg5_width_planted_neighbours_on_a_large_database():
    """A database too large for the CPU oracle to scan (400k x 4096 = 6.5 GB, generated on the device by the
    counter-based generator the bench uses): queries are noisy copies of known rows plus k - 1 planted near copies,
    so the true top-k is known by construction; everything else is ~sqrt(2) away.  Returned distances are
    re-derived on the CPU from regenerated rows."""
    import torch
    from agplace_b200 import synth
    n, nq, d, k = 400_000, 512, 4096, 8
    ix = agp().IndexFlatL2(d)
    ix.reserve(n)
    rng = np.random.default_rng(3)
    src = np.sort(rng.choice(n // 2, size=nq, replace=False)).astype(np.int64)
    base = synth.counter_rows(src, d, seed=17)                                  # the rows the queries are copies of
    planted_ids = n - nq * (k - 1) + np.arange(nq * (k - 1))                    # tail rows get overwritten with near copies
    planted = (np.repeat(base, k - 1, axis=0) + np.float32(2e-2 / np.sqrt(d)) * rng.standard_normal((nq * (k - 1), d)).astype(np.float32)).astype(np.float32)
    step = 65536
    for a in range(0, n, step):
        b = min(n, a + step)
        x = synth.counter_rows_device(a, b, d, seed=17, device=torch.device("cuda", ix.device))
        lo = np.searchsorted(planted_ids, a)
        hi = np.searchsorted(planted_ids, b)
        if hi > lo:
            x[planted_ids[lo:hi] - a] = torch.from_numpy(planted[lo:hi]).to(x.device)
        ix.add(x)
    xq = (base + np.float32(1e-2 / np.sqrt(d)) * rng.standard_normal((nq, d)).astype(np.float32)).astype(np.float32)
    D, I = ix.search(xq, k)
    assert ix.get_stats() == (nq, 0)
    want = np.concatenate([src[:, None], planted_ids.reshape(nq, k - 1)], axis=1)
    assert (I[:, 0] == src).all()
    np.testing.assert_array_equal(np.sort(I, axis=1), np.sort(want, axis=1))
    rows = np.concatenate([base[:, None, :], planted.reshape(nq, k - 1, d)], axis=1)          # [nq, k, d] in `want` order
    d64 = ((xq[:, None, :].astype(np.float64) - rows.astype(np.float64)) ** 2).sum(-1)
    order = np.argsort(d64, axis=1, kind="stable")
    np.testing.assert_array_equal(I, np.take_along_axis(want, order, 1))
    np.testing.assert_allclose(D, np.take_along_axis(d64, order, 1), rtol=2e-5)


# ------------------------------------------------------------------------------------------ certification constant
def _adversarial_sets(d, rng):
    """(name, xb, xq) triples that stress one term of screen_band each."""
    n, nq = 1536, 256
    out = []
    g = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    out.append(("gaussian", g, q))
    out.append(("unit_norm", g / np.linalg.norm(g, axis=1, keepdims=True), q / np.linalg.norm(q, axis=1, keepdims=True)))
    # all-positive entries: every product has the same sign, the accumulator grows monotonically and each truncated
    # add loses up to one ulp of the running sum in the same direction -- the worst case for c_acc
    out.append(("all_positive", np.abs(g) + 0.25, np.abs(q) + 0.25))
    # exactly fp16-representable operands (multiples of 2^-10 below 1): the operand-rounding terms vanish, so the
    # band is the accumulation term alone
    lat_b = (rng.integers(1, 1024, size=(n, d)) / 1024.0).astype(np.float32)
    lat_q = (rng.integers(1, 1024, size=(nq, d)) / 1024.0).astype(np.float32)
    out.append(("fp16_lattice_positive", lat_b, lat_q))
    out.append(("fp16_lattice_signed", lat_b * rng.choice([-1.0, 1.0], size=(n, d)).astype(np.float32),
                lat_q * rng.choice([-1.0, 1.0], size=(nq, d)).astype(np.float32)))
    # near duplicates: true distances ~1e-5 of the norms
    nd = g[rng.integers(0, n, nq)] * (1 + 1e-3 * rng.standard_normal((nq, d)).astype(np.float32))
    out.append(("near_duplicates", g, nd))
    # one heavy coordinate per row next to many small ones (worst case for the per-row fp16 scale)
    spiky = g * 1e-3
    spiky[np.arange(n), rng.integers(0, d, n)] = 30.0
    out.append(("spiky", spiky, q))
    return out


@pytest.mark.parametrize("d", [64, 512, 1000, 4096])
def test_screened_distance_stays_inside_the_certified_band(d):
    rng = np.random.default_rng(100 + d)
    worst = {}
    for name, xb, xq in _adversarial_sets(d, rng):
        ix = agp().IndexFlatL2(d, precision="fp16_screen")
        ix.add(xb)
        dis, band = ix.screen_probe(xq)
        assert np.isfinite(band).all() and np.isfinite(dis).all(), name
        xq64, xb64 = xq.astype(np.float64), xb.astype(np.float64)
        true = (xq64 * xq64).sum(1)[:, None] + (xb64 * xb64).sum(1)[None, :] - 2.0 * (xq64 @ xb64.T)
        err = np.abs(dis.astype(np.float64) - true)
        ratio = (err / band[:, None].astype(np.float64)).max()
        worst[name] = float(ratio)
        assert ratio <= 1.0, f"d={d} {name}: |dis~ - dis64| reaches {ratio:.3f} of the certified band"
        # the band must also be useful, not just safe: a few 1e-3 of the norms at most
        scale = (xq64 * xq64).sum(1).max() + (xb64 * xb64).sum(1).max()
        assert band.max() <= 2e-2 * scale, f"d={d} {name}: band {band.max():.3g} vs norms {scale:.3g}"
    print(f"d={d}: max |dis~ - dis64| / band per set: {worst}")


def test_accumulation_error_alone_is_inside_its_term():
    """fp16-lattice operands make every other term of the band zero or negligible: what is left of |dis~ - dis64| is
    the tensor core's fp32 accumulation (plus the fp32 norm / epilogue terms).  Assert it against the c_acc term alone,
    (K/16 + 17) 2^-22 per unit of |q||y| + |y|^2, at the longest K the engine is specified for."""
    rng = np.random.default_rng(7)
    for d in (512, 4096):
        n, nq = 1024, 256
        xb = (rng.integers(1, 1024, size=(n, d)) / 1024.0).astype(np.float32)
        xq = (rng.integers(1, 1024, size=(nq, d)) / 1024.0).astype(np.float32)
        ix = agp().IndexFlatL2(d, precision="fp16_screen")
        ix.add(xb)
        dis, band = ix.screen_probe(xq)
        xq64, xb64 = xq.astype(np.float64), xb.astype(np.float64)
        qn, yn = (xq64 * xq64).sum(1), (xb64 * xb64).sum(1)
        true = qn[:, None] + yn[None, :] - 2.0 * (xq64 @ xb64.T)
        d_pad = (d + 63) // 64 * 64
        c_acc = (d_pad / 16 + 17) * 2.0 ** -22
        term = 2.0 * c_acc * (np.sqrt(qn)[:, None] * np.sqrt(yn.max()) + yn.max()) + 2.0 ** -20 * (qn[:, None] + yn.max())
        ratio = (np.abs(dis.astype(np.float64) - true) / term).max()
        print(f"d={d}: accumulation error reaches {ratio:.3f} of its term")
        assert ratio <= 1.0
