"""Parity tests proper: the CUDA engine (through the C ABI) against the CPU oracle -- needs a B200.

Tolerances are BASELINE.json's: neighbour indices identical except ties whose fp32 distances
differ by < 1e-5 relative; distances within 1e-4 relative (plus, in the near-duplicate regime only,
an absolute floor of a few fp32 ulps of |q|^2 + |x|^2 -- the expansion form that faiss itself uses
for nq >= 20 is not more accurate than that, SURVEY.md finding 3)."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatl2_oracle as orc
from tests.helpers import make_mining_problem, reference_recall_loop

GOLDEN = sorted((Path(__file__).parent / "golden").glob("*.npz"))
MODES = ["auto", "fp16_screen", "3xtf32", "3xfp16", "fp32_simt", "exact_diff"]
FLT_MAX = np.float32(3.4028234663852886e38)


def agp():
    import agplace_b200
    return agplace_b200


def search(xb, xq, k, precision="auto", chunks=1):
    ix = agp().IndexFlatL2(xb.shape[1], precision=precision)
    for part in np.array_split(xb, chunks):
        ix.add(part)
    assert ix.ntotal == len(xb)
    return ix.search(xq, k)


@pytest.mark.parametrize("precision", MODES)
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_golden_vectors(path, precision):
    g = np.load(path)
    D, I = search(g["xb"], g["xq"], int(g["k"]), precision)
    if path.stem.startswith(("gauss", "pad")):
        ok, msg = orc.compare_knn(D, I, g["D"], g["I"], rel_d=1e-5)
        assert ok, msg
    else:   # exactly representable inputs: bit-exact distances, canonical tie order, padding
        np.testing.assert_array_equal(I, g["I"])
        np.testing.assert_array_equal(D, g["D"])


SHAPES = [  # nq, n, d, k
    (1, 1, 1, 1), (1, 1000, 256, 10), (1, 37, 256, 1), (3, 5, 3, 8), (19, 999, 64, 20), (20, 999, 64, 20),
    (127, 255, 8, 5), (128, 256, 32, 32), (129, 257, 33, 33), (300, 5000, 255, 50), (257, 4097, 512, 100),
    (64, 1500, 100, 256), (50, 3000, 16, 257), (40, 700, 24, 512), (1000, 130, 48, 20), (21, 70000, 32, 10),
    (300, 20000, 64, 400), (257, 9000, 128, 512),      # 256 < k <= 512: the 1024-slot instantiation of the screen kernel
]


@pytest.mark.parametrize("precision", ["auto", "fp16_screen", "3xtf32", "3xfp16", "fp32_simt"])
@pytest.mark.parametrize("nq,n,d,k", SHAPES)
def test_random_shapes_match_oracle(nq, n, d, k, precision):
    rng = np.random.default_rng(nq * 7 + n * 3 + d + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    D, I = search(xb, xq, k, precision)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (nq, k) and I.shape == (nq, k)
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir)
    assert ok, msg
    D64, I64 = orc.knn_fp64(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64)
    assert ok, "vs fp64 truth: " + msg
    real = I >= 0
    assert np.all(np.diff(D, axis=1)[real[:, 1:]] >= 0)
    if k > n:
        assert (I[:, n:] == -1).all() and (D[:, n:] == FLT_MAX).all()


def test_unit_norm_descriptors_cfg1_shape_and_recall_identical():
    from agplace_b200 import recall, synth
    ev = synth.make_eval_set("cfg1", correlated=0.16)        # R@1 ~ 21 %, R@20 ~ 57 % on the CPU oracle
    args = SimpleNamespace(features_dim=256, recall_values=[1, 5, 10, 20])
    r_gpu, s_gpu = recall.compute_recall(args, ev.queries_features, ev.database_features, ev)
    r_dev, _ = recall.compute_recall(args, ev.queries_features, ev.database_features, ev, on_device_recall=True)
    r_cpu, s_cpu = recall.compute_recall(args, ev.queries_features, ev.database_features, ev, index_cls=orc.IndexFlatL2)
    np.testing.assert_array_equal(r_gpu, r_cpu)
    np.testing.assert_array_equal(r_dev, r_cpu)
    assert s_gpu == s_cpu and 5 < r_cpu[0] < 100
    D, I = search(ev.database_features, ev.queries_features, 20)
    Dr, Ir = orc.knn_fp32(ev.queries_features, ev.database_features, 20)
    ok, msg = orc.compare_knn(D, I, Dr, Ir)
    assert ok, msg


@pytest.mark.parametrize("precision", ["auto", "fp32_simt"])
@pytest.mark.parametrize("sigma", [3e-2, 3e-3])
def test_near_duplicate_regime(sigma, precision):
    from agplace_b200 import synth
    xb = synth.descriptors(4000, 256, 5, "db")
    xq, src = synth.clustered_queries(xb, 200, sigma, 6)
    D, I = search(xb, xq, 10, precision)
    assert (I[:, 0] == src).all(), "the perturbed source row must be the nearest neighbour"
    Dr, Ir = orc.knn_fp32(xq, xb, 10)
    # expansion form on both sides: absolute floor of 8 ulp of (|q|^2 + |x|^2) for fp32 FMA; the
    # tensor-core path gets 32 ulp because tcgen05 accumulates with round-toward-zero (measured
    # ~1e-6 relative to the norms over a 192-MMA chain, DESIGN.md "Numerics")
    ulps = 8 if precision == "fp32_simt" else 32
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=ulps * 2.0 ** -24)
    assert ok, msg
    # small batches take the exact difference form (like faiss): tight agreement with fp64 truth
    D1, I1 = search(xb, xq[:7], 10, "auto")
    D64, I64 = orc.knn_fp64(xq[:7], xb, 10)
    ok, msg = orc.compare_knn(D1, I1, D64.astype(np.float32), I64, rel_d=2e-6)
    assert ok, msg


def test_huge_norms_zeros_and_exact_ties_across_tiles():
    rng = np.random.default_rng(8)
    xb = rng.standard_normal((3000, 64)).astype(np.float32) * 100.0       # |x|^2 ~ 6e5
    xb[100] = 0; xb[2900] = 0
    xq = rng.standard_normal((150, 64)).astype(np.float32) * 100.0
    xq[0] = 0
    D, I = search(xb, xq, 30)
    Dr, Ir = orc.knn_fp32(xq, xb, 30)
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    assert I[0, 0] == 100 and I[0, 1] == 2900 and D[0, 0] == 0 and D[0, 1] == 0   # tie -> lower id first
    # identical rows spread over many 256-row tiles and splits: ties resolve to ascending ids everywhere
    base = rng.integers(-3, 4, size=(1, 32)).astype(np.float32)
    xb2 = np.repeat(base, 5000, axis=0)
    xq2 = rng.integers(-3, 4, size=(40, 32)).astype(np.float32)
    for precision in ("auto", "3xtf32", "fp32_simt"):
        D2, I2 = search(xb2, xq2, 64, precision)
        np.testing.assert_array_equal(I2, np.tile(np.arange(64), (40, 1)))
        assert (D2 == D2[:, :1]).all()


def test_screen_certified_band_and_fallback():
    """The single-pass fp16 screen must return the fp64-true neighbour set whatever the data looks like: a
    band that does not fit (mass duplicates) or rows the fp16 plane cannot hold (mixed magnitudes, tiny or
    huge scales) are answered by the exact fallback, visible in get_stats()."""
    rng = np.random.default_rng(21)
    # (a) well-behaved descriptors: no fallback at all
    xb = rng.standard_normal((6000, 128)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = rng.standard_normal((300, 128)).astype(np.float32); xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    ix = agp().IndexFlatL2(128, precision="fp16_screen"); ix.add(xb)
    D, I = ix.search(xq, 25)
    assert ix.get_stats() == (300, 0)
    D64, I64 = orc.knn_fp64(xq, xb, 25)
    ok, msg = orc.compare_knn(D, I, D64.astype(np.float32), I64)
    assert ok, msg
    # (b) 3 distinct rows repeated 2000 times: every band overflows -> all queries re-run exactly, ids ascending
    xb2 = xb[rng.integers(0, 3, 6000)]
    ix2 = agp().IndexFlatL2(128, precision="fp16_screen"); ix2.add(xb2)
    D2, I2 = ix2.search(xq, 25)
    assert ix2.get_stats()[1] == 300
    Dr, Ir = orc.knn_fp32(xq, xb2, 25)
    ok, msg = orc.compare_knn(D2, I2, Dr, Ir, xq=xq, xb=xb2, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    # (c) magnitudes spread over 12 decades inside one database, and queries likewise
    scale_b = (10.0 ** rng.uniform(-6, 6, size=(4000, 1))).astype(np.float32)
    xb3 = (rng.standard_normal((4000, 64)).astype(np.float32)) * scale_b
    xq3 = xb3[rng.integers(0, 4000, 120)] * (1 + 1e-3 * rng.standard_normal((120, 64)).astype(np.float32))
    ix3 = agp().IndexFlatL2(64, precision="fp16_screen"); ix3.add(xb3[:1000]); ix3.add(xb3[1000:])
    D3, I3 = ix3.search(xq3, 5)
    D64, I64 = orc.knn_fp64(xq3, xb3, 5)
    ok, msg = orc.compare_knn(D3, I3, D64.astype(np.float32), I64, xq=xq3, xb=xb3, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    # (d) reset picks a new database scale
    ix3.reset(); ix3.add(xb[:, :64] * 1e-3)
    D4, I4 = ix3.search(xq[:, :64] * 1e-3, 10)
    D64, I64 = orc.knn_fp64(xq[:, :64] * 1e-3, xb[:, :64] * 1e-3, 10)
    ok, msg = orc.compare_knn(D4, I4, D64.astype(np.float32), I64)
    assert ok, msg


def test_add_in_chunks_reset_and_reuse():
    rng = np.random.default_rng(12)
    xb = rng.standard_normal((5000, 96)).astype(np.float32)
    xq = rng.standard_normal((64, 96)).astype(np.float32)
    D1, I1 = search(xb, xq, 25, chunks=1)
    D7, I7 = search(xb, xq, 25, chunks=7)
    np.testing.assert_array_equal(I1, I7); np.testing.assert_array_equal(D1, D7)
    ix = agp().IndexFlatL2(96)
    D0, I0 = ix.search(xq, 4)                                            # empty index
    assert (I0 == -1).all() and (D0 == FLT_MAX).all()
    ix.add(xb[:10]); ix.reset(); assert ix.ntotal == 0
    ix.add(xb)
    D2, I2 = ix.search(xq, 25)
    np.testing.assert_array_equal(I1, I2); np.testing.assert_array_equal(D1, D2)
    xb_copy = xb.copy(); ix2 = agp().IndexFlatL2(96); ix2.add(xb_copy); xb_copy[:] = 0   # index owns a copy
    np.testing.assert_array_equal(ix2.search(xq, 25)[1], I1)


def test_wrapper_coercions_like_faiss():
    rng = np.random.default_rng(13)
    xb64 = rng.standard_normal((400, 20))                                 # float64
    xq_f = np.asfortranarray(rng.standard_normal((30, 20)).astype(np.float32))
    ix = agp().IndexFlatL2(20)
    ix.add(xb64); ix.add(xb64[::2])                                        # non-contiguous view
    assert ix.ntotal == 600 and ix.d == 20 and ix.is_trained
    D, I = ix.search(xq_f, 7)
    Dr, Ir = orc.knn_fp32(np.ascontiguousarray(xq_f), np.concatenate([xb64, xb64[::2]]).astype(np.float32), 7)
    ok, msg = orc.compare_knn(D, I, Dr, Ir)
    assert ok, msg
    Dp, Ip = np.empty((30, 7), np.float32), np.empty((30, 7), np.int64)
    Do, Io = ix.search(xq_f, 7, D=Dp, I=Ip)
    assert Do is Dp and Io is Ip
    np.testing.assert_array_equal(Ip, I)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((2, 19), np.float32))
    with pytest.raises(AssertionError):
        ix.search(np.zeros((2, 21), np.float32), 3)
    with pytest.raises(AssertionError):
        ix.search(xq_f, 0)
    with pytest.raises(RuntimeError):
        ix.search(xq_f, 100000)


def test_torch_tensors_cpu_and_cuda():
    import torch
    rng = np.random.default_rng(14)
    xb = rng.standard_normal((3000, 128)).astype(np.float32)
    xq = rng.standard_normal((100, 128)).astype(np.float32)
    Dr, Ir = orc.knn_fp32(xq, xb, 15)
    ix = agp().IndexFlatL2(128)
    ix.add(torch.from_numpy(xb).cuda())
    D, I = ix.search(torch.from_numpy(xq).cuda(), 15)
    assert D.is_cuda and I.is_cuda and I.dtype == torch.int64
    ok, msg = orc.compare_knn(D.cpu().numpy(), I.cpu().numpy(), Dr, Ir)
    assert ok, msg
    Dc, Ic = ix.search(torch.from_numpy(xq), 15)                           # CPU tensor in -> CPU tensor out
    assert not Dc.is_cuda
    np.testing.assert_array_equal(Ic.numpy(), I.cpu().numpy())
    with torch.cuda.stream(torch.cuda.Stream()):                           # runs on the caller's current stream
        D2, I2 = ix.search(torch.from_numpy(xq).cuda(), 15)
    torch.cuda.synchronize()
    assert torch.equal(I2, I)


def test_recall_kernel_matches_reference_loop():
    a = agp()
    rng = np.random.default_rng(15)
    nq, k = 500, 20
    I = rng.integers(0, 300, size=(nq, k)).astype(np.int64)
    I[5, 3:] = -1
    pos = np.empty(nq, dtype=object)
    for q in range(nq):
        pos[q] = rng.integers(0, 300, size=rng.integers(0, 12)).astype(np.int64)
    vals = [1, 5, 10, 20]
    hits = a.recall_hits(I, pos, vals)
    expect = reference_recall_loop(I, pos, vals)
    np.testing.assert_array_equal(hits / nq * 100, expect)
    import torch
    hits_dev = a.recall_hits(torch.from_numpy(I).cuda(), pos, vals)
    np.testing.assert_array_equal(hits_dev, hits)


@pytest.mark.parametrize("mode", ["partial", "full"])
def test_mined_triplets_identical_to_oracle(mode):
    from agplace_b200 import mining
    p = make_mining_problem(31, database_num=1500, queries_num=120, d=256)
    out = []
    for index_cls in (None, orc.IndexFlatL2):
        miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10,
                                    neg_samples_num=1000, index_cls=index_cls)
        np.random.seed(0)
        f = miner.compute_triplets_partial if mode == "partial" else miner.compute_triplets_full
        out.append(f(p.cache, 60))
    np.testing.assert_array_equal(out[0], out[1])


def test_batched_mining_identical_to_the_per_query_loop():
    """one best_of_lists + one search_masked per refresh must mine exactly the triplets the
    reference's per-query loop (two fresh indexes per query) mines -- same RNG draws, same arithmetic, same ties."""
    from agplace_b200 import mining
    p = make_mining_problem(33, database_num=3000, queries_num=400, d=256)
    # duplicate some database rows so that exact distance ties occur among the negatives
    p.cache[100:140] = p.cache[200:240]
    out = []
    for batched in (False, True):
        miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10,
                                    neg_samples_num=1000)
        np.random.seed(3)
        f = miner.compute_triplets_partial_batched if batched else miner.compute_triplets_partial
        out.append(f(p.cache, 300))
    np.testing.assert_array_equal(out[0], out[1])
    assert out[0].shape == (300, 12) and out[0].dtype == np.int64


def test_batched_mining_with_long_exclusion_lists():
    """k + longest exclusion list > 256 left the tensor-core screen for the expansion-form fp32 tiles
    (different distances and near-tie order), and > 512 failed outright.  Both regimes must still mine exactly the
    per-query loop's triplets: the engine re-ranks in the exact difference form, the miner routes over-long lists
    through the reference's own per-query call."""
    from agplace_b200 import mining
    p = make_mining_problem(37, database_num=2000, queries_num=120, d=128)
    p.cache[100:140] = p.cache[200:240]                       # exact ties among the negatives
    rng = np.random.default_rng(5)
    for q in range(p.queries_num):                            # soft positives: 0, ~300 or ~700 of the 2000 rows
        extra = rng.choice(p.database_num, size=(0, 600, 1400)[q % 3], replace=False)
        p.soft[q] = np.union1d(p.soft[q], extra).astype(np.int64)
    out = []
    for batched in (False, True):
        miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10,
                                    neg_samples_num=1000)
        np.random.seed(9)
        f = miner.compute_triplets_partial_batched if batched else miner.compute_triplets_partial
        out.append(f(p.cache, 90))
    np.testing.assert_array_equal(out[0], out[1])
    # the engine call itself: 300 < k + |exclude| <= 512 answers, identical to a fresh index over the surviving rows
    import agplace_b200
    xb = p.cache[:1000]
    xq = p.cache[p.database_num:p.database_num + 40]
    excl = [np.sort(rng.choice(1000, size=rng.integers(260, 480), replace=False)).astype(np.int64) for _ in range(40)]
    ix = agplace_b200.IndexFlatL2(p.d); ix.add(xb)
    D, I = ix.search_masked(xq, 12, excl)
    for q in range(40):
        keep = np.setdiff1d(np.arange(1000), excl[q])
        one = agplace_b200.IndexFlatL2(p.d); one.add(xb[keep])
        D1, I1 = one.search(xq[q:q + 1], 12)
        np.testing.assert_array_equal(I[q], keep[I1[0]])
        np.testing.assert_array_equal(D[q], D1[0])
    with pytest.raises(RuntimeError, match="AGP_MAX_K"):
        ix.search_masked(xq[:25], 12, [np.arange(600, dtype=np.int64)] * 25)


def test_counter_based_rows_are_identical_on_cpu_and_gpu():
    """bench.py generates cfg4 / cfg5 shards on the device and regenerates sampled rows on the CPU for verification:
    the two generators must agree bit for bit."""
    import torch
    from agplace_b200 import synth
    for d, seed, (a, b) in [(512, 3, (9_990_000, 9_991_111)), (4096, 4, (123_456, 123_700)), (33, 1, (0, 500))]:
        dev = synth.counter_rows_device(a, b, d, seed, torch.device("cuda", 0)).cpu().numpy()
        cpu = synth.counter_rows(np.arange(a, b), d, seed)
        np.testing.assert_array_equal(dev, cpu)
        norms = np.linalg.norm(cpu.astype(np.float64), axis=1)
        assert abs(norms.mean() - 1.0) < 0.02 and norms.std() < 2.0 / np.sqrt(d)


def test_full_batched_mining_identical_to_the_per_query_loop():
    """compute_triplets_full (kitti360:1022-1049) batched through search_subset: same triplets AND the same neg_cache
    over two consecutive refreshes (the second one feeds on the first one's neg_cache)."""
    from agplace_b200 import mining
    p = make_mining_problem(35, database_num=2500, queries_num=300, d=256)
    p.cache[100:140] = p.cache[200:240]          # exact distance ties among the negatives
    out, caches = [], []
    for batched in (False, True):
        miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10,
                                    neg_samples_num=600)
        np.random.seed(5)
        f = miner.compute_triplets_full_batched if batched else miner.compute_triplets_full
        out.append([f(p.cache, 200).copy(), f(p.cache, 200).copy()])
        caches.append(miner.neg_cache)
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    for a, b in zip(caches[0], caches[1]):
        np.testing.assert_array_equal(a, b)
    assert out[0][0].shape == (200, 12) and out[0][0].dtype == np.int64


def test_search_subset_equals_a_fresh_index_over_each_list():
    import agplace_b200
    rng = np.random.default_rng(19)
    xb = rng.standard_normal((900, 48)).astype(np.float32)
    xb[300:330] = xb[10:40]                      # duplicates: ties resolve by position in the list
    xq = rng.standard_normal((70, 48)).astype(np.float32)
    cands = [np.sort(rng.choice(900, size=rng.integers(0, 400), replace=False)).astype(np.int64) for _ in range(70)]
    cands[3] = np.array([], dtype=np.int64)      # empty list: all padding
    cands[4] = np.array([7, 7, 310, 20], dtype=np.int64)   # repeated ids and fewer than k candidates
    ix = agplace_b200.IndexFlatL2(48); ix.add(xb)
    for k in (1, 10, 70):
        D, I = ix.search_subset(xq, k, cands)
        assert D.shape == (70, k) and I.dtype == np.int64
        for q in range(70):
            Dr, Ir = orc.knn_fp32(xq[q:q + 1], xb[cands[q]].reshape(-1, 48), k) if len(cands[q]) else (
                np.full((1, k), np.float32(3.4028234663852886e38)), np.full((1, k), -1, dtype=np.int64))
            ok, msg = orc.compare_knn(D[q:q + 1], I[q:q + 1], Dr, Ir)
            assert ok, f"k={k} query {q}: {msg}"
    # bit-identical to the engine's own one-query search over the gathered rows (the reference call pattern)
    D, I = ix.search_subset(xq[:8], 10, cands[8:16])
    for q in range(8):
        one = agplace_b200.IndexFlatL2(48); one.add(xb[cands[8 + q]])
        D1, I1 = one.search(xq[q:q + 1], 10)
        np.testing.assert_array_equal(I[q], I1[0])
        np.testing.assert_array_equal(D[q], D1[0])
    with pytest.raises(RuntimeError):            # ids must be rows of the index
        ix.search_subset(xq[:1], 5, [np.array([900], dtype=np.int64)])


def test_search_masked_equals_search_on_the_surviving_rows():
    import agplace_b200
    rng = np.random.default_rng(17)
    xb = rng.standard_normal((700, 64)).astype(np.float32)
    xq = rng.standard_normal((90, 64)).astype(np.float32)
    exclude = [np.sort(rng.choice(700, size=rng.integers(0, 40), replace=False)).astype(np.int64) for _ in range(90)]
    def check(xb_, xq_, exclude_, k):
        ix = agplace_b200.IndexFlatL2(64); ix.add(xb_)
        D, I = ix.search_masked(xq_, k, exclude_)
        for q in range(len(xq_)):
            keep = np.setdiff1d(np.arange(len(xb_)), exclude_[q])
            Dr, Ir = orc.knn_fp32(xq_[q:q + 1], xb_[keep], k)      # nq = 1: faiss's exact difference form
            Ir = np.where(Ir >= 0, keep[np.clip(Ir, 0, len(keep) - 1)], -1)
            ok, msg = orc.compare_knn(D[q:q + 1], I[q:q + 1], Dr, Ir)
            assert ok, f"query {q}: {msg}"
    check(xb, xq, exclude, 10)
    # fewer survivors than k: padded (FLT_MAX, -1) like faiss; nq < 20 goes through the difference-form path
    few = [np.arange(25, dtype=np.int64), np.array([], dtype=np.int64), np.arange(5, 30, dtype=np.int64)]
    check(xb[:30], xq[:3], few, 10)
    with pytest.raises(RuntimeError):                          # k + longest exclusion list is bounded by AGP_MAX_K
        big = agplace_b200.IndexFlatL2(64); big.add(xb)
        big.search_masked(xq[:1], 10, [np.arange(600, dtype=np.int64)])
    bd, bp = agplace_b200.best_of_lists(xq[:3], xb[[4, 9, 9, 2, 7]], np.array([0, 2, 2, 5]))
    d0 = ((xb[[4, 9]].astype(np.float64) - xq[0]) ** 2).sum(1)
    d2 = ((xb[[9, 2, 7]].astype(np.float64) - xq[2]) ** 2).sum(1)
    assert bp.tolist() == [int(np.argmin(d0)), -1, int(np.argmin(d2))]
    np.testing.assert_allclose(bd[[0, 2]], [d0.min(), d2.min()], rtol=2e-6)


def test_virtual_shards_merge_equals_single_index():
    """8 virtual shards on one GPU exercise id bases + the cross-shard merge kernel (multi-GPU row)."""
    import ctypes
    import torch
    from agplace_b200 import _lib
    from agplace_b200.sharded import shard_bounds
    rng = np.random.default_rng(16)
    n, nq, d, k, G = 20000, 333, 64, 100, 8
    xb = rng.integers(-4, 5, size=(n, d)).astype(np.float32)               # lattice: many exact ties across shards
    xq = rng.integers(-4, 5, size=(nq, d)).astype(np.float32)
    single = agp().IndexFlatL2(d); single.add(xb)
    Ds, Is = single.search(xq, k)
    xq_dev = torch.from_numpy(xq).cuda()
    Dl = torch.empty((G, nq, k), dtype=torch.float32, device="cuda")
    Il = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
    keep = []
    for g, (a, b) in enumerate(shard_bounds(n, G)):
        sh = agp().IndexFlatL2(d); sh.add(xb[a:b]); sh.set_id_base(a)
        sh.search(xq_dev, k, D=Dl[g], I=Il[g])
        keep.append(sh)
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    lib = _lib.load()
    for id_bound in (n, 0):
        _lib.check(lib.agp_merge_topk(0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), nq, k, G,
                                      ctypes.c_void_p(Dl.data_ptr()), nq * k, ctypes.c_void_p(Il.data_ptr()), nq * k, id_bound,
                                      ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr())), "agp_merge_topk")
        torch.cuda.synchronize()
        np.testing.assert_array_equal(D.cpu().numpy(), Ds)
        np.testing.assert_array_equal(I.cpu().numpy(), Is)
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    np.testing.assert_array_equal(Is, Ir)


def test_full_size_cfg2_properties():
    """BASELINE cfg2 at full size (100k x 512, 20k queries, k = 50): size-independent properties plus an
    oracle check on a query sample and a device-side cross-check against the fp32 SIMT path."""
    import torch
    from agplace_b200 import synth
    c = synth.CONFIGS["cfg2"]
    xb = synth.descriptors(c["n"], c["d"], c["seed"], "db")
    xq = synth.descriptors(c["nq"], c["d"], c["seed"] + 7, "q")
    xq[:64] = xb[1000:1064] * 1.0                                          # planted exact matches
    ix = agp().IndexFlatL2(c["d"]); ix.add(xb)
    D, I = ix.search(xq, c["k"])
    assert np.all(np.diff(D, axis=1) >= 0), "sortedness"
    assert ((I >= 0) & (I < c["n"])).all()
    assert all(len(set(row)) == c["k"] for row in I[::97]), "no duplicate ids in a result row"
    np.testing.assert_array_equal(I[:64, 0], np.arange(1000, 1064))        # self-match is rank 0 ...
    assert (D[:64, 0] <= 8 * 2.0 ** -24 * 2).all()                          # ... at distance ~0 (expansion-form floor)
    sample = np.arange(0, c["nq"], 101)
    Dr, Ir = orc.knn_fp32(xq[sample], xb, c["k"])
    ok, msg = orc.compare_knn(D[sample], I[sample], Dr, Ir, xq=xq[sample], xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg
    simt = agp().IndexFlatL2(c["d"], precision="fp32_simt"); simt.add(xb)
    D2, I2 = simt.search(xq[:4096], c["k"])
    # fp32 FMA expansion form over d = 512 is itself ~11 ulp of (|q|^2 + |x|^2) away from the exact value
    ok, msg = orc.compare_knn(D[:4096], I[:4096], D2, I2, xq=xq[:4096], xb=xb, abs_floor_eps=32 * 2.0 ** -24)
    assert ok, "tensor-core vs fp32 SIMT: " + msg
    # idempotence: same call, same bits
    Db, Ib = ix.search(xq, c["k"])
    np.testing.assert_array_equal(Ib, I); np.testing.assert_array_equal(Db, D)


@pytest.mark.parametrize("precision", MODES)
def test_engine_reproduces_the_published_output_of_the_faiss_tutorial(precision):
    """Real faiss output (tests/golden/faiss_tutorial_1flat.json: faiss's tutorial/python/1-Flat.py as published on
    its wiki) reproduced by the CUDA engine through the C ABI: every neighbour id, distances to printed precision."""
    from tests.helpers import faiss_tutorial_data
    xb, xq, pub = faiss_tutorial_data()
    ix = agp().IndexFlatL2(64, precision=precision); ix.add(xb)
    D, I = ix.search(xb[:5], 4)
    np.testing.assert_array_equal(I, np.array(pub["sanity_I"]))
    np.testing.assert_allclose(D, np.array(pub["sanity_D"]), rtol=1e-5, atol=2e-5)
    D, I = ix.search(xq, 4)
    np.testing.assert_array_equal(I[:5], np.array(pub["search_I_first5"]))
    np.testing.assert_array_equal(I[-5:], np.array(pub["search_I_last5"]))
    Dr, Ir = orc.knn_fp32(xq, xb, 4)
    ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg


def test_screen_variants_return_identical_results():
    """The screen kernel's alternative code paths -- the branchy scan (also the fallback when a candidate bundle
    straddles a 4 GB line), rounds without the pair exchange, sweeps without the first-tile bootstrap, 512-slot lists --
    are selected per index through agp_index_set_knob (the library reads no environment variable); the exact finish
    makes every variant return the same bits.  Shapes cover whole waves, the split remainder and a ragged last tile."""
    rng = np.random.default_rng(23)
    for (n, nq, d, k) in [(30011, 19200, 64, 20), (9000, 1500, 128, 50), (5000, 300, 32, 100)]:
        xb = rng.standard_normal((n, d)).astype(np.float32)
        xq = rng.standard_normal((nq, d)).astype(np.float32)
        ix = agp().IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
        D0, I0 = ix.search(xq, k)
        sample = np.arange(0, nq, max(1, nq // 200))
        Dr, Ir = orc.knn_fp32(xq[sample], xb, k)
        ok, msg = orc.compare_knn(D0[sample], I0[sample], Dr, Ir, xq=xq[sample], xb=xb, abs_floor_eps=8 * 2.0 ** -24)
        assert ok, msg
        for flags, e in [(1, 0), (4, 0), (8, 0), (13, 0), (0, 16), (5, 16)]:
            ix.set_knob("screen_flags", flags)
            ix.set_knob("screen_e", e)
            D, I = ix.search(xq, k)
            np.testing.assert_array_equal(I, I0, err_msg=f"flags={flags} E={e} shape={(n, nq, d, k)}")
            np.testing.assert_array_equal(D, D0)
        assert ix.get_stats()[1] == 0
    with pytest.raises(RuntimeError, match="unknown knob"):
        ix.set_knob("skip_epi", 1)      # result-changing probes do not exist in the product build


def test_native_library_was_used():
    from agplace_b200 import _lib
    before = _lib.kernel_launches()
    search(np.random.default_rng(1).standard_normal((300, 32)).astype(np.float32),
           np.random.default_rng(2).standard_normal((40, 32)).astype(np.float32), 5)
    assert _lib.kernel_launches() > before


def test_cuda_tensor_search_is_asynchronous():
    """N1 contract (include/agpknn.h): a device-in / device-out search only enqueues work -- no host synchronisation
    anywhere on the path (the screen's overflow fallback runs on the device).  The call must return while the GPU is
    still busy, and an all-overflow batch (mass duplicates) must behave the same way."""
    import time
    import torch
    import agplace_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    xb = torch.randn((400_000, 256), generator=g, device=dev)
    xq = torch.randn((40_000, 256), generator=g, device=dev)
    ix = agplace_b200.IndexFlatL2(256, device=0); ix.add(xb)
    ix.search(xq[:4096], 20)                     # builds the fp16 plane, sizes the scratch buffers
    ix.search(xq, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D, I = ix.search(xq, 20)
    t_call = time.perf_counter() - t0
    busy_at_return = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    assert busy_at_return, "the search call waited for the GPU"
    assert t_call < 0.5 * t_total, f"call took {t_call * 1e3:.2f} ms of {t_total * 1e3:.2f} ms"
    # all queries overflow (3 distinct rows x many copies): still no host round trip, still exact
    xb2 = xb[torch.randint(0, 3, (60_000,), generator=g, device=dev)]
    ix2 = agplace_b200.IndexFlatL2(256, device=0); ix2.add(xb2)
    ix2.search(xq[:512], 25)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D2, I2 = ix2.search(xq[:512], 25)
    t_call = time.perf_counter() - t0
    busy_at_return = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    assert busy_at_return or t_call < 2e-3
    assert ix2.get_stats()[1] == 1024
    Dr, Ir = orc.knn_fp32(xq[:64].cpu().numpy(), xb2.cpu().numpy(), 25)
    ok, msg = orc.compare_knn(D2[:64].cpu().numpy(), I2[:64].cpu().numpy(), Dr, Ir, xq=xq[:64].cpu().numpy(), xb=xb2.cpu().numpy(),
                              abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg


@pytest.mark.parametrize("n,d,nq,k", [(20_000, 64, 40_000, 10),      # copy-bound: [3/4 | 1/4]
                                      (60_000, 256, 19_000, 30),     # search-bound, about one wave: [18 | 24 | rest] tiles
                                      (60_000, 256, 30_000, 30),     # search-bound, 1.6 waves: ramp [9 | 18 | 37] tiles + waves
                                      (3_000, 512, 61_000, 5)])      # more than three waves: one wave per chunk
def test_every_automatic_chunk_schedule_returns_the_device_path_bits(n, d, nq, k):
    """The host pipeline picks its chunk schedule from the batch size and from which stage limits it (agpknn.cu:
    search_host_pipelined); each regime, and the staging-piece and explicit-cut knobs, return the bits of one
    device-resident search."""
    import torch
    import agplace_b200
    rng = np.random.default_rng(n + nq)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    ix = agplace_b200.IndexFlatL2(d); ix.add(xb)
    Dd, Id = ix.search(torch.from_numpy(xq).cuda(), k)
    Dd, Id = Dd.cpu().numpy(), Id.cpu().numpy()
    for knobs in ({}, {"pipe_sched": 1}, {"pipe_piece_kb": 1024}, {"pipe_cut1": 2304, "pipe_cut2": 6912, "pipe_cut3": nq - 1}):
        for name, v in knobs.items():
            ix.set_knob(name, v)
        D, I = ix.search(xq, k)
        for name in knobs:
            ix.set_knob(name, 0)
        np.testing.assert_array_equal(I, Id, err_msg=str(knobs))
        np.testing.assert_array_equal(D, Dd, err_msg=str(knobs))


@pytest.mark.parametrize("pinned", [False, True])
def test_host_pipeline_returns_the_device_path_bits(pinned):
    """numpy in / numpy out goes through the chunked H2D | compute | D2H pipeline (agp_index_search with host buffers):
    every chunk schedule returns exactly what one device-resident search returns."""
    import torch
    import agplace_b200
    rng = np.random.default_rng(41)
    n, nq, d, k = 60_000, 21_000, 256, 30          # 21.5 MB of queries: the automatic schedule cuts three chunks
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    ix = agplace_b200.IndexFlatL2(d); ix.add(xb)
    Dd, Id = ix.search(torch.from_numpy(xq).cuda(), k)
    Dd, Id = Dd.cpu().numpy(), Id.cpu().numpy()
    xin = torch.from_numpy(xq).pin_memory().numpy() if pinned else xq
    for chunk in (0, 1000, 4096, 7777, 30_000):
        ix.set_knob("pipe_chunk", chunk)
        D, I = ix.search(xin, k)
        np.testing.assert_array_equal(I, Id, err_msg=f"pipe_chunk={chunk}")
        np.testing.assert_array_equal(D, Dd)
        if pinned:
            Dp = torch.empty((nq, k), dtype=torch.float32).pin_memory()
            Ip = torch.empty((nq, k), dtype=torch.int64).pin_memory()
            ix.search(xin, k, D=Dp.numpy(), I=Ip.numpy())
            np.testing.assert_array_equal(Ip.numpy(), Id)
            np.testing.assert_array_equal(Dp.numpy(), Dd)
    sample = np.arange(0, nq, 211)
    Dr, Ir = orc.knn_fp32(xq[sample], xb, k)
    ok, msg = orc.compare_knn(Dd[sample], Id[sample], Dr, Ir, xq=xq[sample], xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg


def test_lockstep_sweep_is_only_a_hint():
    """Large planes are swept in step (the TMA producers of a full wave meet every few tiles at a global counter).  The
    meeting points must never be able to hang: two searches that share one GPU (two host threads; virtual shards of a
    multi-device index) each leave part of the other's grid non-resident.  Forced on for a small plane here; results
    must be the bits of the default configuration."""
    import threading
    import agplace_b200
    rng = np.random.default_rng(61)
    n, nq, d, k = 70_000, 19_200, 64, 20                   # 75 pair tiles: one full wave + a remainder
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    ref = agplace_b200.IndexFlatL2(d); ref.add(xb)
    D0, I0 = ref.search(xq, k)
    out = {}

    def worker(tag):
        ix = agplace_b200.IndexFlatL2(d); ix.add(xb)
        ix.set_knob("screen_lockstep", 4)
        for _ in range(4):
            out[tag] = ix.search(xq, k)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive(), "a search with lockstep enabled did not return"
    for tag, (D, I) in out.items():
        np.testing.assert_array_equal(I, I0, err_msg=f"thread {tag}")
        np.testing.assert_array_equal(D, D0)
    multi = agplace_b200.IndexFlatL2(d, devices=[0, 0]); multi.add(xb)
    multi.set_knob("screen_lockstep", 4)
    D, I = multi.search(xq, k)
    np.testing.assert_array_equal(I, I0)
    np.testing.assert_array_equal(D, D0)


def test_concurrent_calls_on_one_index_take_turns():
    """faiss lets several threads search one index at the same time; the engine's scratch buffers and streams belong to the
    call in flight, so the C ABI serialises the calls on one index (a per-index lock).  Host threads hammering ONE index
    with different batch sizes (screen path, small-batch path, host pipeline) all get the single-threaded bits."""
    import threading
    import agplace_b200
    rng = np.random.default_rng(67)
    n, d = 30_000, 128
    xb = rng.standard_normal((n, d)).astype(np.float32)
    ix = agplace_b200.IndexFlatL2(d); ix.add(xb)
    jobs = [(rng.standard_normal((nq, d)).astype(np.float32), k) for nq, k in ((1, 10), (7, 3), (300, 20), (2500, 50), (12_000, 10))]
    want = [ix.search(xq, k) for xq, k in jobs]
    errors = []

    def worker(tag):
        try:
            for rep in range(6):
                j = (tag + rep) % len(jobs)
                D, I = ix.search(*jobs[j])
                np.testing.assert_array_equal(I, want[j][1], err_msg=f"thread {tag} job {j}")
                np.testing.assert_array_equal(D, want[j][0])
        except Exception as e:      # noqa: BLE001 -- reported by the main thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive()
    assert not errors, errors[0]


@pytest.mark.parametrize("n,d,nq,k", [(12_000, 64, 300, 10),        # 2 pair tiles x 47 database tiles over 74 pairs: segments of 1-2 tiles
                                      (5_000, 64, 5_000, 10),       # 20 x 20
                                      (30_000, 128, 16_128, 50),    # 63 x 118: the case equal ranges serve worst
                                      (30_000, 128, 10_240, 1),     # 40 x 118
                                      (70_000, 64, 24_000, 100),    # one full wave + 20 tiles of remainder
                                      (9_000, 576, 4_200, 20),      # d_pad > 512: queries streamed with the database
                                      (300, 32, 2_100, 300)])       # 2 database tiles, k > 256 (1024-slot lists)
def test_balanced_remainder_returns_the_bits_of_equal_ranges(n, d, nq, k):
    """The remainder of a batch (pair tiles beyond whole waves) is either split into equal database ranges per pair tile or
    cut into one contiguous segment per CTA pair that crosses pair-tile boundaries (ScreenParams::balanced).  The exact
    finish makes the result independent of the decomposition: forced on, forced off and automatic return the same bits,
    and those equal the oracle's neighbours."""
    import torch
    import agplace_b200
    rng = np.random.default_rng(n + nq + k)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    ix = agplace_b200.IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
    xq_dev = torch.from_numpy(xq).cuda()
    out = {}
    for mode in (0, 1, -1):
        ix.set_knob("screen_balanced", mode)
        D, I = ix.search(xq_dev, k)
        out[mode] = (D.cpu().numpy(), I.cpu().numpy())
    for mode in (1, -1):
        np.testing.assert_array_equal(out[mode][1], out[0][1], err_msg=f"screen_balanced={mode}")
        np.testing.assert_array_equal(out[mode][0], out[0][0], err_msg=f"screen_balanced={mode}")
    sample = np.arange(0, nq, max(1, nq // 200))
    Dr, Ir = orc.knn_fp32(xq[sample], xb, k)
    ok, msg = orc.compare_knn(out[1][0][sample], out[1][1][sample], Dr, Ir, xq=xq[sample], xb=xb, abs_floor_eps=32 * 2.0 ** -24)
    assert ok, msg
    assert ix.get_stats()[1] == 0          # nothing went through the exact fallback
