"""Host-side mirrors of the reference call sites, run against the CPU oracle index (CPU only)."""
from types import SimpleNamespace

import numpy as np
import pytest

from agplace_b200 import mining, recall, synth
from oracle import flatl2_oracle as orc
from tests.helpers import make_mining_problem, reference_recall_loop


def test_compute_recall_matches_reference_loop():
    ev = synth.make_eval_set(dict(n=800, nq=120, d=32, k=20, seed=11, side=200.0), correlated=0.6)
    args = SimpleNamespace(features_dim=32, recall_values=[1, 5, 10, 20])
    recalls, s = recall.compute_recall(args, ev.queries_features, ev.database_features, ev, index_cls=orc.IndexFlatL2)
    _, I = orc.knn_fp32(ev.queries_features, ev.database_features, 20)
    expect = reference_recall_loop(I, ev.get_positives(), args.recall_values)
    np.testing.assert_array_equal(recalls, expect)
    assert s.startswith("R@1: ") and s.count("R@") == 4
    assert np.all(np.diff(recalls) >= 0) and 0 < recalls[-1] <= 100


def test_recall_semantics_by_hand():
    class DS:
        queries_num = 3
        def get_positives(self):
            return np.array([np.array([7, 2]), np.array([], dtype=np.int64), np.array([5])], dtype=object)
    class FakeIndex:
        def __init__(self, d): pass
        def add(self, x): pass
        def search(self, x, k):
            I = np.array([[9, 2, 1, 0], [1, 2, 3, 4], [0, 1, 2, 5]], dtype=np.int64)
            return np.zeros((3, 4), np.float32), I
    args = SimpleNamespace(features_dim=4, recall_values=[1, 2, 4])
    r, _ = recall.compute_recall(args, np.zeros((3, 4), np.float32), np.zeros((9, 4), np.float32), DS(), index_cls=FakeIndex)
    # q0 hits at rank 1 (< 2), q1 has no positives (miss), q2 hits at rank 3 (< 4)
    np.testing.assert_allclose(r, np.array([0, 1, 2]) / 3 * 100)


def test_positives_to_csr_roundtrip():
    from agplace_b200 import positives_to_csr
    pos = np.array([np.array([3, 1]), np.array([], dtype=np.int64), np.array([9])], dtype=object)
    off, ids = positives_to_csr(pos)
    assert off.tolist() == [0, 2, 2, 3] and ids.tolist() == [3, 1, 9]
    off, ids = positives_to_csr([])
    assert off.tolist() == [0] and ids.size == 0


def test_ram_efficient_matrix_gather_semantics():
    m = mining.RAMEfficient2DMatrix((5, 3))
    m[[0, 3]] = np.arange(6, dtype=np.float64).reshape(2, 3)
    assert m[0].dtype == np.float32 and m[1] is None
    g = m[np.array([3, 0])]
    np.testing.assert_array_equal(g, [[3, 4, 5], [0, 1, 2]])


@pytest.mark.parametrize("mode", ["partial", "full"])
def test_mining_loop_is_deterministic_and_consistent(mode):
    p = make_mining_problem(21)
    def run(cache):
        miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10,
                                    neg_samples_num=200, index_cls=orc.IndexFlatL2)
        np.random.seed(0)
        f = miner.compute_triplets_partial if mode == "partial" else miner.compute_triplets_full
        return f(cache, 40)
    t1 = run(p.cache)
    rm = mining.RAMEfficient2DMatrix(p.cache.shape)
    rm[list(range(len(p.cache)))] = p.cache
    t2 = run(rm)
    np.testing.assert_array_equal(t1, t2)
    assert t1.shape == (40, 12) and t1.dtype == np.int64
    for row in t1:
        q, pos, negs = row[0], row[1], row[2:]
        assert pos in p.hard[q]
        assert not np.isin(negs, p.soft[q]).any(), "negatives must exclude soft positives"
        # the best positive is the nearest hard positive in feature space (fp64 check)
        dpos = ((p.cache[p.hard[q]].astype(np.float64) - p.cache[q + p.database_num]) ** 2).sum(1)
        assert p.hard[q][np.argmin(dpos)] == pos
        dneg = ((p.cache[negs].astype(np.float64) - p.cache[q + p.database_num]) ** 2).sum(1)
        assert np.all(np.diff(dneg) >= -1e-9), "hardest negatives come nearest first"
