"""CPU checks of the measurement helpers: the counter-based generator (numpy == torch integer ops), the sampled
verification bench.py runs at full size (it must notice a wrong id, a wrong distance and a missed neighbour), and the
bounded CPU sample of the reference arm."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from agplace_b200 import synth  # noqa: E402
from oracle import flatl2_oracle as orc  # noqa: E402


def test_counter_rows_numpy_equals_torch_and_is_row_addressable():
    import torch
    for d, seed in ((512, 3), (4096, 4), (33, 1)):
        a = synth.counter_rows(np.arange(1000, 1400), d, seed)
        b = synth.counter_rows_device(1000, 1400, d, seed, torch.device("cpu")).numpy()
        np.testing.assert_array_equal(a, b)
        ids = np.array([1399, 1000, 1234, 1234])
        np.testing.assert_array_equal(synth.counter_rows(ids, d, seed), a[ids - 1000])      # any row, any order, repeated
        assert a.dtype == np.float32 and abs(np.linalg.norm(a, axis=1).mean() - 1.0) < 0.03
    assert not np.array_equal(synth.counter_rows(np.arange(4), 64, 1), synth.counter_rows(np.arange(4), 64, 2))


def _small_config():
    return dict(n=60_000, nq=300, d=32, k=10, seed=5, name="t", desc="t", device_generated=True)


def test_sampled_verification_accepts_the_truth_and_rejects_tampering():
    c = _small_config()
    xq = bench.host_queries(c)
    xb = synth.counter_rows(np.arange(c["n"]), c["d"], c["seed"])
    D, I = orc.knn_fp32(xq, xb, c["k"])
    ok = bench.verify_sample(c, xq, D, I)
    assert ok["ok"] and ok["ids_identical"] and ok["queries"] == 64 and ok["rows_regenerated"] > 0
    qs = np.unique(np.linspace(0, len(xq) - 1, 64).astype(np.int64))
    q = int(qs[7])
    # a wrong id at some rank (a far row in place of a neighbour)
    I2 = I.copy(); I2[q, 3] = (I[q, 3] + 12345) % c["n"]
    assert not bench.verify_sample(c, xq, D, I2)["ok"]
    # a wrong distance
    D2 = D.copy(); D2[q, 0] *= 1.01
    assert not bench.verify_sample(c, xq, D2, I)["ok"]
    # a missed neighbour: the true nearest dropped, everything shifted up by one (the lists stay sorted and plausible)
    D3, I3 = orc.knn_fp32(xq, xb, c["k"] + 1)
    D4, I4 = D.copy(), I.copy()
    D4[q], I4[q] = D3[q, 1:], I3[q, 1:]
    assert not bench.verify_sample(c, xq, D4, I4)["ok"]
    # unsorted output
    D5, I5 = D.copy(), I.copy()
    D5[q, [2, 5]], I5[q, [2, 5]] = D[q, [5, 2]], I[q, [5, 2]]
    assert not bench.verify_sample(c, xq, D5, I5)["ok"]


def test_cpu_sample_keeps_the_query_block_and_scales_by_rows():
    c = _small_config()
    orc.build()
    xq = bench.host_queries(c)
    xb, sample, scale = bench.cpu_sample(c, xq, 0.05, orc)
    assert len(sample) == min(c["nq"], 4096)
    assert 1024 <= len(xb) <= c["n"] and abs(scale - len(xb) / c["n"]) < 1e-12
    np.testing.assert_array_equal(xb, synth.counter_rows(np.arange(len(xb)), c["d"], c["seed"]))
