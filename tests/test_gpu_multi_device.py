"""Single-process multi-device index (agp_index_create_multi): one handle, one process,
rows split over several devices, per-shard lists merged on the home device -- must return the bits a one-device
index returns.  With one GPU visible the shards are virtual (the same device listed several times: peer copies
degrade to device copies, everything else is the real code path); with >= 2 GPUs real devices are used too."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatl2_oracle as orc


def _device_sets():
    import torch
    sets = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append([0, 1])
    if torch.cuda.device_count() >= 4:
        sets.append([1, 0, 3, 2])
    return sets


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_multi_device_index_equals_single(metric):
    import torch
    import agplace_b200 as agp
    cls = agp.IndexFlatL2 if metric == "l2" else agp.IndexFlatIP
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((30011, 96)).astype(np.float32)
    xb[5000:5040] = xb[100:140]                      # exact ties across shards: resolved by global id
    xq = rng.standard_normal((1500, 96)).astype(np.float32)
    single = cls(96, device=0)
    single.add(xb[:20000]); single.add(xb[20000:])
    Ds, Is = single.search(xq, 40)
    Ds1, Is1 = single.search(xq[:3], 10)             # nq < 20: difference-form path on every shard
    Dsk, Isk = single.search(xq[:50], 512)
    current = torch.cuda.current_device()
    for devices in _device_sets():
        ix = cls(96, devices=devices)
        assert ix.device == devices[0]
        ix.add(xb[:20000]); ix.add(xb[20000:])       # two batches: two id ranges per shard
        assert ix.ntotal == len(xb)
        D, I = ix.search(xq, 40)                     # numpy in -> numpy out
        np.testing.assert_array_equal(I, Is, err_msg=str(devices))
        np.testing.assert_array_equal(D, Ds)
        D1, I1 = ix.search(xq[:3], 10)
        np.testing.assert_array_equal(I1, Is1)
        np.testing.assert_array_equal(D1, Ds1)
        Dk, Ik = ix.search(xq[:50], 512)
        np.testing.assert_array_equal(Ik, Isk)
        np.testing.assert_array_equal(Dk, Dsk)
        xt = torch.from_numpy(xq).to(f"cuda:{devices[0]}")
        Dt, It = ix.search(xt, 40)                   # CUDA in -> CUDA out on the home device, asynchronous
        assert Dt.device.index == devices[0]
        np.testing.assert_array_equal(It.cpu().numpy(), Is)
        np.testing.assert_array_equal(Dt.cpu().numpy(), Ds)
        ix.reset()
        assert ix.ntotal == 0
        ix.add(torch.from_numpy(xb[:7000]).to(f"cuda:{devices[0]}"))      # device-resident rows, one batch
        D2, I2 = ix.search(xq[:64], 5)
        ref = cls(96, device=0); ref.add(xb[:7000])
        Dr, Ir = ref.search(xq[:64], 5)
        np.testing.assert_array_equal(I2, Ir)
        np.testing.assert_array_equal(D2, Dr)
        del ix
        # the library switches devices while it works but hands the caller's current device back at every entry point
        assert torch.cuda.current_device() == current


def test_multi_device_index_through_the_reference_call_site(monkeypatch):
    """AGP_DEVICES makes the unmodified call site (test.py:27-32: IndexFlatL2(d); add; search) use several devices."""
    from types import SimpleNamespace
    import agplace_b200 as agp
    from agplace_b200 import recall, synth
    ev = synth.make_eval_set(dict(n=6000, nq=700, d=256, k=20, seed=0, side=600.0), correlated=0.3)
    args = SimpleNamespace(features_dim=256, recall_values=[1, 5, 10, 20])
    r_one, s_one = recall.compute_recall(args, ev.queries_features, ev.database_features, ev)
    monkeypatch.setenv("AGP_DEVICES", "0,0")
    assert agp.IndexFlatL2(256).devices == [0, 0]
    r_multi, s_multi = recall.compute_recall(args, ev.queries_features, ev.database_features, ev)
    r_cpu, s_cpu = recall.compute_recall(args, ev.queries_features, ev.database_features, ev, index_cls=orc.IndexFlatL2)
    np.testing.assert_array_equal(r_multi, r_one)
    np.testing.assert_array_equal(r_multi, r_cpu)
    assert s_multi == s_one == s_cpu


def test_multi_device_large_batch_pipeline():
    """Host queries big enough for the chunked pipeline (several chunks, staging ring reused across devices)."""
    import agplace_b200 as agp
    rng = np.random.default_rng(8)
    xb = rng.standard_normal((40000, 256)).astype(np.float32)
    xq = rng.standard_normal((21000, 256)).astype(np.float32)
    single = agp.IndexFlatL2(256, device=0); single.add(xb)
    Ds, Is = single.search(xq, 30)
    for devices in _device_sets():
        ix = agp.IndexFlatL2(256, devices=devices); ix.add(xb)
        for chunk in (0, 3000):
            ix.set_knob("pipe_chunk", chunk)
            D, I = ix.search(xq, 30)
            np.testing.assert_array_equal(I, Is, err_msg=f"{devices} chunk={chunk}")
            np.testing.assert_array_equal(D, Ds)
        assert ix.get_stats()[1] == 0
    sample = np.arange(0, len(xq), 301)
    Dr, Ir = orc.knn_fp32(xq[sample], xb, 30)
    ok, msg = orc.compare_knn(Ds[sample], Is[sample], Dr, Ir, xq=xq[sample], xb=xb, abs_floor_eps=8 * 2.0 ** -24)
    assert ok, msg


def test_multi_device_edge_cases():
    """Empty index, fewer rows than shards, k beyond the row count, faiss's input coercions, and the entry points that are
    one-device only."""
    import agplace_b200 as agp
    FLT_MAX = np.float32(3.4028234663852886e38)
    rng = np.random.default_rng(11)
    ix = agp.IndexFlatL2(24, devices=[0, 0, 0])
    xq = rng.standard_normal((33, 24)).astype(np.float32)
    D, I = ix.search(xq, 5)                                   # empty: faiss padding
    assert (I == -1).all() and (D == FLT_MAX).all()
    xb = rng.standard_normal((2, 24))                         # float64, 2 rows on 3 shards (one shard stays empty)
    ix.add(xb)
    assert ix.ntotal == 2
    D, I = ix.search(xq.astype(np.float64), 4)                # k > ntotal
    ref = agp.IndexFlatL2(24); ref.add(xb)
    Dr, Ir = ref.search(xq, 4)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    assert (I[:, 2:] == -1).all()
    ix.add(np.asfortranarray(rng.standard_normal((700, 24)).astype(np.float32)))     # non-contiguous input
    assert ix.ntotal == 702
    with pytest.raises(AssertionError):
        ix.search(xq[:, :20], 3)
    with pytest.raises(RuntimeError, match="exceeds"):
        ix.search(xq, 513)
    with pytest.raises(RuntimeError, match="one-device"):
        ix.search_masked(xq, 3, [np.array([0])] * len(xq))
    with pytest.raises(RuntimeError, match="one-device"):
        ix.search_subset(xq, 3, [np.array([0, 1])] * len(xq))
    Dp = np.empty((33, 7), np.float32); Ip = np.empty((33, 7), np.int64)
    D2, I2 = ix.search(xq, 7, D=Dp, I=Ip)                     # preallocated outputs, like faiss
    assert D2 is Dp and I2 is Ip
