"""N > 1 host logic on CPU: world_size-2 gloo processes drive ShardedIndexFlatL2 with the CPU oracle
as the local index and a numpy merge as the checker; results must equal one unsharded oracle index."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_ip(rank, world, port, shard, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from agplace_b200.sharded import ShardedIndexFlatIP
        from oracle import flatl2_oracle as orc
        from tests.helpers import numpy_merge
        rng = np.random.default_rng(78)
        xb = rng.integers(-5, 6, size=(401, 12)).astype(np.float32)      # lattice: exact product ties across shards
        xq = rng.integers(-5, 6, size=(33, 12)).astype(np.float32)
        ix = ShardedIndexFlatIP(12, shard=shard, index_cls=orc.IndexFlatIP, merge_fn=numpy_merge,
                                result_device=torch.device("cpu"))
        assert ix.metric_type == 0
        ix.add(xb[:250]); ix.add(xb[250:])
        D, I = ix.search(xq, 9)
        D2, I2 = ix.search(xq, 450)                                      # k > ntotal: (-FLT_MAX, -1) padding survives the merge
        np.savez(Path(out_dir) / f"r{rank}.npz", D=D, I=I, D2=D2, I2=I2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["db", "query"])
def test_two_rank_sharded_inner_product_equals_single(tmp_path, shard):
    world, port = 2, _free_port()
    mp.spawn(_worker_ip, args=(world, port, shard, str(tmp_path)), nprocs=world, join=True)
    from oracle import flatl2_oracle as orc
    rng = np.random.default_rng(78)
    xb = rng.integers(-5, 6, size=(401, 12)).astype(np.float32)
    xq = rng.integers(-5, 6, size=(33, 12)).astype(np.float32)
    Dr, Ir = orc.knn_ip_fp32(xq, xb, 9)
    Dr2, Ir2 = orc.knn_ip_fp32(xq, xb, 450)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(got["D"], Dr)
        np.testing.assert_array_equal(got["I"], Ir)       # ties across shards resolve by global id
        np.testing.assert_array_equal(got["D2"], Dr2)
        np.testing.assert_array_equal(got["I2"], Ir2)


def _worker(rank, world, port, shard, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from agplace_b200.sharded import ShardedIndexFlatL2, shard_bounds
        from oracle import flatl2_oracle as orc
        from tests.helpers import numpy_merge
        rng = np.random.default_rng(77)
        xb = rng.integers(-5, 6, size=(501, 16)).astype(np.float32)      # lattice: exact ties across shards
        xq = rng.integers(-5, 6, size=(45, 16)).astype(np.float32)
        ix = ShardedIndexFlatL2(16, shard=shard, index_cls=orc.IndexFlatL2, merge_fn=numpy_merge,
                                result_device=torch.device("cpu"))
        ix.add(xb[:300]); ix.add(xb[300:])                               # two chunks: per-chunk id bases
        assert ix.ntotal == 501
        if shard == "db":
            assert ix.local.ntotal == sum(b - a for a, b in (shard_bounds(300, world)[rank], shard_bounds(201, world)[rank]))
        D, I = ix.search(xq, 12)
        D2, I2 = ix.search(torch.from_numpy(xq), 60)                     # k spanning shards, torch in -> torch out
        if shard == "db":                                                # dst = r: only rank r receives the result (like dist.gather)
            for r in range(world):
                Dg, Ig = ix.search(xq, 12, dst=r)
                if rank == r:
                    assert np.array_equal(Dg, D) and np.array_equal(Ig, I)
                else:
                    assert Dg is None and Ig is None
        np.savez(Path(out_dir) / f"r{rank}.npz", D=D, I=I, D2=D2.numpy(), I2=I2.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["db", "query"])
def test_two_rank_sharded_equals_single(tmp_path, shard):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, shard, str(tmp_path)), nprocs=world, join=True)
    from oracle import flatl2_oracle as orc
    rng = np.random.default_rng(77)
    xb = rng.integers(-5, 6, size=(501, 16)).astype(np.float32)
    xq = rng.integers(-5, 6, size=(45, 16)).astype(np.float32)
    Dr, Ir = orc.knn_fp32(xq, xb, 12)
    Dr2, Ir2 = orc.knn_fp32(xq, xb, 60)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(got["D"], Dr)
        np.testing.assert_array_equal(got["D2"], Dr2)
        np.testing.assert_array_equal(got["I"], Ir)       # ties across shards resolve by global id
        np.testing.assert_array_equal(got["I2"], Ir2)


def test_shard_bounds_cover_everything():
    from agplace_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _worker_interleaved(rank, world, port, out_dir):
    """add / search / add / search / reset / add / search on a local index that honours set_id_base like the engine
    (a stale engine id_base corrupted the global ids after the second add or a reset)."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from agplace_b200.sharded import ShardedIndexFlatL2
        from oracle import flatl2_oracle as orc
        from tests.helpers import numpy_merge

        class BasedOracleIndex(orc.IndexFlatL2):
            id_base = 0

            def set_id_base(self, base):
                self.id_base = int(base)

            def search(self, x, k, **kw):
                D, I = super().search(x, k, **kw)
                return D, np.where(I >= 0, I + self.id_base, I)

        rng = np.random.default_rng(79)
        xb = rng.integers(-5, 6, size=(700, 8)).astype(np.float32)
        xq = rng.integers(-5, 6, size=(30, 8)).astype(np.float32)
        ix = ShardedIndexFlatL2(8, shard="db", index_cls=BasedOracleIndex, merge_fn=numpy_merge, result_device=torch.device("cpu"))
        out = {}
        ix.add(xb[:300])
        out["D1"], out["I1"] = ix.search(xq, 7)            # one chunk per rank: the engine applies the base
        ix.add(xb[300:500])
        out["D2"], out["I2"] = ix.search(xq, 7)            # two non-contiguous chunks on every rank: the base must be cleared
        ix.reset()
        assert ix.local.id_base == 0 and ix.ntotal == 0
        ix.add(xb[500:])
        out["D3"], out["I3"] = ix.search(xq, 7)            # ids restart at 0 after reset
        ix.add(xb[:100])
        out["D4"], out["I4"] = ix.search(xq, 7)
        np.savez(Path(out_dir) / f"r{rank}.npz", **out)
    finally:
        dist.destroy_process_group()


def test_two_rank_add_search_add_search_and_reset(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_interleaved, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import flatl2_oracle as orc
    rng = np.random.default_rng(79)
    xb = rng.integers(-5, 6, size=(700, 8)).astype(np.float32)
    xq = rng.integers(-5, 6, size=(30, 8)).astype(np.float32)
    want = {1: xb[:300], 2: xb[:500], 3: xb[500:], 4: np.concatenate([xb[500:], xb[:100]])}
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        for step, rows in want.items():
            Dr, Ir = orc.knn_fp32(xq, rows, 7)
            np.testing.assert_array_equal(got[f"D{step}"], Dr, err_msg=f"step {step}")
            np.testing.assert_array_equal(got[f"I{step}"], Ir, err_msg=f"step {step}")
