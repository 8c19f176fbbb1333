"""Row-sharded index over NCCL on real GPUs (skipped unless >= 2 devices are visible):
ShardedIndexFlatL2 on 2 ranks must return exactly what one unsharded index returns."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shard, out_dir, metric="l2"):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from agplace_b200.sharded import ShardedIndexFlatIP, ShardedIndexFlatL2
        rng = np.random.default_rng(5)
        xb = rng.standard_normal((30011, 128)).astype(np.float32)
        xq = rng.standard_normal((777, 128)).astype(np.float32)
        ix = (ShardedIndexFlatIP if metric == "ip" else ShardedIndexFlatL2)(128, device=rank, shard=shard)
        ix.add(xb[:20000]); ix.add(xb[20000:])
        D, I = ix.search(xq, 40)                                   # numpy in -> numpy out
        Dt, It = ix.search(torch.from_numpy(xq).cuda(), 100)       # CUDA in -> CUDA out
        extra = {}
        if shard == "db":                                          # the chunked host pipeline (H2D | search + all-gather + merge | D2H)
            ix.PIPELINE_MIN_BYTES, ix.PIPELINE_CHUNK = 0, 300
            D3, I3 = ix.search(xq, 40)
            assert np.array_equal(D3, D) and np.array_equal(I3, I)
            D4, I4 = ix.search(torch.from_numpy(xq).cuda(), 40)   # CUDA in, chunked: exchange through peer memory on a second stream
            assert np.array_equal(D4.cpu().numpy(), D) and np.array_equal(I4.cpu().numpy(), I)
            extra["peer_memory"] = np.array([0 if getattr(ix, "_peer", None) in (None, False) else 1])
            for r in range(world):                                 # dst = r: lists travel to rank r alone, only r merges and copies out
                D5, I5 = ix.search(xq, 40, dst=r)                  # host pipeline
                D6, I6 = ix.search(torch.from_numpy(xq).cuda(), 40, dst=r)      # device-resident, chunked
                if rank == r:
                    assert np.array_equal(D5, D) and np.array_equal(I5, I)
                    assert np.array_equal(D6.cpu().numpy(), D) and np.array_equal(I6.cpu().numpy(), I)
                else:
                    assert D5 is None and I5 is None and D6 is None and I6 is None
            ix.PIPELINE_MIN_BYTES, ix.PIPELINE_CHUNK = 16 << 20, 18944
            D7, I7 = ix.search(xq, 40, dst=0)                      # small batch: the generic path (NCCL all-gather, merge on dst only)
            assert (np.array_equal(D7, D) and np.array_equal(I7, I)) if rank == 0 else (D7 is None and I7 is None)
            D8, I8 = ix.search(xq, 40)                             # and everybody again afterwards (slot protocol intact)
            assert np.array_equal(D8, D) and np.array_equal(I8, I)
        if shard == "query":                                       # results left partitioned by query: this rank's slice only
            Dl, Il = ix.search(xq, 40, gather=False)
            extra = dict(Dl=Dl, Il=Il)
        np.savez(Path(out_dir) / f"r{rank}.npz", D=D, I=I, D2=Dt.cpu().numpy(), I2=It.cpu().numpy(), **extra)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["db", "query"])
def test_nccl_sharded_equals_single(tmp_path, shard):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, shard, str(tmp_path)), nprocs=world, join=True)
    import agplace_b200 as agp
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((30011, 128)).astype(np.float32)
    xq = rng.standard_normal((777, 128)).astype(np.float32)
    single = agp.IndexFlatL2(128, device=0)
    single.add(xb)
    Ds, Is = single.search(xq, 40)
    Ds2, Is2 = single.search(xq, 100)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(got["I"], Is)
        np.testing.assert_array_equal(got["D"], Ds)
        np.testing.assert_array_equal(got["I2"], Is2)
        np.testing.assert_array_equal(got["D2"], Ds2)
        if shard == "db" and "peer_memory" in got.files:
            print("peer-memory exchange used on rank", r, ":", bool(got["peer_memory"][0]))
        if shard == "query":
            from agplace_b200.sharded import shard_bounds
            a, b = shard_bounds(len(xq), world)[r]
            np.testing.assert_array_equal(got["Il"], Is[a:b])
            np.testing.assert_array_equal(got["Dl"], Ds[a:b])


def test_nccl_sharded_inner_product_equals_single(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, "db", str(tmp_path), "ip"), nprocs=world, join=True)
    import agplace_b200 as agp
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((30011, 128)).astype(np.float32)
    xq = rng.standard_normal((777, 128)).astype(np.float32)
    single = agp.IndexFlatIP(128, device=0)
    single.add(xb)
    Ds, Is = single.search(xq, 40)
    Ds2, Is2 = single.search(xq, 100)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(got["I"], Is)
        np.testing.assert_array_equal(got["D"], Ds)
        np.testing.assert_array_equal(got["I2"], Is2)
        np.testing.assert_array_equal(got["D2"], Ds2)
