"""Mining refresh wall-clock (development aid): the reference's per-query loop (two fresh indexes + two searches per
query, datasets/datasets_ws_kitti360.py:1056-1137) on the CUDA engine and on the CPU oracle port, vs the batched N2 path.
Shape: the reference's defaults -- 256-d descriptors, cache_refresh_rate = 1000 queries, neg_samples_num = 1000."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from agplace_b200 import mining
from oracle import flatl2_oracle as orc
from tests.helpers import make_mining_problem

p = make_mining_problem(7, database_num=20000, queries_num=4000, d=256)
R = 1000
res = {}
for name, kw, meth in (("loop_gpu_engine", {}, "compute_triplets_partial"), ("loop_cpu_oracle", dict(index_cls=orc.IndexFlatL2), "compute_triplets_partial"),
                       ("batched_gpu_engine", {}, "compute_triplets_partial_batched")):
    miner = mining.TripletMiner(p.d, p.database_num, p.queries_num, p.hard, p.soft, negs_num_per_query=10, neg_samples_num=1000, **kw)
    best = None
    for rep in range(3):
        np.random.seed(5)
        t0 = time.perf_counter()
        t = getattr(miner, meth)(p.cache, R)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res[name] = dict(seconds=round(best, 4), us_per_query=round(best / R * 1e6, 1), checksum=int(t.sum()))
# one index lifecycle of the reference's two mining calls: IndexFlatL2(d); add(rows); search(1 query, k); drop
import agplace_b200 as agp
rng = np.random.default_rng(0)
for name, rows, k in (("lifecycle_1000x256_k10", 1000, 10), ("lifecycle_3x256_k1", 3, 1)):
    xb = rng.standard_normal((rows, 256)).astype(np.float32)
    q = rng.standard_normal((1, 256)).astype(np.float32)
    for cls, tag in ((agp.IndexFlatL2, "gpu_engine"), (orc.IndexFlatL2, "cpu_oracle")):
        for _ in range(50):
            ix = cls(256); ix.add(xb); ix.search(q, k)
        t0 = time.perf_counter()
        for _ in range(500):
            ix = cls(256); ix.add(xb); ix.search(q, k)
        res[f"{name}_{tag}_us"] = round((time.perf_counter() - t0) / 500 * 1e6, 1)
res["identical"] = len({v["checksum"] for v in res.values() if isinstance(v, dict)}) == 1
print(json.dumps(res))
