"""Development aid (GPU): build a config's index and run a few device-resident searches (target of ncu captures).
    python scripts/one_search.py cfg2 [n_rows] [n_searches]"""
import sys
sys.path.insert(0, ".")
import torch
import agplace_b200 as agp
import bench

c = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
if len(sys.argv) > 2 and int(sys.argv[2]) > 0:
    c["n"] = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
ix = agp.IndexFlatL2(c["d"], device=0)
bench.fill_index(ix, ix, c, 0, c["n"], dev, 1, "single")
xq = torch.from_numpy(bench.host_queries(c)).to(dev)
for _ in range(reps):
    D, I = ix.search(xq, c["k"])
torch.cuda.synchronize()
print("ok", c["name"], c["n"], tuple(D.shape))
