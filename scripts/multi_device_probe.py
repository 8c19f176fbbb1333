"""Development aid (GPU box with N devices): cfg4 through the single-process multi-device index
(IndexFlatL2(d, devices=[0..N-1])) -- device-resident step and numpy e2e, verified on a sample."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import agplace_b200 as agp
import bench

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
c = bench.workload(sys.argv[2] if len(sys.argv) > 2 else "cfg4")
devices = list(range(n_dev))
ix = agp.IndexFlatL2(c["d"], devices=devices) if n_dev > 1 else agp.IndexFlatL2(c["d"], device=0)
t0 = time.perf_counter()
step_rows = 1 << 18
from agplace_b200 import synth
for a in range(0, c["n"], step_rows):
    b = min(c["n"], a + step_rows)
    ix.add(synth.counter_rows_device(a, b, c["d"], c["seed"], torch.device("cuda", 0)))
torch.cuda.synchronize()
add_s = time.perf_counter() - t0
xq = bench.host_queries(c)
xq_d = torch.from_numpy(xq).to("cuda:0")
k = c["k"]
D, I = ix.search(xq_d, k)
ver = bench.verify_sample(c, xq, D.cpu().numpy(), I.cpu().numpy())
for _ in range(2):
    ix.search(xq_d, k)
torch.cuda.synchronize()
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
    ix.search(xq_d, k)
torch.cuda.synchronize()
dev_ms = (time.perf_counter() - t0) / reps * 1e3
ix.search(xq, k)
t0 = time.perf_counter()
for _ in range(reps):
    De, Ie = ix.search(xq, k)
e2e_ms = (time.perf_counter() - t0) / reps * 1e3
same = bool(np.array_equal(Ie, I.cpu().numpy()))
print(json.dumps(dict(workload=c["name"], devices=devices, single_process=True, add_s=round(add_s, 1), device_step_ms=round(dev_ms, 2),
                      device_qps=round(c["nq"] / dev_ms * 1e3), numpy_e2e_ms=round(e2e_ms, 2), e2e_qps=round(c["nq"] / e2e_ms * 1e3),
                      numpy_equals_device=same, verify=ver)))
