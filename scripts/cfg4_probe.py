"""Development aid (GPU): cfg4 on one GPU -- device step time for several query-chunk sizes of the screen launch."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import agplace_b200 as agp
import bench

c = bench.workload("cfg4")
if len(sys.argv) > 1:
    c["n"] = int(sys.argv[1])
dev = torch.device("cuda", 0)
ix = agp.IndexFlatL2(c["d"], device=0)
bench.fill_index(ix, ix, c, 0, c["n"], dev, 1, "single")
xq = bench.host_queries(c)
xq_d = torch.from_numpy(xq).to(dev)
k = c["k"]
for chunk, lock in ((0, 0), (0, 16), (0, 64), (0, 256), (0, 1024), (65536, 0), (65536, 64), (9472, 0)):
    ix.set_knob("screen_chunk", chunk)
    ix.set_knob("screen_lockstep", lock)
    ix.search(xq_d, k); torch.cuda.synchronize()
    s = bench.ClockSampler(0); s.start()
    ix.set_profiling(True); ix.get_profile_phases(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ix.search(xq_d, k)
    e1.record(); torch.cuda.synchronize()
    ph = ix.get_profile_phases(reset=True); ix.set_profiling(False)
    print(json.dumps(dict(screen_chunk=chunk, lockstep=lock, step_ms=round(e0.elapsed_time(e1) / 3, 2), phases={p: (round(v[0] / 3, 2), v[1]) for p, v in ph.items()}, clocks=s.stop())), flush=True)
ix.set_knob("screen_chunk", 0); ix.set_knob("screen_lockstep", 0)
ix.search(xq, k)
t0 = time.perf_counter()
for _ in range(3):
    ix.search(xq, k)
torch.cuda.synchronize()
print(json.dumps(dict(numpy_e2e_ms=round((time.perf_counter() - t0) / 3 * 1e3, 2))))
