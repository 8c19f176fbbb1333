"""Throughput of the screen path over query-batch sizes (development aid; exercises the partial-wave scheduler)."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
import torch
import agplace_b200 as agp

n, d, k = 100000, 512, 50
rng = np.random.default_rng(1)
xb = rng.standard_normal((n, d)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
ix = agp.IndexFlatL2(d); ix.add(xb)
for nq in [int(a) for a in sys.argv[1:]] or [1000, 2500, 5000, 10000, 20000, 40000]:
    xq = rng.standard_normal((nq, d)).astype(np.float32); xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    xq_d = torch.from_numpy(xq).cuda()
    for _ in range(3):
        ix.search(xq_d, k)
    torch.cuda.synchronize()
    ix.set_profiling(True); ix.get_profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ix.search(xq_d, k)
    e1.record(); torch.cuda.synchronize()
    kms, kn = ix.get_profile(reset=True); ix.set_profiling(False)
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps(dict(nq=nq, step_ms=round(ms, 4), kernel_ms=round(kms / max(kn, 1), 4), mqps=round(nq / ms / 1e3, 2),
                          tflops=round(2.0 * nq * n * d / (kms / max(kn, 1) * 1e-3) / 1e12, 1))), flush=True)
