"""Development aid: cfg2 through the numpy (pageable host memory) call path the reference uses, next to the pinned /
device-resident paths bench.py reports."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
import agplace_b200 as agp

n, nq, d, k = 100000, 20000, 512, 50
rng = np.random.default_rng(1)
xb = rng.standard_normal((n, d)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
xq = rng.standard_normal((nq, d)).astype(np.float32); xq /= np.linalg.norm(xq, axis=1, keepdims=True)
ix = agp.IndexFlatL2(d); ix.add(xb)
xq_pin = torch.from_numpy(xq).pin_memory()
xq_dev = xq_pin.cuda()
out = {}
for name, fn in [("numpy pageable in, numpy out", lambda: ix.search(xq, k)),
                 ("torch pinned in, torch CPU out", lambda: ix.search(xq_pin, k)),
                 ("CUDA in, CUDA out", lambda: ix.search(xq_dev, k))]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    out[name] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
print(json.dumps(out))
