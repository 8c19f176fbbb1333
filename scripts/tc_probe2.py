"""Development probe: time the fused kernel under env-selected variants (key=value,... per variant on argv)."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, ".")
import torch
import agplace_b200 as agp
from agplace_b200 import synth

def run(cfg, reps=5):
    c = synth.CONFIGS[cfg]
    xb = synth.descriptors(c["n"], c["d"], c["seed"], "db")
    xq = synth.descriptors(c["nq"], c["d"], c["seed"] + 7, "q")
    ix = agp.IndexFlatL2(c["d"], precision=os.environ.get("PREC", "3xtf32"))
    ix.add(xb)
    xq_dev = torch.from_numpy(xq).cuda()
    ix.search(xq_dev, c["k"]); torch.cuda.synchronize()
    ix.set_profiling(True); ix.get_profile()
    t0 = time.perf_counter()
    for _ in range(reps):
        D, I = ix.search(xq_dev, c["k"])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms, n = ix.get_profile()
    flops = 2.0 * c["nq"] * c["n"] * c["d"]
    return dict(cfg=cfg, kernel_ms=round(ms / max(n, 1), 4), step_ms=round(wall, 4), tflops=round(flops / (ms / max(n, 1) * 1e-3) / 1e12, 1)), I.cpu().numpy(), D.cpu().numpy()

if __name__ == "__main__":
    cfg = sys.argv[1]
    ref = None
    for variant in sys.argv[2:]:
        for kv in variant.split(","):
            if "=" in kv:
                k, v = kv.split("="); os.environ[k] = v
        r, I, D = run(cfg)
        r["variant"] = variant
        if ref is None: ref = (I, D)
        r["same_as_first"] = bool((I == ref[0]).all() and (D == ref[1]).all())
        print(json.dumps(r), flush=True)
