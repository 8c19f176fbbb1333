"""Development probe: time the fused tcgen05 kernel on a BASELINE config under env-selected variants."""
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
    ix = agp.IndexFlatL2(c["d"], precision="3xtf32")
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
    return dict(cfg=cfg, kernel_ms=ms / max(n, 1), step_ms=wall, tflops=flops / (ms / max(n, 1) * 1e-3) / 1e12,
                bk=os.environ.get("AGP_TC_BK", "32"), skip=os.environ.get("AGP_TC_SKIP_MMA", "0")), I.cpu().numpy()

if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["cfg2"]
    for cfg in cfgs:
        out = {}
        for bk in ("32", "16"):
            for skip in ("0", "1"):
                os.environ["AGP_TC_BK"] = bk; os.environ["AGP_TC_SKIP_MMA"] = skip
                r, I = run(cfg)
                print(json.dumps(r), flush=True)
                if skip == "0": out[bk] = I
        print("bk16 == bk32 indices:", bool((out["16"] == out["32"]).all()))
