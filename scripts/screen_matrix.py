"""Development aid: one process, one resident cfg2-shaped index, a matrix of screen-kernel variants selected through
the AGP_SCREEN_* environment switches (read per search).  Prints kernel/step time per variant, checks every variant's
(D, I) against the first one bit for bit, then one debug-counter pass per variant (stderr)."""
import sys, json, os
import numpy as np
sys.path.insert(0, ".")
import torch
import agplace_b200 as agp

n, nq, d, k = 100000, 20000, 512, 50
if len(sys.argv) > 1:
    n, nq, d, k = (int(a) for a in sys.argv[1].split("x"))
rng = np.random.default_rng(1)
xb = rng.standard_normal((n, d)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
xq = rng.standard_normal((nq, d)).astype(np.float32); xq /= np.linalg.norm(xq, axis=1, keepdims=True)
ix = agp.IndexFlatL2(d, precision="fp16_screen"); ix.add(xb)
xq_d = torch.from_numpy(xq).cuda()

variants = [dict(), dict(FLAGS=1), dict(FLAGS=4), dict(FLAGS=8), dict(E=16), dict(SCHED=6), dict(SCHED=10)]
variants = variants + variants          # second pass: order / warm-up effects show as a difference between the passes
if len(sys.argv) > 2:
    variants = [dict()] + [json.loads(a) for a in sys.argv[2:]]
ref = None
for v in variants:
    for key in ("FLAGS", "E", "SCHED", "STAGES"):
        os.environ.pop("AGP_SCREEN_" + key, None)
    for key, val in v.items():
        os.environ["AGP_SCREEN_" + key] = str(val)
    os.environ.pop("AGP_TC_DEBUG", None)
    for _ in range(15):
        D, I = ix.search(xq_d, k)
    torch.cuda.synchronize()
    ix.set_profiling(True); ix.get_profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        D, I = ix.search(xq_d, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    kms, kn = ix.get_profile(reset=True); ix.set_profiling(False)
    D, I = D.cpu().numpy(), I.cpu().numpy()
    if ref is None:
        ref = (D, I)
    same = bool((I == ref[1]).all() and (D == ref[0]).all())
    print(json.dumps(dict(variant=v, step_ms=round(ms, 4), kernel_ms=round(kms / max(kn, 1), 4),
                          tflops=round(2.0 * nq * n * d / (kms / max(kn, 1) * 1e-3) / 1e12, 1), identical=same,
                          stats=ix.get_stats())), flush=True)
    os.environ["AGP_TC_DEBUG"] = "1"
    sys.stderr.write(f"--- {v}\n"); sys.stderr.flush()
    ix.search(xq_d, k)
    torch.cuda.synchronize()
