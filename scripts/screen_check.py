"""On-GPU probe of the single-pass certified screen (development aid): parity vs the CPU oracle on a few
shapes, then cfg2-shaped timing with the kernel's debug counters."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, ".")
import agplace_b200 as agp
from oracle import flatl2_oracle as orc


def run(nq, n, d, k, precision="fp16_screen", seed=0, unit=True, dup=0):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    if unit:
        xb /= np.linalg.norm(xb, axis=1, keepdims=True) + 1e-12
        xq /= np.linalg.norm(xq, axis=1, keepdims=True) + 1e-12
    if dup:
        xb[: n // 2] = xb[rng.integers(0, dup, n // 2)]      # heavy duplication: exercises the overflow fallback
    ix = agp.IndexFlatL2(d, precision=precision)
    ix.add(xb)
    t = time.time(); D, I = ix.search(xq, k); dt = time.time() - t
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir)
    same = float((I == Ir).mean())
    print(json.dumps(dict(nq=nq, n=n, d=d, k=k, dup=dup, ok=ok, msg=msg, idx_equal=same, stats=ix.get_stats(), sec=round(dt, 4))), flush=True)
    return ok


def perf(nq=20000, n=100000, d=512, k=50, reps=5):
    import torch
    rng = np.random.default_rng(1)
    xb = rng.standard_normal((n, d)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = rng.standard_normal((nq, d)).astype(np.float32); xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    ix = agp.IndexFlatL2(d, precision="fp16_screen")
    ix.add(xb)
    xq_d = torch.from_numpy(xq).cuda()
    for _ in range(3):
        D, I = ix.search(xq_d, k)
    torch.cuda.synchronize()
    ix.set_profiling(True); ix.get_profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        D, I = ix.search(xq_d, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kms, kn = ix.get_profile(reset=True)
    ix.set_profiling(False)
    print(json.dumps(dict(cfg=f"{n}x{nq}x{d} k={k}", step_ms=round(ms, 4), kernel_ms=round(kms / max(kn, 1), 4),
                          tflops=round(2.0 * nq * n * d / (kms / max(kn, 1) * 1e-3) / 1e12, 1), stats=ix.get_stats(),
                          env={k_: v for k_, v in os.environ.items() if k_.startswith("AGP_")})), flush=True)
    return D, I


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        allok = True
        for (nq, n, d, k) in [(40, 5000, 64, 20), (300, 3000, 256, 50), (130, 1000, 100, 100), (64, 700, 512, 256), (33, 50, 16, 60),
                              (2000, 10000, 256, 20), (1000, 40000, 512, 10), (600, 3000, 1024, 10), (257, 513, 40, 5)]:
            try:
                allok &= run(nq, n, d, k)
            except Exception as e:
                allok = False
                print("EXC", nq, n, d, k, repr(e), flush=True)
        allok &= run(200, 4000, 128, 20, dup=3)
        print("ALL OK" if allok else "FAILURES", flush=True)
    if what in ("all", "perf"):
        perf()
