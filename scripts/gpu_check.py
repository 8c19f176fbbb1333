"""Quick on-GPU parity probe (development aid): CUDA engine vs the CPU oracle on a few shapes."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import agplace_b200 as agp
from oracle import flatl2_oracle as orc

def run(nq, n, d, k, precision, seed=0, unit=True):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    if unit:
        xb /= np.linalg.norm(xb, axis=1, keepdims=True) + 1e-12
        xq /= np.linalg.norm(xq, axis=1, keepdims=True) + 1e-12
    ix = agp.IndexFlatL2(d, precision=precision)
    ix.add(xb)
    t = time.time(); D, I = ix.search(xq, k); dt = time.time() - t
    Dr, Ir = orc.knn_fp32(xq, xb, k)
    ok, msg = orc.compare_knn(D, I, Dr, Ir)
    same = float((I == Ir).mean())
    pad = Ir < 0
    rel = float(np.max(np.abs(D - Dr)[~pad] / np.maximum(Dr[~pad], 1e-30))) if (~pad).any() else 0.0
    print(json.dumps(dict(nq=nq, n=n, d=d, k=k, precision=precision, ok=ok, msg=msg, idx_equal=same, max_rel=rel, sec=round(dt, 4))), flush=True)
    return ok

if __name__ == "__main__":
    allok = True
    for prec in ("fp32_simt", "exact_diff", "3xtf32", "auto"):
        for (nq, n, d, k) in [(1, 1000, 256, 10), (5, 300, 33, 3), (40, 5000, 64, 20), (200, 3000, 256, 50),
                              (130, 1000, 100, 100), (64, 700, 512, 256), (33, 50, 16, 60)]:
            if prec == "exact_diff" and nq > 64: continue
            try:
                allok &= run(nq, n, d, k, prec)
            except Exception as e:
                allok = False
                print("EXC", prec, nq, n, d, k, repr(e), flush=True)
    print("ALL OK" if allok else "FAILURES")
