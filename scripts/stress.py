"""Randomised API-sequence stress (GPU): add / search / reset sequences with mixed batch sizes (host mirror -> upload ->
growth), mixed k (all three slot budgets), numpy / torch CPU / torch CUDA inputs, one-device and multi-device indexes
(real devices when the box has several, virtual shards otherwise), every result checked against the oracle.  Hunts the
state bugs fixed-shape tests miss (it found the "current device left on a shard" bug of the multi-device index).
    python scripts/stress.py [seconds] [seed]          (tests/test_gpu_stress.py runs a short budget)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")


def run(budget=120.0, seed=0):
    import torch

    import agplace_b200 as agp
    from oracle import flatl2_oracle as orc

    orc.build()
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    n_checks = n_seq = 0

    def make(n, d, regime):
        if regime == "lattice":
            return rng.integers(-4, 5, size=(n, d)).astype(np.float32)
        x = rng.standard_normal((n, d)).astype(np.float32)
        if regime == "unit":
            x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
        return x

    while time.time() < t_end:
        n_seq += 1
        d = int(rng.choice([1, 7, 32, 64, 200, 256, 512, 513, 1024]))
        regime = str(rng.choice(["gauss", "unit", "lattice"]))
        metric = "ip" if rng.random() < 0.15 else "l2"
        multi = rng.random() < 0.3
        cls = agp.IndexFlatIP if metric == "ip" else agp.IndexFlatL2
        n_gpu = torch.cuda.device_count()
        if multi and n_gpu >= 2 and rng.random() < 0.7:       # real devices when the box has them (peer copies over NVLink)
            devs = [int(v) for v in rng.permutation(n_gpu)[: int(rng.integers(2, min(n_gpu, 4) + 1))]]
            devs = [0] + [v for v in devs if v != 0]           # CUDA-tensor inputs of this script live on cuda:0 = the home device
        else:
            devs = [0] * int(rng.integers(2, 5))
        ix = cls(d, devices=devs) if multi else cls(d)
        rows = np.empty((0, d), np.float32)
        desc = f"seq {n_seq}: d={d} {regime} {metric} multi={devs if multi else None}"
        for step in range(int(rng.integers(2, 7))):
            op = rng.random()
            if op < 0.45 or len(rows) == 0:
                n = int(rng.choice([1, 3, 40, 300, 1500, 6000, 20000])) if d <= 256 else int(rng.choice([1, 3, 40, 300, 1500]))
                x = make(n, d, regime)
                how = rng.random()
                if how < 0.6:
                    ix.add(x)
                elif how < 0.8:
                    ix.add(torch.from_numpy(x))
                else:
                    ix.add(torch.from_numpy(x).cuda())
                rows = np.concatenate([rows, x])
                desc += f" | add {n}"
            elif op < 0.93:
                nq = int(rng.choice([1, 2, 19, 20, 33, 257, 700, 3000] + ([5000, 9000] if d <= 256 else [])))
                k = int(rng.choice([1, 5, 10, 50, 77, 100, 256, 300, 512]))
                xq = make(nq, d, regime)
                how = rng.random()
                if how < 0.6:
                    D, I = ix.search(xq, k)
                elif how < 0.75:
                    D, I = ix.search(torch.from_numpy(xq), k)
                    D, I = D.numpy(), I.numpy()
                else:
                    D, I = ix.search(torch.from_numpy(xq).cuda(), k)
                    D, I = D.cpu().numpy(), I.cpu().numpy()
                desc += f" | search nq={nq} k={k}"
                if metric == "ip":
                    Dr, Ir = orc.knn_ip_fp32(xq, rows, k)
                    real = Ir >= 0
                    tol = 1e-4 * np.abs(Dr) + 64 * 2.0 ** -24 * (np.linalg.norm(xq, axis=1)[:, None] * np.linalg.norm(rows, axis=1).max() + 1e-30)
                    assert np.array_equal(I < 0, Ir < 0), desc
                    assert (np.abs(D - Dr)[real] <= tol[real]).all(), desc
                    if regime == "lattice":
                        assert np.array_equal(I, Ir) and np.array_equal(D, Dr), desc
                else:
                    Dr, Ir = orc.knn_fp32(xq, rows, k)
                    if regime == "lattice":
                        assert np.array_equal(I, Ir) and np.array_equal(D, Dr), desc
                    else:
                        ok, msg = orc.compare_knn(D, I, Dr, Ir, xq=xq, xb=rows, abs_floor_eps=32 * 2.0 ** -24)
                        assert ok, desc + " :: " + msg
                n_checks += 1
            else:
                ix.reset()
                rows = np.empty((0, d), np.float32)
                desc += " | reset"
            assert ix.ntotal == len(rows), desc
        del ix
    return n_seq, n_checks


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n_seq, n_checks = run(budget, seed)
    print(f"stress ok: {n_seq} sequences, {n_checks} searches checked in {budget:.0f} s (seed {seed})")
