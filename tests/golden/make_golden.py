"""Generates tests/golden/*.npz -- committed known-answer vectors for the exact-L2 top-k path.

The AGPlace reference holds NO golden vectors for this path and its arithmetic (faiss-cpu) cannot
be executed in this image ("parity unpinned"), so these fixtures are produced by
the fp64 brute-force ground truth (oracle/flatl2_oracle.py:knn_fp64) on seeded inputs with
well-separated distances, plus hand-constructed cases whose answers are known by construction.
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import flatl2_oracle as orc  # noqa: E402


def save(name, xb, xq, k, D, I, note):
    np.savez_compressed(HERE / f"{name}.npz", xb=xb.astype(np.float32), xq=xq.astype(np.float32), k=np.int64(k),
                        D=D.astype(np.float32), I=I.astype(np.int64), note=np.array(note))


def main():
    # 1. integer lattice: every distance is an exactly representable integer, no rounding anywhere
    rng = np.random.default_rng(1234)
    xb = rng.integers(-8, 9, size=(300, 24)).astype(np.float32)
    xq = rng.integers(-8, 9, size=(37, 24)).astype(np.float32)
    D, I = orc.knn_fp64(xq, xb, 12)
    save("lattice_37x300x24_k12", xb, xq, 12, D, I, "integer coordinates: distances exact in fp32; ties -> lower id")
    # 2. identity rows: query i == database row i -> d = 0 at rank 0, index i
    xb = np.eye(40, dtype=np.float32) * 3.0
    xq = xb[:25].copy()
    D, I = orc.knn_fp64(xq, xb, 5)
    save("identity_25x40x40_k5", xb, xq, 5, D, I, "rank 0 = self at distance 0; remaining ranks tie at 18 -> ascending ids")
    # 3. k > ntotal: padded with (FLT_MAX, -1)
    xb = rng.standard_normal((7, 16)).astype(np.float32)
    xq = rng.standard_normal((21, 16)).astype(np.float32)
    D, I = orc.knn_fp64(xq, xb, 10)
    save("pad_21x7x16_k10", xb, xq, 10, D, I, "k > ntotal: trailing (3.4028235e38, -1)")
    # 4. duplicates: every row appears 3 times -> groups of equal distances ordered by id
    base = rng.integers(-4, 5, size=(50, 8)).astype(np.float32)
    xb = np.concatenate([base, base, base])
    xq = rng.integers(-4, 5, size=(30, 8)).astype(np.float32)
    D, I = orc.knn_fp64(xq, xb, 9)
    save("dups_30x150x8_k9", xb, xq, 9, D, I, "triplicated rows: exact ties resolve to ascending ids")
    # 5. small-batch (nq < 20, faiss difference-form branch) gaussian, well separated (d small, N small)
    xb = rng.standard_normal((64, 4)).astype(np.float32)
    xq = rng.standard_normal((5, 4)).astype(np.float32)
    D, I = orc.knn_fp64(xq, xb, 3)
    save("gauss_5x64x4_k3", xb, xq, 3, D, I, "nq < 20 branch; fp64 truth, compare distances to 1e-5 rel")
    # 6. k = 100 (reservoir handler), lattice
    xb = rng.integers(-6, 7, size=(1000, 32)).astype(np.float32)
    xq = rng.integers(-6, 7, size=(24, 32)).astype(np.float32)
    D, I = orc.knn_fp64(xq, xb, 100)
    save("lattice_24x1000x32_k100", xb, xq, 100, D, I, "k >= 100: reservoir handler; integer distances")


if __name__ == "__main__":
    main()
