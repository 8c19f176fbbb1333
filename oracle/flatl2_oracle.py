"""oracle/flatl2_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracles for the exact-L2 top-k path that AGPlace runs through
``faiss.IndexFlatL2`` (reference call sites: test.py:27-32,
datasets/datasets_ws_kitti360.py:976-993, datasets/datasets_ws_nuscenes.py:1241-1258,
datasets_ws.py:689-706).  faiss-cpu is an un-pinned, un-vendored third-party wheel
(reference README.md:45) that is absent from this image, so this module *restates*
faiss's published IndexFlatL2 algorithm instead of importing it:

* ``O64``  -- :func:`knn_fp64`: fp64 brute force ``sum((x-y)**2)``, stable order by
  (distance, index).  Ground truth for neighbour sets.
* ``O32``  -- :class:`IndexFlatL2`: fp32 restatement of faiss's two code paths
  (``nq < 20``: exact difference form; ``nq >= 20``: norms + sgemm blocks of
  4096 x 1024 + ``xn + yn - 2 ip`` clamped at 0), Top1 / max-heap (k < 100) /
  reservoir (k >= 100) result handlers, ascending (distance, id) output, (FLT_MAX, -1)
  padding, and the SWIG wrapper's input coercions.  The sgemm is numpy/OpenBLAS (faiss
  delegates to BLAS the same way); norms, epilogue and heaps are the C code in
  ``oracle/flatl2_ref.c`` (``oracle/_build/liboracle.so``).  A slow pure-numpy
  variant (:func:`knn_fp32_numpy`) exists to cross-check the C code.

PARITY PIN: the reference holds no golden vectors or tests at this boundary and faiss cannot be
executed here.  The oracle is pinned against (i) the one known-answer vector real faiss has
published for IndexFlatL2 -- the output of its own first tutorial (tutorial/python/1-Flat.py, faiss
wiki "Getting started": 100k x 64 database, ``np.random.seed(1234)``, k = 4; transcribed into
``tests/golden/faiss_tutorial_1flat.json``): all 40 neighbour ids and the 20 printed distances are
reproduced on both code paths (``tests/test_oracle.py``); (ii) ``O64``; (iii) the hand-written
known-answer vectors in ``tests/golden``.  Everything beyond that single published vector (tie
order, padding, the inner-product variant) is "against the restatement".  If ``import faiss`` ever
succeeds, :func:`faiss_available` reports it and tests compare against it too.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)
BLAS_THRESHOLD = 20        # faiss distance_compute_blas_threshold
BS_X, BS_Y = 4096, 1024    # faiss distance_compute_blas_{query,database}_bs
RESERVOIR_MIN_K = 100      # faiss distance_compute_min_k_reservoir

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/flatl2_ref.c -> oracle/_build/liboracle.so (gcc, OpenMP)."""
    src = _HERE / "flatl2_ref.c"
    if _LIB_PATH.exists() and not force and _LIB_PATH.stat().st_mtime >= src.stat().st_mtime:
        return _LIB_PATH
    _LIB_PATH.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fno-fast-math", "-ffp-contract=off",
           "-fPIC", "-shared", "-fvisibility=hidden", "-o", str(_LIB_PATH), str(src), "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        build()
    lib = ctypes.CDLL(str(_LIB_PATH))
    i64, vp, fp, ip = ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
    lib.oracle_norms_l2sqr.argtypes = [fp, fp, i64, i64]
    lib.oracle_handler_new.argtypes = [i64, i64]
    lib.oracle_handler_new.restype = vp
    lib.oracle_handler_free.argtypes = [vp]
    lib.oracle_add_ip_block.argtypes = [vp, i64, i64, i64, i64, fp, fp, fp]
    lib.oracle_handler_finish.argtypes = [vp, fp, ip]
    lib.oracle_knn_l2sqr_seq.argtypes = [fp, fp, i64, i64, i64, i64, fp, ip]
    lib.oracle_knn_l2sqr_blas_c.argtypes = [fp, fp, i64, i64, i64, i64, fp, ip]
    lib.oracle_num_threads.restype = ctypes.c_int
    _lib = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def faiss_available() -> bool:
    try:
        import faiss  # noqa: F401
        return hasattr(faiss, "IndexFlatL2") and "agplace_b200" not in getattr(faiss, "__file__", "")
    except Exception:
        return False


# --------------------------------------------------------------------------------------
# O64: fp64 ground truth
# --------------------------------------------------------------------------------------
def knn_fp64(xq, xb, k, block=256):
    """Exact brute force in fp64; ascending (distance, index); padded (FLT_MAX, -1)."""
    xq = np.asarray(xq, dtype=np.float64)
    xb = np.asarray(xb, dtype=np.float64)
    nq, n = xq.shape[0], xb.shape[0]
    D = np.full((nq, k), float(FLT_MAX), dtype=np.float64)
    I = np.full((nq, k), -1, dtype=np.int64)
    if n == 0:
        return D, I
    bn = (xb * xb).sum(1)
    for i0 in range(0, nq, block):
        q = xq[i0:i0 + block]
        if xb.shape[1] <= 64 or n * q.shape[0] * xb.shape[1] <= 2e8:
            dist = ((q[:, None, :] - xb[None, :, :]) ** 2).sum(-1)
        else:
            # fp64 expansion is accurate to ~1e-16 * norms: fine as ground truth for fp32 work
            dist = (q * q).sum(1)[:, None] + bn[None, :] - 2.0 * (q @ xb.T)
            np.maximum(dist, 0.0, out=dist)
        m = min(k, n)
        order = np.argsort(dist, axis=1, kind="stable")[:, :m]
        D[i0:i0 + block, :m] = np.take_along_axis(dist, order, 1)
        I[i0:i0 + block, :m] = order
    return D, I


# --------------------------------------------------------------------------------------
# O32: fp32 restatement (numpy sgemm + C handlers)
# --------------------------------------------------------------------------------------
def _finish_numpy(dist_f32, k):
    """Canonical selection from a full fp32 distance matrix (ties -> lower id)."""
    nq, n = dist_f32.shape
    D = np.full((nq, k), FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    m = min(k, n)
    if m:
        order = np.argsort(dist_f32, axis=1, kind="stable")[:, :m]
        D[:, :m] = np.take_along_axis(dist_f32, order, 1)
        I[:, :m] = order
    return D, I


def knn_fp32_numpy(xq, xb, k, path=None):
    """Pure-numpy fp32 restatement (materialises nq x N; small sizes only)."""
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    nq = xq.shape[0]
    if path is None:
        path = "seq" if nq < BLAS_THRESHOLD else "blas"
    if xb.shape[0] == 0:
        return _finish_numpy(np.empty((nq, 0), np.float32), k)
    if path == "seq":
        dist = np.empty((nq, xb.shape[0]), dtype=np.float32)
        for i in range(nq):
            t = xb - xq[i]
            dist[i] = np.einsum("ij,ij->i", t, t, dtype=np.float32)
    else:
        xn = np.einsum("ij,ij->i", xq, xq, dtype=np.float32)
        yn = np.einsum("ij,ij->i", xb, xb, dtype=np.float32)
        dist = (xn[:, None] + yn[None, :]) - np.float32(2) * (xq @ xb.T)
        np.maximum(dist, np.float32(0), out=dist)
    return _finish_numpy(dist, k)


def knn_fp32(xq, xb, k, path=None):
    """faiss ``knn_L2sqr`` restated: returns (D f32 [nq,k], I i64 [nq,k])."""
    lib = _load()
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    nq, d = xq.shape
    n = xb.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if nq == 0:
        return D, I
    if path is None:
        path = "seq" if nq < BLAS_THRESHOLD else "blas"
    if path == "seq":
        lib.oracle_knn_l2sqr_seq(_fp(xq), _fp(xb), d, nq, n, k, _fp(D), _ip(I))
        return D, I
    if path == "blas_c":
        lib.oracle_knn_l2sqr_blas_c(_fp(xq), _fp(xb), d, nq, n, k, _fp(D), _ip(I))
        return D, I
    xn = np.empty(nq, dtype=np.float32)
    yn = np.empty(max(n, 1), dtype=np.float32)
    lib.oracle_norms_l2sqr(_fp(xn), _fp(xq), d, nq)
    lib.oracle_norms_l2sqr(_fp(yn), _fp(xb), d, n)
    h = lib.oracle_handler_new(nq, k)
    try:
        ip_buf = np.empty(BS_X * BS_Y, dtype=np.float32)
        for i0 in range(0, nq, BS_X):
            i1 = min(i0 + BS_X, nq)
            for j0 in range(0, n, BS_Y):
                j1 = min(j0 + BS_Y, n)
                ip = ip_buf[: (i1 - i0) * (j1 - j0)].reshape(i1 - i0, j1 - j0)
                np.matmul(xq[i0:i1], xb[j0:j1].T, out=ip)     # sgemm (OpenBLAS)
                lib.oracle_add_ip_block(h, i0, i1, j0, j1, _fp(ip), _fp(xn), _fp(yn))
        lib.oracle_handler_finish(h, _fp(D), _ip(I))
    finally:
        lib.oracle_handler_free(h)
    return D, I


def knn_ip_fp64(xq, xb, k):
    """Exact maximum inner product in fp64; descending product, ties by index; padded (-FLT_MAX, -1)."""
    xq = np.asarray(xq, dtype=np.float64)
    xb = np.asarray(xb, dtype=np.float64)
    nq, n = xq.shape[0], xb.shape[0]
    D = np.full((nq, k), -float(FLT_MAX), dtype=np.float64)
    I = np.full((nq, k), -1, dtype=np.int64)
    m = min(k, n)
    if m:
        ip = xq @ xb.T
        order = np.argsort(-ip, axis=1, kind="stable")[:, :m]
        D[:, :m] = np.take_along_axis(ip, order, 1)
        I[:, :m] = order
    return D, I


def knn_ip_fp32(xq, xb, k, path=None):
    """faiss ``knn_inner_product`` restated (``IndexFlatIP.search``; reference anyloc/utilities.py:446,457): fp32
    ``fvec_inner_product`` per pair for ``nq < 20``, sgemm blocks of 4096 x 1024 otherwise, the k LARGEST products
    per query (min-heap result handler, strict admission: on boundary ties the lower id stays), output descending,
    padded (-FLT_MAX, -1).  Ties inside the list are ordered by ascending id here; faiss's own order of exactly
    equal products cannot be checked in this image (no published vector covers the inner-product metric: restatement only)."""
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    nq, n = xq.shape[0], xb.shape[0]
    D = np.full((nq, k), -FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    m = min(k, n)
    if nq == 0 or m == 0:
        return D, I
    if path is None:
        path = "seq" if nq < BLAS_THRESHOLD else "blas"
    for i0 in range(0, nq, BS_X):
        q = xq[i0:i0 + BS_X]
        if path == "seq":
            ip = np.stack([np.einsum("ij,j->i", xb, qi, dtype=np.float32) for qi in q])
        else:
            ip = np.concatenate([q @ xb[j0:j0 + BS_Y].T for j0 in range(0, n, BS_Y)], axis=1)
        order = np.argsort(-ip, axis=1, kind="stable")[:, :m]
        D[i0:i0 + BS_X, :m] = np.take_along_axis(ip, order, 1)
        I[i0:i0 + BS_X, :m] = order
    return D, I


class IndexFlatL2:
    """Restatement of the SWIG-wrapped ``faiss.IndexFlatL2`` surface the reference uses."""

    def __init__(self, d):
        self.d = int(d)
        self.ntotal = 0
        self.is_trained = True
        self.metric_type = 1  # faiss.METRIC_L2
        self._chunks = []
        self._xb = np.empty((0, self.d), dtype=np.float32)

    _knn = staticmethod(lambda x, xb, k: knn_fp32(x, xb, k))

    def add(self, x):
        n, d = x.shape
        assert d == self.d
        x = np.ascontiguousarray(x, dtype="float32")
        self._chunks.append(x.copy())
        self.ntotal += n
        self._xb = None

    def reset(self):
        self._chunks = []
        self._xb = np.empty((0, self.d), dtype=np.float32)
        self.ntotal = 0

    def _base(self):
        if self._xb is None:
            self._xb = np.concatenate(self._chunks, 0) if self._chunks else np.empty((0, self.d), np.float32)
            self._chunks = [self._xb]
        return self._xb

    def search(self, x, k, *, params=None, D=None, I=None):
        n, d = x.shape
        x = np.ascontiguousarray(x, dtype="float32")
        assert d == self.d
        assert k > 0
        Dn, In = self._knn(x, self._base(), int(k))
        if D is None:
            D = Dn
        else:
            assert D.shape == (n, k)
            D[...] = Dn
        if I is None:
            I = In
        else:
            assert I.shape == (n, k)
            I[...] = In
        return D, I


class IndexFlatIP(IndexFlatL2):
    """Restatement of ``faiss.IndexFlatIP`` (reference anyloc/utilities.py:446): same surface, :func:`knn_ip_fp32`."""

    _knn = staticmethod(lambda x, xb, k: knn_ip_fp32(x, xb, k))

    def __init__(self, d):
        super().__init__(d)
        self.metric_type = 0  # faiss.METRIC_INNER_PRODUCT


# --------------------------------------------------------------------------------------
# comparison helpers shared by the parity tests (tolerances from BASELINE.json north_star)
# --------------------------------------------------------------------------------------
def compare_knn(D, I, D_ref, I_ref, xq=None, xb=None, rel_d=1e-4, rel_tie=1e-5, abs_floor_eps=0.0):
    """Return (ok, message).  Index mismatches are tolerated only where the two rows'
    reference distances at the differing ranks are ties within ``rel_tie`` (relative);
    distances must agree within ``rel_d`` relative (+ ``abs_floor_eps`` * (|q|^2+|x|^2)
    when an absolute floor for the cancellation regime is requested)."""
    D = np.asarray(D); I = np.asarray(I); D_ref = np.asarray(D_ref); I_ref = np.asarray(I_ref)
    if D.shape != D_ref.shape or I.shape != I_ref.shape:
        return False, f"shape mismatch {D.shape} vs {D_ref.shape}"
    floor = 0.0
    if abs_floor_eps and xq is not None and xb is not None:
        qn = (np.asarray(xq, np.float64) ** 2).sum(1)
        bmax = float((np.asarray(xb, np.float64) ** 2).sum(1).max()) if len(xb) else 0.0
        floor = abs_floor_eps * (qn[:, None] + bmax)
    pad = I_ref < 0
    if not np.array_equal(pad, I < 0):
        return False, "padding (-1) positions differ"
    dd = np.abs(D.astype(np.float64) - D_ref.astype(np.float64))
    tol = rel_d * np.abs(D_ref.astype(np.float64)) + floor
    bad = (dd > tol) & ~pad
    if bad.any():
        q, r = np.argwhere(bad)[0]
        return False, f"distance mismatch at q={q} rank={r}: {D[q, r]!r} vs {D_ref[q, r]!r}"
    neq = (I != I_ref) & ~pad
    if neq.any():
        # a mismatching rank is acceptable only if the item we returned is a tie (in the
        # reference's own distances or ours) with the item the reference returned
        for q, r in np.argwhere(neq):
            a, b = float(D[q, r]), float(D_ref[q, r])
            tie_tol = rel_tie * max(abs(a), abs(b)) + (float(np.max(floor[q])) if np.ndim(floor) else floor)
            if abs(a - b) > tie_tol:
                return False, f"index mismatch at q={q} rank={r}: {I[q, r]} vs {I_ref[q, r]} (d {a} vs {b})"
            # and the returned id must exist in the reference list or be a boundary tie
            if I[q, r] not in I_ref[q]:
                kth = float(D_ref[q][~pad[q]][-1])
                if abs(a - kth) > rel_tie * max(abs(a), abs(kth)) + (float(np.max(floor[q])) if np.ndim(floor) else floor):
                    return False, f"id {I[q, r]} at q={q} rank={r} is not a neighbour nor a boundary tie"
    return True, "ok"
