"""Drop-in ``IndexFlatL2`` over libagpknn.so.

Mirrors the SWIG-wrapped ``faiss.IndexFlatL2`` surface that AGPlace uses
(reference test.py:27-32; datasets/datasets_ws_kitti360.py:976-993;
datasets/datasets_ws_nuscenes.py:1241-1258; datasets_ws.py:689-706):

    index = IndexFlatL2(d); index.add(xb); D, I = index.search(xq, k); index.reset()

Same coercions as faiss's Python wrapper (``np.ascontiguousarray(x, 'float32')``, ``assert d ==
self.d``, ``assert k > 0``, optional preallocated ``D=``/``I=``), same return layout (``D`` fp32
[nq, k] squared L2 ascending, ``I`` int64 [nq, k], padded ``(3.4028235e38, -1)``), and the index
copies on ``add``.  ``torch.Tensor`` inputs are accepted like ``faiss.contrib.torch_utils``
(reference anyloc/utilities.py:14,455-456): CUDA tensors stay on the device, run on torch's
current stream and return CUDA tensors.  All arithmetic happens in the CUDA library; there is no
CPU path.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
FLT_MAX = np.float32(3.4028234663852886e38)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def default_device() -> int:
    """AGP_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    for key in ("AGP_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


_PINNED_RESULT_MIN = 1 << 20        # bytes: below this a plain numpy array is cheaper than a pinned block
_PINNED_RESULT_MAX = 1 << 30


def _result_array(shape, dtype):
    """A fresh result array owned by the caller, like faiss returns.  Large results are numpy views of page-locked
    blocks from torch's caching host allocator (the block goes back to that cache when the array is garbage
    collected): the engine's D2H copy then lands directly in the array the caller receives -- no pinned staging copy,
    no first-touch page faults on a fresh pageable allocation."""
    nbytes = shape[0] * shape[1] * np.dtype(dtype).itemsize
    if _PINNED_RESULT_MIN <= nbytes <= _PINNED_RESULT_MAX:
        try:
            import torch
            t = torch.empty(shape, dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.int64, pin_memory=True)
            return t.numpy()          # keeps the tensor (and its pinned block) alive through .base
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


class IndexFlatL2:
    """Exact brute-force squared-L2 index resident in one B200's HBM."""

    _metric = METRIC_L2

    def __init__(self, d, device=None, precision=None, devices=None):
        """``devices=[0, 1, ...]`` (or ``AGP_DEVICES=0,1,...`` in the environment, read here, not in the library) builds ONE
        index over several GPUs of the box from this single process: rows are split across the devices, every search
        runs on all of them and is merged on ``devices[0]`` -- same results as a one-device index, same surface, so the
        reference's ``faiss.IndexFlatL2(d)`` call sites use the whole box unmodified."""
        self.d = int(d)
        self.is_trained = True
        self.metric_type = self._metric
        if devices is None and device is None and os.environ.get("AGP_DEVICES", "") != "":
            devices = [int(v) for v in os.environ["AGP_DEVICES"].split(",") if v.strip() != ""]
        self.devices = [int(v) for v in devices] if devices is not None and len(devices) > 1 else None
        if devices is not None and len(devices) == 1 and device is None:
            device = int(devices[0])
        self.device = self.devices[0] if self.devices else (default_device() if device is None else int(device))
        precision = precision or os.environ.get("AGP_PRECISION", "auto")
        if precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}, got {precision!r}")
        self.precision = precision
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self._foreign_stream = False
        if self.devices:
            ids = (ctypes.c_int * len(self.devices))(*self.devices)
            _lib.check(self._lib.agp_index_create_multi(self.d, len(self.devices), ids, _lib.PRECISION[precision], self._metric,
                                                        ctypes.byref(self._h)), "agp_index_create_multi")
        else:
            _lib.check(self._lib.agp_index_create_metric(self.d, self.device, _lib.PRECISION[precision], self._metric,
                                                         ctypes.byref(self._h)), "agp_index_create_metric")

    # ------------------------------------------------------------------ faiss attributes
    @property
    def ntotal(self) -> int:
        return int(self._lib.agp_index_ntotal(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.agp_index_free(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    # ------------------------------------------------------------------ helpers
    def _use_torch_stream(self, tensor):
        import torch
        if tensor.device.index != self.device:
            raise ValueError(f"tensor is on cuda:{tensor.device.index} but the index lives on cuda:{self.device}")
        stream = torch.cuda.current_stream(tensor.device).cuda_stream
        _lib.check(self._lib.agp_index_set_stream(self._h, ctypes.c_void_p(stream), 0), "agp_index_set_stream")
        self._foreign_stream = True

    def _use_own_stream(self):
        if self._foreign_stream:      # (a fresh index is already on its own stream: the mining loop never pays this call)
            _lib.check(self._lib.agp_index_set_stream(self._h, None, 1), "agp_index_set_stream")
            self._foreign_stream = False

    # ------------------------------------------------------------------ add / reset
    def add(self, x):
        n, d = x.shape
        assert d == self.d
        if _is_torch(x) and x.is_cuda:
            import torch
            x = x.detach().to(torch.float32).contiguous()
            self._use_torch_stream(x)
            _lib.check(self._lib.agp_index_add(self._h, n, ctypes.c_void_p(x.data_ptr()), _lib.MEM_DEVICE), "agp_index_add")
            # x may be a temporary: the copy is queued on torch's stream, which orders it before any reuse
            x.record_stream(__import__("torch").cuda.current_stream(x.device))
            return
        if _is_torch(x):
            x = x.detach().numpy()
        x = np.ascontiguousarray(x, dtype="float32")
        self._use_own_stream()
        _lib.check(self._lib.agp_index_add(self._h, n, ctypes.c_void_p(x.ctypes.data), _lib.MEM_HOST), "agp_index_add")

    def reset(self):
        _lib.check(self._lib.agp_index_reset(self._h), "agp_index_reset")

    def reserve(self, n):
        _lib.check(self._lib.agp_index_reserve(self._h, int(n)), "agp_index_reserve")

    def set_id_base(self, base):
        _lib.check(self._lib.agp_index_set_id_base(self._h, int(base)), "agp_index_set_id_base")

    # ------------------------------------------------------------------ search
    def search(self, x, k, *, params=None, D=None, I=None):
        n, d = x.shape
        assert d == self.d
        assert k > 0
        k = int(k)
        if k > _lib.MAX_K:
            raise RuntimeError(f"k={k} exceeds the engine's maximum of {_lib.MAX_K}")
        if _is_torch(x) and x.is_cuda:
            return self._search_cuda(x, n, k, D, I)
        as_torch = _is_torch(x)
        if as_torch:
            x = x.detach().numpy()
        x = np.ascontiguousarray(x, dtype="float32")
        if D is None:
            Dn = _result_array((n, k), np.float32)
        else:
            Dn = D.numpy() if _is_torch(D) else D
            assert Dn.shape == (n, k)
        if I is None:
            In = _result_array((n, k), np.int64)
        else:
            In = I.numpy() if _is_torch(I) else I
            assert In.shape == (n, k)
        direct = (Dn.dtype == np.float32 and Dn.flags.c_contiguous and In.dtype == np.int64 and In.flags.c_contiguous)
        Dw = Dn if direct else np.empty((n, k), dtype=np.float32)
        Iw = In if direct else np.empty((n, k), dtype=np.int64)
        self._use_own_stream()
        _lib.check(self._lib.agp_index_search(self._h, n, ctypes.c_void_p(x.ctypes.data), _lib.MEM_HOST, k,
                                              ctypes.c_void_p(Dw.ctypes.data), ctypes.c_void_p(Iw.ctypes.data), _lib.MEM_HOST),
                   "agp_index_search")
        if not direct:
            Dn[...] = Dw
            In[...] = Iw
        if as_torch:
            import torch
            return (D if D is not None else torch.from_numpy(Dn)), (I if I is not None else torch.from_numpy(In))
        return (D if D is not None else Dn), (I if I is not None else In)

    def _search_cuda(self, x, n, k, D, I):
        import torch
        x = x.detach().to(torch.float32).contiguous()
        self._use_torch_stream(x)
        if D is None:
            D = torch.empty((n, k), dtype=torch.float32, device=x.device)
        else:
            assert tuple(D.shape) == (n, k) and D.is_cuda and D.dtype == torch.float32 and D.is_contiguous()
        if I is None:
            I = torch.empty((n, k), dtype=torch.int64, device=x.device)
        else:
            assert tuple(I.shape) == (n, k) and I.is_cuda and I.dtype == torch.int64 and I.is_contiguous()
        _lib.check(self._lib.agp_index_search(self._h, n, ctypes.c_void_p(x.data_ptr()), _lib.MEM_DEVICE, k,
                                              ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), _lib.MEM_DEVICE),
                   "agp_index_search")
        x.record_stream(torch.cuda.current_stream(x.device))
        return D, I

    def search_masked(self, x, k, exclude):
        """Batched masked search.  ``exclude[q]`` = row ids of this index that query q must not
        return -- the reference's ``np.setdiff1d(sampled_database_indexes, soft_positives)`` followed by a
        fresh ``IndexFlatL2`` per query (datasets/datasets_ws_kitti360.py:1088-1091, 985-993), for all queries
        in one call.  Returns numpy ``(D fp32 [nq,k], I int64 [nq,k])`` with ids referring to THIS index."""
        n, d = x.shape
        assert d == self.d
        assert k > 0
        assert len(exclude) == n
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype="float32")
        offsets, ids = positives_to_csr(exclude)
        D = np.empty((n, int(k)), dtype=np.float32)
        I = np.empty((n, int(k)), dtype=np.int64)
        self._use_own_stream()
        _lib.check(self._lib.agp_index_search_masked(self._h, n, ctypes.c_void_p(x.ctypes.data), _lib.MEM_HOST, int(k),
                                                     ctypes.c_void_p(offsets.ctypes.data), ctypes.c_void_p(ids.ctypes.data),
                                                     ctypes.c_void_p(D.ctypes.data), ctypes.c_void_p(I.ctypes.data), _lib.MEM_HOST),
                   "agp_index_search_masked")
        return D, I

    def search_subset(self, x, k, candidates):
        """Batched search over per-query candidate subsets (``compute_triplets_full``).
        ``candidates[q]`` = row ids of THIS index that query q may return -- the reference builds a fresh
        ``IndexFlatL2`` over ``cache[neg_indexes]`` per query (datasets/datasets_ws_kitti360.py:985-993 called from
        :1041).  Returns numpy ``(D fp32 [nq,k], I int64 [nq,k])`` where ``I`` holds POSITIONS inside
        ``candidates[q]`` (exactly what the fresh index returns), ties by position, padded ``(3.4028235e38, -1)``."""
        n, d = x.shape
        assert d == self.d
        assert k > 0
        assert len(candidates) == n
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype="float32")
        offsets, ids = positives_to_csr(candidates)
        D = np.empty((n, int(k)), dtype=np.float32)
        I = np.empty((n, int(k)), dtype=np.int64)
        self._use_own_stream()
        _lib.check(self._lib.agp_index_search_subset(self._h, n, ctypes.c_void_p(x.ctypes.data), _lib.MEM_HOST, int(k),
                                                     ctypes.c_void_p(offsets.ctypes.data), ctypes.c_void_p(ids.ctypes.data),
                                                     ctypes.c_void_p(D.ctypes.data), ctypes.c_void_p(I.ctypes.data), _lib.MEM_HOST),
                   "agp_index_search_subset")
        return D, I

    # ------------------------------------------------------------------ instrumentation (bench.py)
    def set_profiling(self, enable: bool):
        _lib.check(self._lib.agp_index_set_profiling(self._h, int(bool(enable))), "agp_index_set_profiling")

    def get_profile(self, reset=True):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self._lib.agp_index_get_profile(self._h, ctypes.byref(ms), ctypes.byref(n), int(reset)), "agp_index_get_profile")
        return ms.value, n.value


    PHASES = ("distance", "prep", "finish", "fallback")

    def get_profile_phases(self, reset=True):
        """{phase: (milliseconds, launches)} accumulated since the last reset (``agp_index_get_profile_phases``)."""
        ms = (ctypes.c_double * 4)()
        n = (ctypes.c_int64 * 4)()
        _lib.check(self._lib.agp_index_get_profile_phases(self._h, ms, n, int(reset)), "agp_index_get_profile_phases")
        return {name: (ms[i], n[i]) for i, name in enumerate(self.PHASES)}

    def set_knob(self, name: str, value: int):
        """Development switch of this index (A/B variants of the screen kernel; ``include/agpknn.h:agp_index_set_knob``)."""
        _lib.check(self._lib.agp_index_set_knob(self._h, name.encode(), int(value)), "agp_index_set_knob")

    def screen_probe(self, x):
        """Diagnostics: the screened distance the tensor-core kernel evaluates for every (query, row) and the certified
        half band per query -- ``(dis~ fp32 [nq, ntotal], band fp32 [nq])``; see ``agp_index_screen_probe``."""
        x = np.ascontiguousarray(x, dtype="float32")
        n, d = x.shape
        assert d == self.d
        dis = np.empty((n, self.ntotal), dtype=np.float32)
        band = np.empty(n, dtype=np.float32)
        self._use_own_stream()
        _lib.check(self._lib.agp_index_screen_probe(self._h, n, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(dis.ctypes.data),
                                                    ctypes.c_void_p(band.ctypes.data)), "agp_index_screen_probe")
        return dis, band

    def get_stats(self):
        """(queries answered by the single-pass screen, queries re-run through the exact fp32 fallback)."""
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._lib.agp_index_get_stats(self._h, ctypes.byref(a), ctypes.byref(b)), "agp_index_get_stats")
        return a.value, b.value


class IndexFlatIP(IndexFlatL2):
    """Exact maximum-inner-product index (``faiss.IndexFlatIP``; reference anyloc/utilities.py:446, the cosine branch
    of ``get_top_k_recall`` -- SURVEY 8f N3).  Same surface as :class:`IndexFlatL2`; ``search`` returns the k LARGEST
    inner products, descending, ties by id, padded ``(-3.4028235e38, -1)``.  ``search_masked`` / ``search_subset`` are
    L2-only (the mining helpers never use an inner-product index)."""

    _metric = METRIC_INNER_PRODUCT


def IndexFlat(d, metric=METRIC_L2, **kw):
    """``faiss.IndexFlat(d, metric)``."""
    if metric == METRIC_L2:
        return IndexFlatL2(d, **kw)
    if metric == METRIC_INNER_PRODUCT:
        return IndexFlatIP(d, **kw)
    raise ValueError(f"unsupported metric {metric!r}")


class StandardGpuResources:
    """Placeholder for ``faiss.StandardGpuResources()`` (reference anyloc/utilities.py:452): the engine owns its
    device memory and streams, there is nothing to configure."""


def index_cpu_to_gpu(res, device, index):
    """``faiss.index_cpu_to_gpu(res, device, index)`` (reference anyloc/utilities.py:453).  Every index of this
    package already lives on a GPU: an empty index is re-created on ``device`` if it was built for another one,
    a populated index must already be there."""
    if index.device == int(device):
        return index
    if index.ntotal:
        raise RuntimeError(f"index holds {index.ntotal} vectors on cuda:{index.device}; create it on cuda:{device} instead")
    return type(index)(index.d, device=int(device), precision=index.precision)


def positives_to_csr(positives_per_query):
    """Reference format (object array / list of unsorted int arrays, test.py:73) -> CSR int64."""
    lens = np.fromiter((len(p) for p in positives_per_query), dtype=np.int64, count=len(positives_per_query))
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    ids = (np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1) for p in positives_per_query])
           if len(lens) and offsets[-1] > 0 else np.empty(0, dtype=np.int64))
    return offsets, np.ascontiguousarray(ids)


def best_of_lists(xq, rows, offsets, device=None):
    """Nearest row of each query's own candidate list on the GPU (N2; the reference's
    ``get_best_positive_index``, datasets/datasets_ws_kitti360.py:976-983, batched).  ``rows`` are the gathered
    candidate features, list q = ``rows[offsets[q]:offsets[q+1]]``.  Returns ``(best_distance fp32 [nq],
    best_position int64 [nq])`` -- position inside the query's own list, first on ties, -1 if empty."""
    lib = _lib.load()
    device = default_device() if device is None else int(device)
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, xq.shape[1])
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nq, d = xq.shape
    assert len(offsets) == nq + 1 and offsets[-1] == len(rows)
    bd = np.empty(nq, dtype=np.float32)
    bp = np.empty(nq, dtype=np.int64)
    _lib.check(lib.agp_best_of_lists(device, nq, d, ctypes.c_void_p(xq.ctypes.data), ctypes.c_void_p(rows.ctypes.data),
                                     ctypes.c_void_p(offsets.ctypes.data), ctypes.c_void_p(bd.ctypes.data),
                                     ctypes.c_void_p(bp.ctypes.data)), "agp_best_of_lists")
    return bd, bp


def recall_hits(predictions, positives_per_query, recall_values, device=None):
    """K5 on the GPU: hits[i] = #queries whose first correct prediction has rank < recall_values[i].

    ``predictions`` is the ``I`` array of ``search`` (numpy or CUDA tensor).  Mirrors the loop at
    reference test.py:72-83 (``recalls[i:] += 1; break``)."""
    lib = _lib.load()
    device = default_device() if device is None else int(device)
    ns = np.ascontiguousarray(recall_values, dtype=np.int32)
    hits = np.zeros(len(ns), dtype=np.int64)
    offsets, ids = positives_to_csr(positives_per_query)
    nq, k = predictions.shape
    if _is_torch(predictions) and predictions.is_cuda:
        import torch
        pred = predictions.to(torch.int64).contiguous()
        d_off = torch.from_numpy(offsets).to(pred.device)
        d_ids = torch.from_numpy(ids if len(ids) else np.zeros(1, np.int64)).to(pred.device)
        stream = torch.cuda.current_stream(pred.device).cuda_stream
        _lib.check(lib.agp_recall_at_n(pred.device.index, ctypes.c_void_p(stream), ctypes.c_void_p(pred.data_ptr()), _lib.MEM_DEVICE,
                                       nq, k, ctypes.c_void_p(d_off.data_ptr()), ctypes.c_void_p(d_ids.data_ptr()),
                                       ctypes.c_void_p(ns.ctypes.data), len(ns), ctypes.c_void_p(hits.ctypes.data)), "agp_recall_at_n")
        return hits
    if _is_torch(predictions):
        predictions = predictions.numpy()
    pred = np.ascontiguousarray(predictions, dtype=np.int64)
    _lib.check(lib.agp_recall_at_n(device, None, ctypes.c_void_p(pred.ctypes.data), _lib.MEM_HOST, nq, k,
                                   ctypes.c_void_p(offsets.ctypes.data), ctypes.c_void_p(ids.ctypes.data),
                                   ctypes.c_void_p(ns.ctypes.data), len(ns), ctypes.c_void_p(hits.ctypes.data)), "agp_recall_at_n")
    return hits


def radius_neighbors(database_xy, queries_xy, radius, device=None, return_csr=False):
    """GPU restatement of the reference's positives computation,
    ``NearestNeighbors().fit(database_utms).radius_neighbors(queries_utms, radius=r, return_distance=False)``
    (datasets/datasets_ws_kitti360.py:613-618, 740-745): fp64, inclusive radius.  Returns an object array of int64 id
    arrays like sklearn (ids ascending; sklearn's order is the tree's), or ``(offsets, ids)`` CSR with
    ``return_csr=True`` -- the layout ``recall_hits`` and the C ABI consume."""
    lib = _lib.load()
    device = default_device() if device is None else int(device)
    db = np.ascontiguousarray(database_xy, dtype=np.float64)
    q = np.ascontiguousarray(queries_xy, dtype=np.float64)
    assert db.ndim == 2 and q.ndim == 2 and db.shape[1] == q.shape[1]
    n, dim = db.shape
    nq = q.shape[0]
    counts = np.zeros(nq, dtype=np.int64)
    _lib.check(lib.agp_radius_count(device, n, dim, ctypes.c_void_p(db.ctypes.data), nq, ctypes.c_void_p(q.ctypes.data),
                                    float(radius), ctypes.c_void_p(counts.ctypes.data)), "agp_radius_count")
    offsets = np.zeros(nq + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    ids = np.empty(max(int(offsets[-1]), 1), dtype=np.int64)
    _lib.check(lib.agp_radius_fill(device, n, dim, ctypes.c_void_p(db.ctypes.data), nq, ctypes.c_void_p(q.ctypes.data),
                                   float(radius), ctypes.c_void_p(offsets.ctypes.data), ctypes.c_void_p(ids.ctypes.data)),
               "agp_radius_fill")
    ids = ids[: int(offsets[-1])]
    if return_csr:
        return offsets, ids
    out = np.empty(nq, dtype=object)
    for i in range(nq):
        out[i] = ids[offsets[i]:offsets[i + 1]]
    return out
