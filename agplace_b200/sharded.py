"""Row-sharded multi-GPU ``IndexFlatL2``: one process per GPU, one exchange step.

North-star scheme (BASELINE.json): the database is split row-wise across the
ranks of a ``torch.distributed`` group, every rank searches the (replicated) query batch against
its shard with the single-GPU engine, the per-shard ``(D fp32, I int64 global)[nq, k]`` lists are
exchanged with ONE all-gather of a packed byte buffer (NCCL over NVLink on GPUs, gloo in the CPU
tests), and every rank runs the K4 merge kernel (``agp_merge_topk``) so all ranks return the same
canonical (distance, id) ordered result as a single index would.  ``search(..., dst=r)`` delivers
the merged result to rank ``r`` only (the lists travel to that rank alone; the other ranks return
``(None, None)``), which is what an evaluation that consumes the neighbours in one process needs.
On GPUs the exchange of large batches goes through symmetric peer memory (copy-engine pushes per
query chunk beside the next chunk's search, one signal barrier, one merge): ``_PeerExchange``.

``shard="query"`` is the zero-compute-redundancy alternative for databases that fit one GPU:
every rank holds the whole database and searches only its slice of the queries; the all-gather
then just concatenates.

The reference has no multi-GPU retrieval (faiss-cpu, one process: SURVEY.md section 2), so this
module extends the drop-in surface rather than mirroring a reference file.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .index import METRIC_INNER_PRODUCT, METRIC_L2, IndexFlatIP, IndexFlatL2, _is_torch


def _cuda_merge(D_lists, d_stride, I_lists, i_stride, nq, k, n_lists, id_bound, metric=METRIC_L2, out=None):
    """K4 across shards on the GPU through the C ABI (device tensors in, device tensors out)."""
    import torch
    if not D_lists.is_cuda:
        raise RuntimeError("agplace_b200 has no CPU merge: per-shard lists must be CUDA tensors "
                           "(tests inject a merge_fn for the gloo/CPU host-logic checks)")
    dev = D_lists.device
    if out is not None:
        D, I = out
    else:
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.agp_merge_topk_metric(dev.index, ctypes.c_void_p(stream), nq, k, n_lists, ctypes.c_void_p(D_lists.data_ptr()),
                                         d_stride, ctypes.c_void_p(I_lists.data_ptr()), i_stride, id_bound, int(metric),
                                         ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr())), "agp_merge_topk_metric")
    return D, I


def shard_bounds(n, world):
    """Contiguous row ranges: shard g holds [g*ceil(n/G), min((g+1)*ceil(n/G), n))."""
    per = -(-n // world) if n else 0
    return [(min(g * per, n), min((g + 1) * per, n)) for g in range(world)]


class _PeerExchange:
    """The exchange step over NVLink peer memory instead of an NCCL kernel: every rank owns a symmetric buffer
    ``[2 slots][world][list bytes]`` (torch symmetric memory: the same allocation mapped into every rank's address
    space).  As soon as a query chunk's local search has finished, the rank PUSHES that chunk's rows of its per-shard
    lists into row ``rank`` of every peer's buffer with plain device-to-device copies -- copy engines, no SMs, so the
    pushes of chunk c run while the persistent screen kernel of chunk c + 1 owns every SM and nothing competes with it.
    After the last chunk ONE signal barrier makes all arrivals visible and ONE K4 launch merges the ``world`` lists.
    (Merging per chunk was measured slower: a merge kernel queued between two screen launches delays the next launch's
    CTA pairs, and the lockstep sweep makes all pairs wait for the late ones -- 8 GPUs, cfg4: 127 ms per step against
    114 ms with one NCCL all-gather at the end.)  Slots alternate per search: a peer can only overwrite slot ``s``
    after it has passed the barrier of the search in between, which this rank enters after its merge of slot ``s``."""

    def __init__(self, group, device, list_bytes):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.list_bytes = (int(list_bytes) + 15) // 16 * 16
        g = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(2 * self.world * self.list_bytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, g.group_name)
        self.peers = [self.hdl.get_buffer(p, (2, self.world, self.list_bytes), torch.uint8) for p in range(self.world)]
        if any(t.data_ptr() == 0 for t in self.peers):
            raise RuntimeError("symmetric memory returned an unmapped peer buffer")
        self.n = 0

    def begin(self, nq, k):
        """Start one search: returns (slot, d_bytes) -- D lists live at [0, nq k 4), I lists at [d_bytes, ...) of a row."""
        d_bytes = (nq * k * 4 + 7) // 8 * 8
        assert d_bytes + nq * k * 8 <= self.list_bytes
        s = self.n & 1
        self.n += 1
        return s, d_bytes

    def push(self, s, d_bytes, a, k, D_loc, I_loc, dst=None):
        """Rows [a, a + m) of this rank's lists -> row ``rank`` of slot ``s`` on every rank, or only on rank ``dst``
        (current stream, copy engines)."""
        import torch
        m_k = D_loc.numel()
        Db = D_loc.reshape(-1).view(torch.uint8)
        Ib = I_loc.reshape(-1).view(torch.uint8)
        targets = [(self.rank + off) % self.world for off in range(self.world)] if dst is None else [dst]
        for peer in targets:                               # start with myself, then the ring: spreads the link load
            row = self.peers[peer][s, self.rank]
            row[a * k * 4: a * k * 4 + m_k * 4].copy_(Db, non_blocking=True)
            row[d_bytes + a * k * 8: d_bytes + a * k * 8 + m_k * 8].copy_(Ib, non_blocking=True)

    def finish(self, s):
        """All pushes of all ranks have landed once this returns (on the stream): the local slot [world, list_bytes]."""
        self.hdl.barrier(channel=s)
        return self.buf.view(2, self.world, self.list_bytes)[s]


class ShardedIndexFlatL2:
    _metric = METRIC_L2
    _local_cls = IndexFlatL2

    def __init__(self, d, group=None, device=None, precision=None, shard="db", index_cls=None, merge_fn=None,
                 result_device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        if shard not in ("db", "query"):
            raise ValueError("shard must be 'db' or 'query'")
        self.d = int(d)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard = shard
        self.is_trained = True
        self.metric_type = self._metric
        self._ntotal = 0
        index_cls = index_cls or self._local_cls
        kwargs = {}
        if index_cls is self._local_cls:
            kwargs = dict(device=device, precision=precision)
        self.local = index_cls(self.d, **kwargs)
        self._merge = merge_fn or _cuda_merge
        self._chunks = []          # (local_start, global_start, count) of this rank's rows, add order
        self._local_rows = 0
        self._result_device = result_device   # torch device for the exchange (None: cuda:<index device>)
        self.phase_events = None              # bench.py: set to [] to collect (name, start_event, end_event) per search

    @property
    def ntotal(self):
        return self._ntotal

    # ------------------------------------------------------------------ add
    def add(self, x):
        """Every rank passes the same full chunk; each keeps its contiguous slice of it."""
        n, d = x.shape
        assert d == self.d
        if self.shard == "query":
            self.local.add(x)
            self._ntotal += n
            return
        lo, hi = shard_bounds(n, self.world)[self.rank]
        if hi > lo:
            self.local.add(x[lo:hi])
            self._record_chunk(self._ntotal + lo, hi - lo)
        self._ntotal += n

    def _record_chunk(self, global_start, count):
        if self._chunks and self._chunks[-1][1] + self._chunks[-1][2] == global_start:
            ls, gs, c = self._chunks[-1]            # contiguous with the previous chunk: extend it
            self._chunks[-1] = (ls, gs, c + count)
        else:
            self._chunks.append((self._local_rows, int(global_start), count))
        self._local_rows += count

    def add_local(self, x_local, global_start, n_global_added):
        """Rank-local variant for databases that never exist in one piece (generated per shard):
        this rank's rows get ids global_start.. ; ``n_global_added`` rows were added job-wide."""
        assert self.shard == "db"
        n, d = x_local.shape
        assert d == self.d
        if n:
            self.local.add(x_local)
            self._record_chunk(int(global_start), n)
        self._ntotal += int(n_global_added)

    def reset(self):
        self.local.reset()
        if hasattr(self.local, "set_id_base"):
            self.local.set_id_base(0)
        self._chunks = []
        self._local_rows = 0
        self._ntotal = 0

    # ------------------------------------------------------------------ search
    def _phase(self, name):
        """Context manager recording a CUDA-event pair around one phase of search() when ``phase_events`` is a list."""
        import contextlib
        import torch
        if self.phase_events is None or not torch.cuda.is_available():
            return contextlib.nullcontext()

        @contextlib.contextmanager
        def cm():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            yield
            e1.record()
            self.phase_events.append((name, e0, e1))
        return cm()

    def phase_ms(self, reset=True):
        """{phase: total milliseconds} of the collected events (synchronises)."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.phase_events or []:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        if reset and self.phase_events is not None:
            self.phase_events = []
        return out

    def _engine_applies_base(self):
        """One contiguous chunk on this rank: the engine adds its global start itself (agp_index_set_id_base).  The base
        is re-derived from the chunk table before EVERY search and cleared otherwise, so no stale base survives an
        add / search / add or a reset (it is engine state, not something to remember here)."""
        return len(self._chunks) == 1 and hasattr(self.local, "set_id_base") and self.shard == "db"

    def _to_global(self, I_local, base_applied=False):
        """local row -> global id through the per-chunk bases (monotone, so tie order is preserved)."""
        import torch
        if base_applied:
            return I_local      # id_base already applied by the engine
        starts = torch.tensor([c[0] for c in self._chunks] or [0], dtype=torch.int64, device=I_local.device)
        gbase = torch.tensor([c[1] - c[0] for c in self._chunks] or [0], dtype=torch.int64, device=I_local.device)
        which = torch.searchsorted(starts, I_local.clamp(min=0), right=True) - 1
        return torch.where(I_local < 0, I_local, I_local + gbase[which.clamp(min=0)])

    def search(self, x, k, *, params=None, D=None, I=None, gather=True, dst=None):
        """``gather=False`` (query-split only): return just this rank's slice of the results -- rows
        ``shard_bounds(nq, world)[rank]`` of ``(D, I)`` -- and skip the all-gather; the results then stay
        partitioned by query like the work was (no collective on the data path).
        ``dst=r`` (row-sharded only; like ``torch.distributed.gather``): only rank ``r`` receives ``(D, I)``, every other
        rank returns ``(None, None)`` -- the per-shard lists travel to rank ``r`` alone, only that rank merges and copies
        the result to its host (the reference evaluates in ONE process: test.py:27-32 is called from the training process)."""
        import torch
        import torch.distributed as dist
        nq, d = x.shape
        assert d == self.d
        assert k > 0
        k = int(k)
        as_numpy = not _is_torch(x)
        if self.shard == "query":
            lo, hi = shard_bounds(nq, self.world)[self.rank]
            per = -(-nq // self.world) if nq else 0
            xs = x[lo:hi]
        else:
            xs = x
        if dst is not None:
            assert self.shard == "db" and D is None and I is None and 0 <= int(dst) < self.world
            dst = int(dst)
        base_applied = self._engine_applies_base()
        if hasattr(self.local, "set_id_base"):
            self.local.set_id_base(self._chunks[0][1] if base_applied else 0)
        native = self.shard == "db" and isinstance(self.local, IndexFlatL2) and self._merge is _cuda_merge and torch.cuda.is_available()
        if native and as_numpy and D is None and I is None and nq * d * 4 >= self.PIPELINE_MIN_BYTES:
            return self._search_host_pipelined(np.ascontiguousarray(x, dtype=np.float32), k, base_applied, dst)
        if native and _is_torch(x) and x.is_cuda and D is None and I is None and nq > self.PIPELINE_CHUNK + self.PIPELINE_CHUNK // 4:
            return self._search_device_chunked(x, k, base_applied, dst)
        if isinstance(self.local, IndexFlatL2) and not (_is_torch(xs) and xs.is_cuda) and xs.shape[0]:
            # host queries: one H2D copy, then everything (search, exchange, merge) stays on the GPU
            xh = xs if _is_torch(xs) else torch.from_numpy(np.ascontiguousarray(xs, dtype=np.float32))
            xs = xh.to(torch.device("cuda", self.local.device), dtype=torch.float32, non_blocking=True)
        if xs.shape[0]:
            with self._phase("local_search"):
                D_loc, I_loc = self.local.search(xs, k)
        else:
            D_loc, I_loc = np.empty((0, k), np.float32), np.empty((0, k), np.int64)
        if not _is_torch(D_loc):
            dev = self._result_device
            if dev is None:
                dev = torch.device("cuda", self.local.device) if hasattr(self.local, "device") else torch.device("cpu")
            D_loc = torch.from_numpy(np.ascontiguousarray(D_loc)).to(dev)
            I_loc = torch.from_numpy(np.ascontiguousarray(I_loc)).to(dev)
        dev = D_loc.device

        if self.shard == "query" and not gather:
            if as_numpy:
                return D_loc.cpu().numpy(), I_loc.cpu().numpy()
            return D_loc, I_loc
        if self.shard == "query":
            # pad every slice to `per` rows, all-gather, drop the padding
            rows = per
        else:
            I_loc = self._to_global(I_loc, base_applied)
            rows = nq
        d_bytes = (rows * k * 4 + 7) // 8 * 8
        i_bytes = rows * k * 8
        with self._phase("pack"):
            send = torch.zeros(d_bytes + i_bytes, dtype=torch.uint8, device=dev)
            n_loc = D_loc.shape[0] * k
            send[: n_loc * 4].view(torch.float32).copy_(D_loc.reshape(-1))
            send[d_bytes: d_bytes + n_loc * 8].view(torch.int64).copy_(I_loc.reshape(-1))
            recv = torch.empty(self.world * (d_bytes + i_bytes), dtype=torch.uint8, device=dev)
        with self._phase("all_gather"):
            dist.all_gather_into_tensor(recv, send, group=self.group)        # the single exchange step
        recv2 = recv.view(self.world, d_bytes + i_bytes)

        if self.shard == "query":
            bounds = shard_bounds(nq, self.world)
            Dg = torch.cat([recv2[g, : (b - a) * k * 4].view(torch.float32).view(b - a, k) for g, (a, b) in enumerate(bounds)])
            Ig = torch.cat([recv2[g, d_bytes: d_bytes + (b - a) * k * 8].view(torch.int64).view(b - a, k)
                            for g, (a, b) in enumerate(bounds)])
        else:
            stride = d_bytes + i_bytes
            D_lists = recv.view(torch.float32)                            # list g starts at g * stride / 4 floats
            # int64 view must start 8-byte aligned: d_bytes is a multiple of 8
            I_lists = recv.view(torch.int64)[d_bytes // 8:]
            id_bound = self._ntotal
            if dst is not None and self.rank != dst:
                return None, None
            with self._phase("merge"):
                if self._metric == METRIC_L2:
                    Dg, Ig = self._merge(D_lists, stride // 4, I_lists, stride // 8, nq, k, self.world, id_bound)
                else:
                    Dg, Ig = self._merge(D_lists, stride // 4, I_lists, stride // 8, nq, k, self.world, id_bound, self._metric)

        if D is not None:
            (D if _is_torch(D) else torch.from_numpy(D)).copy_(Dg)
            Dg = D
        if I is not None:
            (I if _is_torch(I) else torch.from_numpy(I)).copy_(Ig)
            Ig = I
        if as_numpy and D is None:
            Dg = Dg.cpu().numpy()
        if as_numpy and I is None:
            Ig = Ig.cpu().numpy()
        return Dg, Ig


    # ------------------------------------------------------------------ host-buffer pipeline (row-sharded)
    PIPELINE_MIN_BYTES = 16 << 20
    PIPELINE_CHUNK = 18944           # queries per chunk: one wave of 256-query pair tiles on a 148-SM B200

    USE_PEER_MEMORY = True           # exchange through symmetric peer memory (copy engines); False / unavailable: NCCL all-gather

    def _peer_exchange(self, dev, list_bytes):
        """The symmetric-memory exchange object, (re)built collectively when a larger list size is needed; None if this
        torch build / topology cannot provide it (every rank takes the same decision: allocation is all-or-nothing)."""
        import torch
        import torch.distributed as dist
        if not self.USE_PEER_MEMORY:
            return None
        ex = getattr(self, "_peer", None)
        if ex is not None and ex is not False and ex.list_bytes >= list_bytes:
            return ex
        if ex is False:
            return None
        ok = torch.ones(1, device=dev)
        try:
            new = _PeerExchange(self.group, dev, list_bytes)
        except Exception:                                   # noqa: BLE001 -- any failure means "not available here"
            new = None
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() < 1:
            self._peer = False
            return None
        self._peer = new
        return new

    def _exchange_and_merge(self, D_loc, I_loc, nq, k, out=None):
        """The one exchange step: this shard's lists go to every rank, K4 merge (ties by global id)."""
        import torch
        import torch.distributed as dist
        dev = D_loc.device
        d_bytes = (nq * k * 4 + 7) // 8 * 8
        i_bytes = nq * k * 8
        with self._phase("pack"):
            send = torch.empty(d_bytes + i_bytes, dtype=torch.uint8, device=dev)
            send[: nq * k * 4].view(torch.float32).copy_(D_loc.reshape(-1))
            send[d_bytes:].view(torch.int64).copy_(I_loc.reshape(-1))
            recv = torch.empty(self.world * (d_bytes + i_bytes), dtype=torch.uint8, device=dev)
        with self._phase("all_gather"):
            dist.all_gather_into_tensor(recv, send, group=self.group)
        stride = d_bytes + i_bytes
        with self._phase("merge"):
            args = (recv.view(torch.float32), stride // 4, recv.view(torch.int64)[d_bytes // 8:], stride // 8, nq, k, self.world, self._ntotal)
            if out is not None and self._merge is _cuda_merge:
                return _cuda_merge(*args, self._metric, out=out)
            return self._merge(*args) if self._metric == METRIC_L2 else self._merge(*args, self._metric)

    def _merge_slot(self, slot, d_bytes, nq, k, out):
        import torch
        stride = slot.shape[1]
        flat = slot.reshape(-1)                           # [world * list_bytes]: list g starts at g * stride bytes
        return _cuda_merge(flat.view(torch.float32), stride // 4, flat.view(torch.int64)[d_bytes // 8:], stride // 8, nq, k, self.world,
                           self._ntotal, self._metric, out=out)

    def _search_device_chunked(self, x, k, base_applied, dst=None):
        """CUDA queries in, CUDA results out, one wave of query tiles at a time: while the (persistent, SM-filling) screen
        kernel of chunk c + 1 runs on the main stream, the lists of chunk c travel to every rank through peer memory on a
        second stream (copy engines only); one barrier + one merge launch after the last chunk."""
        import torch
        nq = x.shape[0]
        dev = x.device
        ex = self._peer_exchange(dev, (nq * k * 4 + 7) // 8 * 8 + nq * k * 8)
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_s_comm"):
            self._s_comm = torch.cuda.Stream(dev)
        comm = self._s_comm
        x = x.detach().to(torch.float32).contiguous()
        chunk = self.PIPELINE_CHUNK
        if ex is None:                                    # no peer memory here: one NCCL all-gather + merge at the end
            with self._phase("local_search"):
                D_loc, I_loc = self.local.search(x, k)
            out = self._exchange_and_merge(D_loc, self._to_global(I_loc, base_applied), nq, k)
            return out if dst is None or self.rank == dst else (None, None)
        mine = dst is None or self.rank == dst
        D = torch.empty((nq, k), dtype=torch.float32, device=dev) if mine else None
        I = torch.empty((nq, k), dtype=torch.int64, device=dev) if mine else None
        slot_id, d_bytes = ex.begin(nq, k)
        comm.wait_stream(main)
        for a in range(0, nq, chunk):
            b = min(nq, a + chunk)
            with self._phase("local_search"):
                D_loc, I_loc = self.local.search(x[a:b], k)
            I_loc = self._to_global(I_loc, base_applied)
            ev = torch.cuda.Event(); ev.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                ex.push(slot_id, d_bytes, a, k, D_loc, I_loc, dst)
            D_loc.record_stream(comm); I_loc.record_stream(comm)
        main.wait_stream(comm)
        with self._phase("exchange_barrier"):
            slot = ex.finish(slot_id)
        if mine:
            with self._phase("merge"):
                self._merge_slot(slot, d_bytes, nq, k, (D, I))
        return D, I

    def _search_host_pipelined(self, x, k, base_applied, dst=None):
        """numpy in / numpy out on every rank, as a pipeline over query chunks: host copy into a pinned buffer + H2D
        (copy stream) | local search + all-gather + merge (compute stream; nothing in it synchronises the host) | D2H
        into the pinned result arrays (copy-out stream).  Every rank issues the same sequence of collectives."""
        import torch
        from .index import _result_array
        nq = x.shape[0]
        dev = torch.device("cuda", self.local.device)
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._stage = [None, None]
        chunk = self.PIPELINE_CHUNK
        first = max(256, chunk // 4)                    # a short first chunk: the GPU starts after a quarter-wave of host copy
        cuts = [0] + list(range(first, nq, chunk)) + [nq] if nq > first + chunk // 2 else [0, nq]
        mine = dst is None or self.rank == dst
        D = _result_array((nq, k), np.float32) if mine else None
        I = _result_array((nq, k), np.int64) if mine else None
        Dt, It = (torch.from_numpy(D), torch.from_numpy(I)) if mine else (None, None)
        pinned_out = mine and Dt.is_pinned() and It.is_pinned()
        xt = torch.from_numpy(x)
        xq_dev = torch.empty((nq, self.d), dtype=torch.float32, device=dev)
        stage_ev = [None, None]
        pending = []
        ex = self._peer_exchange(dev, (nq * k * 4 + 7) // 8 * 8 + nq * k * 8)
        slot_id, d_bytes = ex.begin(nq, k) if ex is not None else (0, 0)
        self._s_out.wait_stream(main)
        self._s_in.wait_stream(main)                        # xq_dev's block may still be in use by work queued on main
        for c in range(len(cuts) - 1):
            a, b = cuts[c], cuts[c + 1]
            st = self._stage[c & 1]
            if st is None or st.shape[0] < b - a:
                st = self._stage[c & 1] = torch.empty((max(chunk, b - a), self.d), dtype=torch.float32, pin_memory=True)
            if stage_ev[c & 1] is not None:
                stage_ev[c & 1].synchronize()              # the DMA that last read this pinned buffer is done
            st[: b - a].copy_(xt[a:b])                      # host memcpy: overlaps the GPU work of the previous chunks
            with torch.cuda.stream(self._s_in):
                xq_dev[a:b].copy_(st[: b - a], non_blocking=True)
                ev_in = torch.cuda.Event(); ev_in.record()
            stage_ev[c & 1] = ev_in
            main.wait_event(ev_in)
            with self._phase("local_search"):
                D_loc, I_loc = self.local.search(xq_dev[a:b], k)
            I_loc = self._to_global(I_loc, base_applied)
            if ex is not None:                          # peer memory: copy-engine pushes beside the next chunk's search
                ev = torch.cuda.Event(); ev.record(main)
                with torch.cuda.stream(self._s_out):
                    self._s_out.wait_event(ev)
                    ex.push(slot_id, d_bytes, a, k, D_loc, I_loc, dst)
                D_loc.record_stream(self._s_out); I_loc.record_stream(self._s_out)
                continue
            Dg, Ig = self._exchange_and_merge(D_loc, I_loc, b - a, k)      # NCCL fallback: all-gather + merge per chunk, in line
            if not mine:
                continue
            ev = torch.cuda.Event(); ev.record(main)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev)
                Dt[a:b].copy_(Dg, non_blocking=pinned_out)
                It[a:b].copy_(Ig, non_blocking=pinned_out)
            Dg.record_stream(self._s_out); Ig.record_stream(self._s_out)
            pending.append((Dg, Ig))
        if ex is not None:                              # one barrier + one merge launch, then the results leave in four pieces
            main.wait_stream(self._s_out)
            slot = ex.finish(slot_id)
            if mine:
                Dg = torch.empty((nq, k), dtype=torch.float32, device=dev)
                Ig = torch.empty((nq, k), dtype=torch.int64, device=dev)
                self._merge_slot(slot, d_bytes, nq, k, (Dg, Ig))
                self._s_out.wait_stream(main)
                with torch.cuda.stream(self._s_out):
                    Dt.copy_(Dg, non_blocking=pinned_out)
                    It.copy_(Ig, non_blocking=pinned_out)
                Dg.record_stream(self._s_out); Ig.record_stream(self._s_out)
        self._s_out.synchronize()
        xq_dev.record_stream(self._s_in)
        return D, I


class ShardedIndexFlatIP(ShardedIndexFlatL2):
    """Row- or query-sharded ``IndexFlatIP``: same exchange, the merge orders the per-shard lists by descending product
    (ties by global id) and pads with ``(-3.4028235e38, -1)``."""

    _metric = METRIC_INNER_PRODUCT
    _local_cls = IndexFlatIP
