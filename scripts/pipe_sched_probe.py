"""Development aid: chunk schedules of the host-buffer search pipeline (agp_index_search with numpy in / numpy out) on
mid-sized batches.  Explicit chunk boundaries through the knobs pipe_cut1..3; "auto" is the library's own schedule.
    python scripts/pipe_sched_probe.py [reps] [sweep|cuts]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

import agplace_b200 as agp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(1)


def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(ts[len(ts) // 2] * 1e3, 3)


T = 256          # queries per pair tile
mode = sys.argv[2] if len(sys.argv) > 2 else "sweep"
if mode == "sweep":      # automatic schedule vs the round-2 schedule (knob pipe_sched = 1) over batch sizes
    for name, n, d, k, sizes in (("100kx512 k=50", 100000, 512, 50, [6000, 12000, 15000, 16000, 18000, 20000, 23000, 26000, 32000, 40000, 57000, 80000]),
                                 ("50kx256 k=20", 50000, 256, 20, [8000, 16000, 20000, 30000, 40000, 80000])):
        xb = unit(n, d)
        ix = agp.IndexFlatL2(d)
        ix.add(xb)
        for nq in sizes:
            xq = unit(nq, d)
            xq_dev = torch.from_numpy(xq).cuda()
            res = {"nq": nq, "device": timed(lambda: (ix.search(xq_dev, k), torch.cuda.synchronize()))}
            ix.set_knob("pipe_sched", 1)
            D0, I0 = ix.search(xq, k)
            res["r2"] = timed(lambda: ix.search(xq, k))
            ix.set_knob("pipe_sched", 0)
            D, I = ix.search(xq, k)
            assert np.array_equal(I, I0) and np.array_equal(D, D0), (name, nq)
            res["auto"] = timed(lambda: ix.search(xq, k))
            res["r2_again"] = (ix.set_knob("pipe_sched", 1), timed(lambda: ix.search(xq, k)))[1]
            ix.set_knob("pipe_sched", 0)
            print(json.dumps({name: res}), flush=True)
        del ix
    sys.exit(0)

if mode == "small":      # transfer-dominated shapes (the reference's own sizes): automatic schedule vs fixed cuts / 8 MB staging pieces
    for name, n, d, k, nq in (("10kx256 k=20", 10000, 256, 20, 2000), ("10kx256 k=20", 10000, 256, 20, 8000), ("10kx256 k=20", 10000, 256, 20, 20000),
                              ("30kx256 k=20", 30000, 256, 20, 6000), ("30kx256 k=20", 30000, 256, 20, 16000), ("100kx512 k=50", 100000, 512, 50, 20000)):
        xb, xq = unit(n, d), unit(nq, d)
        ix = agp.IndexFlatL2(d)
        ix.add(xb)
        xq_dev = torch.from_numpy(xq).cuda()
        res = {"nq": nq, "device": timed(lambda: (ix.search(xq_dev, k), torch.cuda.synchronize())), "auto": timed(lambda: ix.search(xq, k))}
        ix.set_knob("pipe_piece_kb", 8192)
        res["auto, 8 MB pieces"] = timed(lambda: ix.search(xq, k))
        ix.set_knob("pipe_sched", 1)
        res["r2 schedule, 8 MB pieces"] = timed(lambda: ix.search(xq, k))
        ix.set_knob("pipe_sched", 0)
        ix.set_knob("pipe_piece_kb", 0)
        for c1 in (nq // 4, nq // 2, nq):
            c1 = (c1 + 255) // 256 * 256 if c1 < nq else nq
            ix.set_knob("pipe_cut1", c1)
            res["one chunk" if c1 >= nq else f"{c1}|rest"] = timed(lambda: ix.search(xq, k))
        ix.set_knob("pipe_cut1", 0)
        res["auto again"] = timed(lambda: ix.search(xq, k))
        print(json.dumps({name: res}), flush=True)
        del ix
    sys.exit(0)

cases = [
    ("cfg2 100kx512 nq=20000 k=50", 100000, 512, 20000, 50,
     [(), (9 * T, 27 * T), (9 * T, 27 * T, 51 * T), (9 * T, 27 * T, 45 * T), (5 * T, 14 * T, 32 * T), (9 * T, 32 * T), (18 * T, 42 * T),
      (10 * T, 30 * T, 50 * T), (6 * T, 18 * T, 42 * T), (4 * T, 13 * T, 31 * T)]),
    ("100kx512 nq=8000 k=50", 100000, 512, 8000, 50, [(), (4 * T, 13 * T), (9 * T,), (5 * T, 14 * T), (3 * T, 9 * T, 18 * T)]),
    ("100kx512 nq=40000 k=50", 100000, 512, 40000, 50,
     [(), (9 * T, 27 * T, 64 * T), (9 * T, 27 * T, 82 * T), (18 * T, 54 * T, 100 * T), (9 * T, 36 * T, 83 * T)]),
    ("cfg1-like 10kx256 nq=8000 k=20", 10000, 256, 8000, 20, [(), (9 * T,), (4 * T, 13 * T)]),
]
for name, n, d, nq, k, scheds in cases:
    xb, xq = unit(n, d), unit(nq, d)
    ix = agp.IndexFlatL2(d)
    ix.add(xb)
    xq_dev = torch.from_numpy(xq).cuda()
    res = {"device-resident": timed(lambda: (ix.search(xq_dev, k), torch.cuda.synchronize()))}
    D0, I0 = ix.search(xq, k)
    for cuts in scheds:
        for i in range(3):
            ix.set_knob(f"pipe_cut{i + 1}", cuts[i] if i < len(cuts) else 0)
        D, I = ix.search(xq, k)
        assert np.array_equal(I, I0) and np.array_equal(D, D0), (name, cuts)
        res["auto" if not cuts else "|".join(str(c // T) for c in cuts) + " tiles"] = timed(lambda: ix.search(xq, k))
    print(json.dumps({name: res}), flush=True)
    del ix
