"""Development aid: the remainder cost model of search_screen (waves x (range + per-item overhead)) -- device-resident
search time over batch sizes for several values of the overhead constant (knob screen_item_overhead, tenths of a tile).
    python scripts/split_model_probe.py [reps] [balanced|overhead]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import torch

import agplace_b200 as agp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = sys.argv[2] if len(sys.argv) > 2 else "balanced"
rng = np.random.default_rng(2)


def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def timed(ix, xq, k, n=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ix.search(xq, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, n, d, k, sizes in (("100kx512 k=50", 100000, 512, 50, [1000, 2048, 2560, 4096, 5120, 6144, 8192, 10240, 12288, 14336, 16128, 17920, 20000, 23040, 25600, 28160]),
                             ("10kx256 k=20 (cfg1 db)", 10000, 256, 20, [2000, 4096, 8192, 12288, 16128]),
                             ("1Mx128 k=10", 1000000, 128, 10, [2048, 5120, 8192, 12288, 16128]),
                             ("100kx512 k=10 (cfg3)", 100000, 512, 10, [1000]),
                             ("20kx4096 k=100", 20000, 4096, 100, [2048, 8192, 16128]),
                             ("200kx64 k=10", 200000, 64, 10, [4096, 8192, 16128])):
    ix = agp.IndexFlatL2(d)
    ix.add(unit(n, d))
    for nq in sizes:
        xq = torch.from_numpy(unit(nq, d)).cuda()
        D0, I0 = ix.search(xq, k)
        res = {"nq": nq, "tiles": -(-nq // 256)}
        # interleaved A/B (the clock drifts under load): mode "overhead" = the earlier per-item constant (3 tiles) vs the
        # automatic one; mode "balanced" = equal ranges only (screen_balanced = 0) vs automatic (balanced when cheaper)
        knob, variants, labels = (("screen_item_overhead", (30, 0), {30: "c0=3", 0: "auto"}) if mode == "overhead" else
                                  ("screen_balanced", (0, -1), {0: "equal ranges", -1: "auto"}))
        best = {v: 1e9 for v in variants}
        for rnd in range(max(2, reps // 4)):
            for ov in variants if rnd % 2 == 0 else variants[::-1]:
                ix.set_knob(knob, ov)
                if rnd == 0:
                    D, I = ix.search(xq, k)
                    assert torch.equal(I, I0) and torch.equal(D, D0), (name, nq, ov)
                    timed(ix, xq, k, 2)
                best[ov] = min(best[ov], timed(ix, xq, k))
        res.update({labels[ov]: round(best[ov], 4) for ov in variants})
        ix.set_knob("screen_item_overhead", 0)
        ix.set_knob("screen_balanced", -1)
        print(json.dumps({name: res}), flush=True)
    del ix
