"""Development aid (GPU): per-shape timing of the search path -- device step, dominant kernel, phases, numpy e2e --
with the instrumented kernel's cycle report once per shape, and a sweep of the host pipeline's chunk size.

    python scripts/shape_probe.py cfg1 cfg3 cfg2 [--counters] [--pipe 0,2560,5120,10240]
"""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

import agplace_b200 as agp
from agplace_b200 import synth


def probe(name, counters, pipe, reps, sched=(), first_cuts=()):
    c = dict(synth.CONFIGS[name])
    n, nq, d, k = c["n"], c["nq"], c["d"], c["k"]
    xb = synth.descriptors(n, d, c["seed"], "db")
    xq = synth.descriptors(nq, d, c["seed"] + 7, "q")
    ix = agp.IndexFlatL2(d)
    ix.add(xb)
    xq_d = torch.from_numpy(xq).cuda()
    for _ in range(3):
        ix.search(xq_d, k)
    torch.cuda.synchronize()
    if counters:
        ix.set_knob("cycle_counters", 1)
        ix.search(xq_d, k)
        torch.cuda.synchronize()
        ix.set_knob("cycle_counters", 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ix.search(xq_d, k)
    host_enqueue_ms = (time.perf_counter() - t0) / reps * 1e3      # host time to enqueue one search (the GPU runs behind)
    torch.cuda.synchronize()
    ix.set_profiling(True)
    ix.get_profile_phases(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.search(xq_d, k)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / reps
    ph = ix.get_profile_phases(reset=True)
    ix.set_profiling(False)
    kms = ph["distance"][0] / max(ph["distance"][1], 1) * (ph["distance"][1] / reps)
    out = dict(cfg=name, step_ms=round(step, 4), kernel_ms=round(kms, 4), tflops=round(2.0 * nq * n * d / (kms * 1e-3) / 1e12, 1),
               phases_ms={p: round(v[0] / reps, 4) for p, v in ph.items()}, kernel_share=round(kms / step, 3), host_enqueue_ms=round(host_enqueue_ms, 4))
    if sched:
        sw = {}
        for m in sched:
            ix.set_knob("screen_sched", m)
            for _ in range(3):
                ix.search(xq_d, k)
            torch.cuda.synchronize()
            ix.set_profiling(True); ix.get_profile_phases(reset=True)
            for _ in range(reps):
                ix.search(xq_d, k)
            torch.cuda.synchronize()
            p2 = ix.get_profile_phases(reset=True); ix.set_profiling(False)
            sw[str(m)] = round(p2["distance"][0] / reps, 4)
        ix.set_knob("screen_sched", 0)
        out["kernel_ms_by_sched_quarters"] = sw
    e2e = {}
    for chunk in pipe:
        ix.set_knob("pipe_chunk", chunk)
        for _ in range(3):
            ix.search(xq, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ix.search(xq, k)
        torch.cuda.synchronize()
        e2e[str(chunk)] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    ix.set_knob("pipe_chunk", 0)
    out["numpy_e2e_ms_by_pipe_chunk"] = e2e
    emin = {}
    for kb in (512, 1 << 20):
        ix.set_knob("pipe_min_kb", kb)
        for _ in range(3):
            ix.search(xq, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ix.search(xq, k)
        torch.cuda.synchronize()
        emin["staged" if kb == 512 else "driver_pageable"] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    ix.set_knob("pipe_min_kb", 512)
    out["numpy_e2e_ms_small_call"] = emin
    e2f = {}
    for first in first_cuts:
        ix.set_knob("pipe_first", first)
        for _ in range(3):
            ix.search(xq, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ix.search(xq, k)
        torch.cuda.synchronize()
        e2f[str(first)] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    ix.set_knob("pipe_first", 0)
    if e2f:
        out["numpy_e2e_ms_by_first_chunk"] = e2f
    # where the host path spends its time: pinned in / pinned preallocated out (pure pipeline), pageable in + reused
    # pageable out (no page faults on the results), and the default call (fresh numpy results)
    xp = torch.from_numpy(xq).pin_memory().numpy()
    Dp, Ip = torch.empty((nq, k), dtype=torch.float32).pin_memory().numpy(), torch.empty((nq, k), dtype=torch.int64).pin_memory().numpy()
    Dn, In = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    variants = {"pinned_in_pinned_out": lambda: ix.search(xp, k, D=Dp, I=Ip), "pageable_in_pinned_out": lambda: ix.search(xq, k, D=Dp, I=Ip),
                "pinned_in_reused_pageable_out": lambda: ix.search(xp, k, D=Dn, I=In), "pageable_in_reused_pageable_out": lambda: ix.search(xq, k, D=Dn, I=In),
                "pageable_in_fresh_out": lambda: ix.search(xq, k)}
    ix.set_profiling(True)
    ix.get_profile_phases(reset=True)
    for _ in range(reps):
        ix.search(xq, k)
    torch.cuda.synchronize()
    ph = ix.get_profile_phases(reset=True)
    ix.set_profiling(False)
    out["numpy_path_phases_ms"] = {p: (round(v[0] / reps, 4), v[1] // reps) for p, v in ph.items()}
    vt = {}
    for nm, fn in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        vt[nm] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    out["host_path_variants_ms"] = vt
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cfgs", nargs="*", default=["cfg1", "cfg3", "cfg2"])
    ap.add_argument("--counters", action="store_true")
    ap.add_argument("--pipe", default="0")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sched", default="")
    ap.add_argument("--first", default="")
    a = ap.parse_args()
    for name in a.cfgs:
        probe(name, a.counters, [int(x) for x in a.pipe.split(",")], a.reps, [int(x) for x in a.sched.split(",") if x], [int(x) for x in a.first.split(",") if x])
