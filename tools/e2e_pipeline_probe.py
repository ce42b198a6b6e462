#!/usr/bin/env python
"""Probe of the pipelined multi-chunk path: per-chunk completion times and where host threads block."""
import os
import sys
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
from synaptor_b200 import _device as D
from synaptor_b200.proc import tasks as T

shape = (512, 512, 512)
pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=2.0, fg=0.03)).pin_memory()
base = torch.from_numpy(orc.synth_base(shape, seed=1000).view(np.int64)).pin_memory()
pn, bn = pred.numpy(), base.numpy().view(np.uint64)

spent = defaultdict(list)


def wrap(mod, name):
    fn = getattr(mod, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            spent[name].append((time.perf_counter() - t0) * 1e3)
    setattr(mod, name, w)


for n in ("to_device_f32", "to_device_u64", "chunk_ccs_device", "count_overlaps_device", "extract_faces_device"):
    wrap(D, n)
wrap(T, "_materialise_cc")
_empty = torch.empty


def timed_empty(*a, **k):
    if k.get("pin_memory"):
        t0 = time.perf_counter()
        r = _empty(*a, **k)
        spent["pinned_empty"].append((time.perf_counter() - t0) * 1e3)
        return r
    return _empty(*a, **k)


torch.empty = timed_empty
workers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in T.cc_overlap_tasks([(pn, bn)] * 6, 0.5, 100):
    pass
torch.cuda.synchronize()
for rep in range(3):
    spent.clear()
    n = 12
    t0 = time.perf_counter()
    ts = []
    for out in T.cc_overlap_tasks([(pn, bn)] * n, 0.5, 100):
        ccs, conts, info, mat = out
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print("rep", rep, "workers", workers, "per chunk %.2f ms" % (tot / n * 1e3), "completion", [round(x * 1e3) for x in ts])
    for (k, v) in spent.items():
        print(f"  {k:24s} n={len(v):3d} mean={np.mean(v):7.2f} max={np.max(v):7.2f} ms")
