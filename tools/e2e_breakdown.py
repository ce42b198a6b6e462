#!/usr/bin/env python
"""Where the end-to-end time of cc_overlap_task (host arrays in, host objects out) goes on this box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as orc
from synaptor_b200.proc import tasks as T

shape = (512, 512, 512)
pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=2.0, fg=0.03)).pin_memory()
base = torch.from_numpy(orc.synth_base(shape, seed=1000).view(np.int64)).pin_memory()
dev = torch.device("cuda", 0)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dp = torch.empty_like(pred, device=dev)
db = torch.empty_like(base, device=dev)
print("H2D pred  537 MB: %.2f ms" % timeit(lambda: dp.copy_(pred, non_blocking=True)))
print("H2D base 1074 MB: %.2f ms" % timeit(lambda: db.copy_(base, non_blocking=True)))
lab = torch.empty(shape, dtype=torch.int32, device=dev)
hl = torch.empty(shape, dtype=torch.int32).pin_memory()
print("D2H labels 537 MB: %.2f ms" % timeit(lambda: hl.copy_(lab, non_blocking=True)))
s2 = torch.cuda.Stream()


def both():
    db.copy_(base, non_blocking=True)
    with torch.cuda.stream(s2):
        hl.copy_(lab, non_blocking=True)


print("H2D base + D2H labels concurrently: %.2f ms" % timeit(both))
pn, bn = pred.numpy(), base.numpy().view(np.uint64)
print("cc_overlap_task (host arrays): %.2f ms" % timeit(lambda: T.cc_overlap_task(pn, bn, 0.5, 100), n=5))
print("cc_task only (host array): %.2f ms" % timeit(lambda: T.cc_task(pn, 0.5, 100, verbose=False), n=5))
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    T.cc_overlap_task(pn, bn, 0.5, 100)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
