#!/usr/bin/env python
"""How much of the 512^3 step is launch overhead: the same syn_chunk_ccs call replayed from a CUDA graph."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
from synaptor_b200 import _device as D

shape = (512, 512, 512)
dev = torch.device("cuda", 0)
pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=2.0, fg=0.03)).to(dev)
base = torch.from_numpy(orc.synth_base(shape, seed=1000).view(np.int64)).to(dev)
t32 = np.float32(0.5)
res = D.chunk_ccs_device(pred, t32, 100, base=base, sync=True)
want = dict(res.counts)


def step():
    D.chunk_ccs_device(pred, t32, 100, base=base, sync=False, out=res)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager, default stream : %.4f ms" % timeit(step))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    print("eager, side stream    : %.4f ms" % timeit(step))
    g = torch.cuda.CUDAGraph()
    step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step()
    print("graph replay          : %.4f ms" % timeit(g.replay))
got = D.read_counts(res)
print("counts equal after replay:", {k: got[k] for k in ("runs", "components", "kept", "pairs")} ==
      {k: want[k] for k in ("runs", "components", "kept", "pairs")})
