#!/usr/bin/env python
"""Where the time of multi.merge_ccs_device goes (single process, 8 chunks of one volume on one GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
from synaptor_b200 import _device as D
from synaptor_b200 import multi

cs = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 256, 256)
grid = (2, 2, 2)
dev = torch.device("cuda", 0)
one = orc.synth_pred(cs, seed=0, sigma=2.0, fg=0.03)
results = []
for gi in np.ndindex(grid):
    beg = tuple(a * b for (a, b) in zip(gi, cs))
    results.append(D.chunk_ccs_device(torch.from_numpy(one).to(dev), np.float32(0.5), 100, offset=beg))
torch.cuda.synchronize()
import cProfile
import pstats
for rep in range(3):
    for r in results:       # fresh labels each time (the merge relabels in place)
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = multi.merge_ccs_device(results, grid, rank=0, world=1, size_thr=100, remap=False)
    torch.cuda.synchronize()
    print("merge_ccs_device: %.3f ms, merged %d, edges %d" % ((time.perf_counter() - t0) * 1e3, out.n_merged, out.n_edges))
pr = cProfile.Profile()
pr.enable()
out = multi.merge_ccs_device(results, grid, rank=0, world=1, size_thr=100, remap=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
