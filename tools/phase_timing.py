#!/usr/bin/env python
"""
Prints the in-kernel phase timers of an instrumented build (make -C synaptor_b200/csrc clean && make TIMING=1):
cycles summed over CTAs (thread 0 clock64 deltas) for k_tile_ccl, per phase, on the 512^3 bench workload.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as orc
from synaptor_b200 import _device as D

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 512, 512)
sigma, fg = (float(sys.argv[4]), float(sys.argv[5])) if len(sys.argv) >= 6 else (2.0, 0.03)
dev = torch.device("cuda", 0)
pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=sigma, fg=fg)).to(dev)
base = torch.from_numpy(orc.synth_base(shape, seed=1000).view(np.int64)).to(dev)
res = None
for _ in range(3):
    res = D.chunk_ccs_device(pred, np.float32(0.5), 100, base=base, sync=True, out=res)
c = res.counts
names = ["load", "starts+scan", "local tables", "unions", "flatten", "slot init", "statistics", "records"]
dbg = c["dbg"]
n = max(1, dbg[8])
print({k: c[k] for k in ("runs", "components", "kept", "pairs", "active_words", "bnd_words", "tile_records", "flagged_tiles")})
print(f"k_tile_ccl: {n} non-empty CTAs; cycles per CTA by phase (thread 0):")
tot = sum(dbg[:8])
for (nm, v) in zip(names, dbg[:8]):
    print(f"  {nm:14s} {v / n:9.0f} cyc  {100.0 * v / max(1, tot):5.1f}%")
print(f"  total          {tot / n:9.0f} cyc = {tot / n / 1.965e3:.2f} us per CTA")
print(f"  union passes (warp 0, both passes): setup {dbg[9] / n:.0f} cyc, rounds {dbg[10] / n:.0f} cyc in {dbg[11] / n:.2f} rounds, "
      f"of which sm_union {dbg[12] / n:.0f} cyc")
print(f"  jump: barrier wait {dbg[13] / n:.0f} cyc, jumping {dbg[14] / n:.0f} cyc in {dbg[15] / n:.2f} rounds")
