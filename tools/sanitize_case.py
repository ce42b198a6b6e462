#!/usr/bin/env python
"""Small cases that exercise every kernel path once (for compute-sanitizer --tool memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
import synaptor_b200 as syn

rng = np.random.default_rng(0)
cases = [((16, 32, 512), 0.06), ((16, 32, 512), 0.5), ((5, 13, 384), 0.3), ((9, 70, 96), 0.3), ((7, 13, 131), 0.3)]
for (shape, p) in cases:
    pred = rng.random(shape, dtype=np.float32)
    base = orc.synth_base(shape, seed=3, block=4)
    bdev = torch.from_numpy(base.view(np.int64)).cuda()
    ccs, conts, info, mat = syn.proc.tasks.cc_overlap_task(pred, bdev, 1.0 - p, 2)
    want = orc.cc_task_c(pred, 1.0 - p, 2)[0]
    assert np.array_equal(ccs, want)
    syn.proc.tasks.cc_task(pred, 1.0 - p, 2, verbose=False)
    syn.proc.tasks.cc_task(pred, 1.0 - p, 2, overlap_seg=base, verbose=False)
    syn.proc.tasks.cc_task(pred, 1.0 - p, 2, connectivity=26, verbose=False)
    syn.proc.tasks.cc_task(pred, 1.0 - p, 2, dil_param=2, verbose=False)
    syn.proc.tasks.overlap_task(ccs, base)
print("sanitize cases ok")
