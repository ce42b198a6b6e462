#!/usr/bin/env python
"""Runs the north-star volume job (bench.py's default workload) a few times on one GPU, nothing else: the command
that the ncu captures of the job use.   python tools/job_probe.py [steps] [--graph]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from synaptor_b200 import multi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
graph = "--graph" in sys.argv
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
job = multi.VolumeJob(bench.VOLUME, bench.CHUNK, bench.THRESH, bench.DUST, bench.SIZE_THR, with_base=True, rank=0, world=1,
                      device=dev, graph=graph)
keep = []
for li, ci in enumerate(job.mine):
    beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
    p = bench.gen_pred_chunk(torch, dev, bench.VOLUME, beg, shp)
    b = bench.gen_base_chunk(torch, dev, bench.VOLUME, beg, shp)
    keep.append((p, b))
    job.set_inputs(li, p, b)
job.run()
assert job.check()
torch.cuda.synchronize()
print("PROBE_STEADY_STATE launches so far:", int(__import__("synaptor_b200")._cabi.lib().syn_launch_count()), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    job.run()
e1.record()
torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / steps, job.info_dict())
