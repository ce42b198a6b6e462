#!/usr/bin/env python
"""One GPU: the volume job with 1 vs 2 volumes in flight (two VolumeJob instances driven alternately).
   python tools/inflight_probe.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from synaptor_b200 import multi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
njobs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
jobs, keep = [], []
for j in range(njobs):
    job = multi.VolumeJob(bench.VOLUME, bench.CHUNK, bench.THRESH, bench.DUST, bench.SIZE_THR, with_base=True, rank=0,
                          world=1, device=dev, graph=True, lanes=lanes)
    for li, ci in enumerate(job.mine):
        if j == 0:
            beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
            keep.append((bench.gen_pred_chunk(torch, dev, bench.VOLUME, beg, shp),
                         bench.gen_base_chunk(torch, dev, bench.VOLUME, beg, shp)))
        job.set_inputs(li, *keep[li])
    job.run()
    assert job.check()
    jobs.append(job)
torch.cuda.synchronize()
for n in range(1, njobs + 1):
    for _ in range(3):
        for j in range(n):
            jobs[j].run(join=False)
    for j in range(n):
        jobs[j].join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        jobs[s % n].run(join=False)
    for j in range(n):
        jobs[j].join()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"lanes {lanes} in flight {n}: {ms:.4f} ms per step = {1.073741824 / ms * 1e3:.1f} Gvoxel/s", [j.info_dict()["flags"] for j in jobs])
