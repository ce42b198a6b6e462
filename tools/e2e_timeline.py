#!/usr/bin/env python
"""Device-side timeline of one cc_overlap_task call on host arrays (the e2e leg of bench.py): when each copy and
kernel group finishes relative to the start of the call.  Re-implements tasks._start_chunk/_finish_chunk with
timing events; run by hand on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as orc
from synaptor_b200 import _device as D
from synaptor_b200.proc import tasks as T
from synaptor_b200.proc.seg import continuation as cont_mod
from synaptor_b200.proc import seg

shape = (512, 512, 512)
pred = torch.from_numpy(orc.synth_pred(shape, seed=0, sigma=2.0, fg=0.03)).pin_memory().numpy()
base = torch.from_numpy(orc.synth_base(shape, seed=1000).view(np.int64)).pin_memory().numpy().view(np.uint64)
dev = torch.device("cuda", 0)
for _ in range(2):
    T.cc_overlap_task(pred, base, 0.5, 100)


def ev(stream):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


for rep in range(3):
    torch.cuda.synchronize()
    host_t0 = time.perf_counter()
    marks = {}
    s_main = torch.cuda.current_stream()
    s_in = D.side_stream(torch, dev, "in")
    e0 = ev(s_main)
    s_in.wait_stream(s_main)
    with torch.cuda.stream(s_in):
        dp = D.to_device_f32(torch, pred, dev)
        marks["pred H2D done"] = ev(s_in)
        db = D.to_device_u64(torch, base, dev)
        marks["base H2D done"] = ev(s_in)
    s_main.wait_event(marks["pred H2D done"])
    res = D.chunk_ccs_device(dp, np.float32(0.5), 100)
    marks["chunk_ccs done"] = ev(s_main)
    h = {"chunk_ccs returned": time.perf_counter() - host_t0}
    faces = cont_mod.faces_from_device(res.labels)       # small copies before the label volume (as tasks._materialise_cc)
    table = res.table_numpy()
    h["faces on host"] = time.perf_counter() - host_t0
    s_out = D.side_stream(torch, dev, "out")
    host, evl = res.labels_numpy_async(s_out)
    marks["labels D2H done"] = ev(s_out)
    conts, _ = cont_mod.continuations_from_faces(faces)
    info = seg.misc.seg_info_from_table(table[:, 0], table[:, 1:])
    h["continuations + table built"] = time.perf_counter() - host_t0
    s_main.wait_event(marks["base H2D done"])
    ov = D.count_overlaps_device(res.labels, db)
    marks["overlaps done"] = ev(s_main)
    h["count_overlaps returned"] = time.perf_counter() - host_t0
    evl.synchronize()
    coo = T._coo_from_result(ov)
    h["COO built (end)"] = time.perf_counter() - host_t0
    torch.cuda.synchronize()
    print(f"-- rep {rep}")
    for (k, e) in marks.items():
        print(f"  device {k:28s} {e0.elapsed_time(e):7.2f} ms")
    for (k, v) in h.items():
        print(f"  host   {k:28s} {v * 1e3:7.2f} ms")
