#!/usr/bin/env python
"""Latency of torch.distributed.all_gather_into_tensor (NCCL) for the slot sizes of the volume job, eager and
replayed from a CUDA graph.  torchrun --nproc-per-node N tools/allgather_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for kb in (8, 64, 256, 1024, 2304, 4096):
    n = kb * 1024
    src = torch.zeros(n, dtype=torch.uint8, device=dev)
    dst = torch.zeros(n * world, dtype=torch.uint8, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(dst, src)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_gather_into_tensor(dst, src)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 50 * 1e3
    graph_us = None
    if "--graph" in sys.argv:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                dist.all_gather_into_tensor(dst, src)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(50):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / 50 * 1e3
    res[kb] = (round(eager, 1), None if graph_us is None else round(graph_us, 1))
if rank == 0:
    print("ALLGATHER_US per-rank KB -> (eager, graph):", res, "world", world, flush=True)
dist.barrier()
dist.destroy_process_group()
