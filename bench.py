#!/usr/bin/env python
"""
bench.py -- throughput of the chunk_ccs + chunk_overlaps hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU path on this box's host cores)

A "step" is one pass of the fused chunk_ccs + chunk_overlaps pipeline over one
512^3 chunk per GPU (BASELINE.json configs[1]: float32 cleft prediction, threshold
0.5, 6-connectivity, dust 100, uint64 base segmentation).  With N GPUs every rank
owns one chunk of an (N*512) x 512 x 512 volume (weak scaling; chunks shard
naturally, no data-path collective in the timed step).  The cross-chunk merge_ccs
exchange (NCCL all_gather of segment counts, boundary faces and tables, device
union-find, LUT remap) is run and timed separately (config.merge_ccs_ms_separate).

value   : Gvoxel/s, whole job, inputs resident in HBM, outputs left in HBM (CUDA events, max over ranks)
e2e     : the same through the public API synaptor_b200.proc.tasks.cc_overlap_task with pinned HOST
          buffers, H2D + D2H + host-side table/continuation materialisation inside the timed region
roofline: the dominant kernel (dense label write + pair histogram, 12 B/voxel algorithmic) timed with
          CUDA events on the launching stream in a second, event-instrumented pass of K steps
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = (512, 512, 512)
THRESH, DUST, CONN = 0.5, 100, 6
METRIC = "Gvoxels/s chunk_ccs+overlaps"
BYTES_PER_VOXEL_PIPELINE = 16   # 4 B pred read + 4 B label write + 8 B base read
BYTES_PER_VOXEL_STREAM_KERNEL = 16  # k_stream: 4 B pred read + 8 B base read + 4 B label (zero-fill) write
CPU_SAMPLE = (256, 256, 256)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_stream launch on the 512^3 workload (ncu --set full,
# profiles/r01m_top_summary.txt: 1.610673 GB read + 491.01 MB written in 331.9 us (under ncu); the algorithmic 16 B/voxel are
# 2.147 GB -- the foreground sectors are not zero-filled by this kernel)
STREAM_TRAFFIC_512 = 2101686120


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/syn_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi needs a moment to start (longer on an 8-GPU box): do not enter the timed region before the
        first sample has been written, or a 7 ms region is over before the sampler runs."""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1]))
                except ValueError:
                    continue
                for (n, v) in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.remove(self.path)
        except OSError:
            pass
        # "under load" = the upper half of the samples (idle samples before/after the region excluded)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's numpy port (reference cost profile: np.unique, find_objects, Counter ...)
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, shape = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as orc
    pred = orc.synth_pred(shape, seed=seed, sigma=2.0, fg=0.03)
    base = orc.synth_base(shape, seed=seed + 1)
    t0 = time.perf_counter()
    ccs, _, info = orc.cc_task_np(pred, THRESH, DUST)
    orc.count_overlaps_np(ccs, base)
    return time.perf_counter() - t0, len(info)


def cpu_baseline(nproc=None, shape=CPU_SAMPLE, reps=1):
    """One reference-style task (cc_task + overlap_task) per process, one chunk each, like the reference's
    one-task-per-chunk deployment.  Returns (Gvoxel/s aggregate, cores, wall seconds)."""
    import multiprocessing as mp
    from oracle import oracle as orc
    cores = len(os.sched_getaffinity(0))
    nproc = nproc or max(1, cores)          # every host core the process may use: one reference-style task per core
    ctx = mp.get_context("fork")
    jobs = [(100 + i, shape) for i in range(nproc * reps)]
    with ctx.Pool(nproc) as pool:
        pool.map(_cpu_worker, [(1, (32, 32, 32))] * nproc)     # import + page-in outside the timing
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    # synthetic-input generation happens inside the workers but outside their timed span; use the sum of the
    # per-task compute spans spread over nproc workers as the wall time of the compute
    compute_wall = sum(r[0] for r in res) / nproc
    vox = float(np.prod(shape)) * len(jobs)
    kind = "port"
    ref_so = orc.ref_ext("_describe") is not None
    sample = (f"{len(jobs)} chunks of {shape[0]}x{shape[1]}x{shape[2]} (same generator/params as the GPU workload), "
              f"one oracle cc_task_np + count_overlaps_np per process on {nproc} processes; numpy/scipy port of the "
              f"reference's Python steps with scipy.ndimage.label for the absent cc3d"
              + (" and the reference's own compiled _describe/_relabel from oracle/_ref" if ref_so else ""))
    return vox / compute_wall / 1e9, nproc, compute_wall, kind, sample, wall


def workload_name(shape):
    return (f"configs[1]: chunk_ccs+chunk_overlaps {shape[0]}x{shape[1]}x{shape[2]} per GPU, "
            "float32 pred (sigma 2, fg 3%), thr 0.5, 6-conn, dust 100, uint64 base (32^3 blocks)")


def run_reference(args, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # each step = one bounded sample of the workload (one 256^3 chunk per host core, all cores busy); a single
    # warm-up sample whatever W is, so that the whole run stays within a few minutes
    vals, walls = [], []
    info = None
    n_warm = 1 if args.warmup >= 1 else 0
    for _ in range(n_warm):
        cpu_baseline(reps=1)
    for _ in range(max(1, args.steps)):
        info = cpu_baseline(reps=1)
        vals.append(info[0])
        walls.append(info[2])
    v = float(np.mean(vals))
    gv, cores, wall, kind, sample, _ = info
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Gvoxel/s", "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": n_warm, "ms_per_step": float(np.mean(walls)) * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32->u32/u64", "data": "synthetic",
        "config": {"workload": workload_name(tuple(args.chunk)), "sample_shape": list(CPU_SAMPLE),
                   "sample": "each step is a bounded CPU sample of the workload: " + sample},
        "cpu_baseline": {"value": v, "unit": "Gvoxel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line, out_fd)
    return 0


# --------------------------------------------------------------------------------------------------
def emit(line, fd):
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # exactly ONE JSON line may reach stdout: libraries (e.g. NCCL's version banner) print there too, so
    # everything else is sent to stderr and the JSON goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the plain syn_chunk_ccs call, not its CUDA graph")
    ap.add_argument("--chunk", type=int, nargs=3, default=list(CHUNK))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, real_stdout)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from oracle import oracle as orc          # synthetic-input generator + cpu_baseline leg only
    import synaptor_b200 as syn
    from synaptor_b200 import _cabi as C
    from synaptor_b200 import _device as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    shape = tuple(args.chunk)
    nvox = int(np.prod(shape))
    # ---- synthetic inputs (pinned host copies for the e2e leg, device copies for the resident leg) ----
    pred_np = orc.synth_pred(shape, seed=rank, sigma=2.0, fg=0.03)
    base_np = orc.synth_base(shape, seed=1000 + rank)
    pred_pin = torch.from_numpy(pred_np).pin_memory()
    base_pin = torch.from_numpy(base_np.view(np.int64)).pin_memory()
    del pred_np, base_np
    pred = pred_pin.to(dev)
    base = base_pin.to(dev)
    t32 = np.float32(THRESH)
    L = C.lib()
    grid = (world, 1, 1)
    offset = (rank * shape[0], 0, 0)

    merge_out = [None]
    if world > 1:
        from synaptor_b200 import multi

    def step(out=None):
        res = D.chunk_ccs_device(pred, t32, DUST, connectivity=CONN, base=base, offset=offset, sync=False, out=out)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed steps replay ONE captured CUDA graph of the syn_chunk_ccs call (D.ChunkPipeline: stream capture of
    # the same C-ABI call on the resident input tensors; 13 kernels + 1 memset per replay) -- what a production loop
    # over equally shaped chunks does.  --no-graph times the plain call instead.
    pipe = None if args.no_graph else D.ChunkPipeline(shape, DUST, thresh=THRESH, connectivity=CONN, offset=offset,
                                                      device=dev, pred=pred, base=base)
    # first call sizes the buffers (sync=True grows capacities if needed)
    res = D.chunk_ccs_device(pred, t32, DUST, connectivity=CONN, base=base, offset=offset, sync=True)
    counts = dict(res.counts)
    run_stream = pipe.stream if pipe is not None else torch.cuda.current_stream()
    if pipe is not None:
        pipe.run()
        assert pipe.counts()["kept"] == counts["kept"]

    def timed_step():
        if pipe is not None:
            pipe.graph.replay()
        else:
            step(res)

    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first_sample()
    with torch.cuda.stream(run_stream):
        for _ in range(2 * args.warmup):        # load already running when the timed steps start (clock samples)
            timed_step()
        launches0 = L.syn_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            timed_step()
        e1.record()
        barrier()
    launches = (args.steps * pipe.launches_per_run) if pipe is not None else (L.syn_launch_count() - launches0)
    ms_total = e0.elapsed_time(e1)
    after = pipe.counts() if pipe is not None else D.read_counts(res)
    assert after["errflags"] == 0 and after["kept"] == counts["kept"] and after["pairs"] == counts["pairs"], after
    # ---- cross-chunk merge_ccs exchange (config 4), timed on its own: NCCL all_gathers of counts, faces and
    # tables + device union-find + LUT remap of every rank's label volume.  Not part of the headline metric
    # (chunk_ccs + chunk_overlaps shard with no data-path collective).
    n_merged, merge_ms = None, None
    if world > 1:
        times = []
        for _ in range(max(3, min(args.steps, 5))):
            res = step(res)
            res.counts = {}
            barrier()
            t0 = time.perf_counter()
            merge_out[0] = multi.merge_ccs_device(res, grid, rank=rank, world=world, size_thr=DUST)
            barrier()
            times.append((time.perf_counter() - t0) * 1e3)
        n_merged = merge_out[0].n_merged
        merge_ms = float(np.median(times))

    # ---- event-instrumented passes: one CUDA event after every kernel of syn_chunk_ccs ----------------
    def profiled(mode):
        C.check(L.syn_profile_enable(mode))
        acc = {}
        for _ in range(args.steps):
            D.chunk_ccs_device(pred, t32, DUST, connectivity=CONN, base=base, offset=offset, sync=False, out=res)
            for (k, v) in C.profile_read().items():
                acc[k] = acc.get(k, 0.0) + v / args.steps
        C.check(L.syn_profile_enable(0))
        return acc
    stage_ms = profiled(1)
    for _ in range(50):                     # ~35 ms more of the same step under the sampler (20 ms sampling period)
        res = step(res)
    torch.cuda.synchronize()
    clocks = sampler.stop()

    # ---- e2e through the public API with pinned host buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        pn, bn = pred_pin.numpy(), base_pin.numpy().view(np.uint64)
        # one public call per chunk: pinned host arrays in (H2D inside), numpy / DataFrame / COO out (D2H + host
        # materialisation inside).  Warm-up in the steady-state pattern of the timed loop (the previous result is
        # still referenced while the next one is produced, so two pinned output blocks circulate).
        n_e2e = max(3, min(args.steps, 6))
        for _ in range(3):
            ccs, conts, info, mat = syn.proc.tasks.cc_overlap_task(pn, bn, THRESH, DUST, offset=offset)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ccs, conts, info, mat = syn.proc.tasks.cc_overlap_task(pn, bn, THRESH, DUST, offset=offset)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        assert len(info) == counts["kept"] and mat.nnz == counts["pairs"]
        # the same chunks through the generator form of the call (cc_overlap_tasks: the next chunk's input copies
        # are queued before the previous chunk is finished, so the PCIe link carries inputs back to back);
        # reported next to the per-call figure, which stays the headline e2e value
        for out in syn.proc.tasks.cc_overlap_tasks(((pn, bn, offset) for _ in range(3)), THRESH, DUST):
            ccs, conts, info, mat = out
        barrier()
        t0 = time.perf_counter()
        for out in syn.proc.tasks.cc_overlap_tasks(((pn, bn, offset) for _ in range(n_e2e)), THRESH, DUST):
            ccs, conts, info, mat = out
        torch.cuda.synchronize()
        e2e_piped_s = (time.perf_counter() - t0) / n_e2e
        assert len(info) == counts["kept"] and mat.nnz == counts["pairs"]
        h2d = pred_pin.numel() * 4 + base_pin.numel() * 8
        d2h = ccs.nbytes + len(info) * 11 * 8 + mat.nnz * 20 + sum(
            2 * (shape[(a + 1) % 3] * shape[(a + 2) % 3]) * 4 for a in range(3))
        e2e = {"seconds": e2e_s, "h2d": h2d, "d2h": d2h, "n": n_e2e, "piped_seconds": e2e_piped_s}

    # ---- reduce over ranks --------------------------------------------------------------------------
    ms_max = ms_total
    e2e_s_max = e2e["seconds"] if e2e else 0.0
    if world > 1:
        t = torch.tensor([ms_total, e2e_s_max], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max, e2e_s_max = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms_max / args.steps
        value = world * nvox / (ms_step * 1e-3) / 1e9
        wl_ms = float(stage_ms['stream'])
        achieved = BYTES_PER_VOXEL_STREAM_KERNEL * nvox / (wl_ms * 1e-3) / 1e9
        pipe_gbs = BYTES_PER_VOXEL_PIPELINE * nvox / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32->u32/u64", "data": "synthetic",
            "config": {"workload": workload_name(shape),
                       "chunks_per_gpu": 1, "grid": list(grid), "merge_ccs_in_step": False,
                       "cuda_graph": pipe is not None,
                       "merge_ccs_ms_separate": merge_ms,
                       "l2": "inputs (1.6 GB/step) exceed the 126 MB L2; no flush needed",
                       "components": counts["components"], "kept": counts["kept"], "pairs": counts["pairs"],
                       "runs": counts["runs"], "fg_voxels": counts["fg_voxels"], "merged_segments": n_merged},
            "roofline": {"bound": "hbm", "kernel": "k_stream<base,scan,tma> (pred + base read [base by TMA bulk copies], threshold, label zero-fill, run prefix)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": STREAM_TRAFFIC_512 if shape == (512, 512, 512) else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_voxel": BYTES_PER_VOXEL_STREAM_KERNEL, "kernel_ms": wl_ms,
                         "pipeline": {"bytes_per_voxel": BYTES_PER_VOXEL_PIPELINE, "achieved": pipe_gbs,
                                      "frac": pipe_gbs / peak, "frac_of_8TBs_nominal": pipe_gbs / 8000.0},
                         "stage_ms": {n: round(float(v), 5) for (n, v) in stage_ms.items()}},
            "clocks": clocks,
            "gpu_launches": int(launches),
        }
        if e2e:
            line["e2e"] = {"value": world * nvox / e2e_s_max / 1e9, "unit": "Gvoxel/s",
                           "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "ms_per_step": e2e_s_max * 1e3, "chunks": int(e2e["n"]),
                           "ms_per_step_generator_call_rank0": e2e["piped_seconds"] * 1e3,
                           "api": "synaptor_b200.proc.tasks.cc_overlap_task (one call per chunk; pinned host arrays in, "
                                  "numpy / DataFrame / COO out)"}
        if not args.no_cpu_baseline and world == 1:
            gv, cores, wall, kind, sample, _ = cpu_baseline()
            line["cpu_baseline"] = {"value": gv, "unit": "Gvoxel/s", "cores": cores, "kind": kind, "sample": sample,
                                    "seconds": wall}
        emit(line, real_stdout)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
