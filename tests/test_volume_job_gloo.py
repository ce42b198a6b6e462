"""
world_size-2 `gloo` test (CPU) of the host-side plumbing of synaptor_b200.multi.VolumeJob: chunk ownership over an
uneven, ragged grid, capacity agreement (all_reduce), the packed-slot exchange (all_gather_into_tensor of uint8
slots), slot addressing through the merge plan, the sizing pass and the overflow / regrow protocol.  The CUDA entry
points are replaced by numpy stand-ins that implement the documented contracts of include/synaptor_b200.h
(syn_chunk_ccs_begin / syn_merge_ccs_packed / syn_merge_ccs_table / syn_chunk_ccs_finish), i.e. the packed slot
format; the result must equal the oracle workflow bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from synaptor_b200 import multi

HDR = 16


class NumpyVolumeJob(multi.VolumeJob):
    """multi.VolumeJob with the four CUDA calls restated in numpy (tests only)."""

    def set_host_inputs(self, li, pred, base):
        self.pred[li], self.base[li] = pred, base

    def _alloc_chunk(self, li):
        self.ws[li] = {"counts": {}}
        self.pairs[li] = None
        self.graphs = None

    def _read_counts(self, li):
        c = {"errflags": 0, "runs": 1, "pairs": 0, "max_label": 0, "max_base": 0, "kept": 0}
        c.update(self.ws[li]["counts"])
        return c

    def _begin(self, li, slot=None, cap_rows=None, cap_entries=None):
        cap_rows, cap_entries = cap_rows or self.cap_rows, cap_entries or self.cap_entries
        slot = slot if slot is not None else self.my_slots[li]
        ci = self.mine[li]
        ccs, _, info = orc.cc_task_c(self.pred[li], float(self.thresh), self.dust, offset=self.chunk_begin[ci],
                                     with_continuations=False)
        ids = np.asarray(info.index, dtype=np.int64)
        table = np.concatenate([ids[:, None], info.to_numpy(dtype=np.int64).reshape(len(ids), 10)], 1)
        rel = np.zeros(int(ccs.max()) + 1, dtype=np.uint32)
        rel[ids] = np.arange(1, len(ids) + 1, dtype=np.uint32)
        raw = slot.numpy()
        raw[:HDR * 8] = 0
        hdr = raw[:HDR * 8].view(np.int64)
        hdr[0] = len(ids)
        if len(ids) > cap_rows:
            hdr[8] |= 1
        tab = raw[HDR * 8:HDR * 8 + cap_rows * 88].view(np.int64).reshape(cap_rows, 11)
        tab[:min(len(ids), cap_rows)] = table[:cap_rows]
        ent = raw[HDR * 8 + cap_rows * 88:HDR * 8 + cap_rows * 88 + cap_entries * 8].view(np.uint32).reshape(cap_entries, 2)
        n = 0
        for (f, (axis, hi)) in enumerate(multi.FACE_ORDER):
            if not (self.face_masks[li] >> f) & 1:
                continue
            face = np.take(ccs, -1 if hi else 0, axis=axis).ravel()
            pos = np.flatnonzero(face)
            hdr[2 + f] = n
            k = len(pos)
            if n + k > cap_entries:
                hdr[8] |= 2
                k = max(0, cap_entries - n)
            ent[n:n + k, 0] = pos[:k]
            ent[n:n + k, 1] = rel[face[pos[:k]]]
            n += len(pos)
            hdr[9 + f] = n
        hdr[1] = n
        self.ws[li]["ccs"], self.ws[li]["rel"] = ccs, rel
        self.ws[li]["counts"] = {"kept": len(ids), "runs": 1}

    def _slot_views(self, slot):
        raw = self.all_slots[slot].numpy()
        hdr = raw[:HDR * 8].view(np.int64)
        tab = raw[HDR * 8:HDR * 8 + self.cap_rows * 88].view(np.int64).reshape(self.cap_rows, 11)
        ent = raw[HDR * 8 + self.cap_rows * 88:].view(np.uint32).reshape(-1, 2)
        return hdr, tab, ent

    def _merge_ids(self):
        plan = self.plan.numpy()
        nchunks, npairs = int(plan[0]), int(plan[1])
        slots = plan[2:2 + nchunks]
        pairs = plan[2 + nchunks:].reshape(npairs, 4)
        info = self.info.numpy()
        info[:] = 0
        views = [self._slot_views(int(s)) for s in slots]
        if any(h[8] or h[0] > self.cap_rows or h[1] > self.cap_entries for (h, _, _) in views):
            info[3] = 1
            return
        kept = np.array([int(h[0]) for (h, _, _) in views], dtype=np.int64)
        base = np.concatenate(([0], np.cumsum(kept)))
        self.id_base.numpy()[:] = base
        total = int(base[-1])
        par = np.arange(total + 1)

        def find(x):
            while par[x] != x:
                par[x] = par[par[x]]
                x = par[x]
            return x
        nedges = 0
        for (ca, fa, cb, fb) in pairs.tolist():
            (ha, _, ea), (hb, _, eb) = views[ca], views[cb]
            A, B = ea[ha[2 + fa]:ha[9 + fa]], eb[hb[2 + fb]:hb[9 + fb]]
            lut = dict(zip(B[:, 0].tolist(), B[:, 1].tolist()))
            for (p, r) in A.tolist():
                if p in lut:
                    a, b = find(int(base[ca]) + r), find(int(base[cb]) + lut[p])
                    nedges += 1
                    if a != b:
                        par[max(a, b)] = min(a, b)
        gmap = np.array([find(i) for i in range(total + 1)])
        rows = np.concatenate([t[:k] for ((_, t, _), k) in zip(views, kept)]) if total else np.zeros((0, 11), np.int64)
        msize = np.zeros(total + 1, dtype=np.int64)
        np.add.at(msize, gmap[1:], rows[:, 1])
        fmap = self.fmap.numpy().view(np.uint32)
        fmap[:total + 1] = np.where(msize[gmap] >= self.size_thr, gmap, 0)
        fmap[0] = 0
        info[0], info[2] = total, nedges
        self._gmap, self._rows = gmap, rows

    def _merge_table(self):
        info = self.info.numpy()
        if info[3]:
            return
        gmap, rows = self._gmap, self._rows
        out = []
        for d in sorted(set(gmap[1:].tolist())):
            mem = np.flatnonzero(gmap[1:] == d)
            S = int(rows[mem, 1].sum())
            if S < self.size_thr:
                continue
            w = rows[mem, 1].astype(np.float64) / np.float64(S)
            cent = [int(orc.kahan_sum(rows[mem, 2 + a].astype(np.float64) * w)) for a in range(3)]
            out.append([d, S, *cent, *rows[mem, 5:8].min(0).tolist(), *rows[mem, 8:11].max(0).tolist()])
        m = self.merged.numpy()
        m[:len(out)] = np.array(out, dtype=np.int64).reshape(-1, 11)
        info[1] = len(out)

    def _finish(self, li):
        ci = self.mine[li]
        if self.info.numpy()[3]:
            return
        base = int(self.id_base.numpy()[ci])
        fmap = self.fmap.numpy().view(np.uint32)
        rel = self.ws[li]["rel"]
        lut = np.where(rel > 0, fmap[np.minimum(base + rel, len(fmap) - 1)], 0).astype(np.uint32)
        self.labels[li].numpy().view(np.uint32)[:] = lut[self.ws[li]["ccs"]]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vshape, cshape = (40, 36, 50), (16, 18, 20)        # 3 x 2 x 3 = 18 chunks, ragged, 9 per rank
        out = {"rank": rank, "runs": []}
        job = NumpyVolumeJob(vshape, cshape, 0.5, 6, 10, with_base=False, rank=rank, world=world,
                             device=torch.device("cpu"), vol_offset=(7, 8, 9), graph=False)
        for (seed, fg) in ((21, 0.04), (22, 0.30)):        # the second, denser volume overflows the first capacities
            vol = orc.synth_pred(vshape, seed=seed, sigma=1.5, fg=fg)
            for li, ci in enumerate(job.mine):
                beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
                sl = tuple(slice(b, b + s) for (b, s) in zip(beg, shp))
                job.set_host_inputs(li, np.ascontiguousarray(vol[sl]), None)
            attempts = 0
            for attempts in range(1, 6):
                job.run()
                if job.check():
                    break
            out["runs"].append({"attempts": attempts, "caps": (job.cap_rows, job.cap_entries),
                                "labels": {ci: job.labels[li].numpy().view(np.uint32).copy() for li, ci in enumerate(job.mine)},
                                "merged": job.merged_table().copy(), "info": job.info_dict(),
                                "maps": [job.chunk_id_map(ci) for ci in range(job.nchunks)]})
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_volume_job_plumbing_world2_gloo():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    vshape, cshape = (40, 36, 50), (16, 18, 20)
    for (k, (seed, fg)) in enumerate(((21, 0.04), (22, 0.30))):
        vol = orc.synth_pred(vshape, seed=seed, sigma=1.5, fg=fg)
        want = orc.volume_job(vol, None, cshape, 0.5, 6, 10, vol_offset=(7, 8, 9))
        wm = {int(a): [int(x) for x in v] for (a, v) in want["merged"].items()}
        seen = set()
        for o in outs:
            r = o["runs"][k]
            assert r["info"]["flags"] == 0 and r["info"]["n_merged"] == len(wm)
            assert {int(x[0]): x[1:].tolist() for x in r["merged"]} == wm
            for ci in range(len(want["chunks"])):
                assert r["maps"][ci] == {int(a): int(b) for (a, b) in want["id_maps"][want["chunks"][ci][0]].items()}
            for (ci, lab) in r["labels"].items():
                assert np.array_equal(lab, want["final"][want["chunks"][ci][1]])
                seen.add(ci)
        assert seen == set(range(18))
    # both ranks agreed on the capacities, and the denser volume made them grow (collectively)
    assert outs[0]["runs"][0]["caps"] == outs[1]["runs"][0]["caps"]
    assert outs[0]["runs"][1]["caps"] == outs[1]["runs"][1]["caps"]
    assert outs[0]["runs"][1]["caps"][1] > outs[0]["runs"][0]["caps"][1]       # more face entries
    assert outs[0]["runs"][1]["attempts"] >= 2


def _worker_idle(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vshape, cshape = (24, 20, 16), (24, 20, 16)          # ONE chunk, two ranks: rank 1 owns nothing
        vol = orc.synth_pred(vshape, seed=5, sigma=1.5, fg=0.15)
        job = NumpyVolumeJob(vshape, cshape, 0.5, 6, 10, with_base=False, rank=rank, world=world,
                             device=torch.device("cpu"), graph=False)
        for li, ci in enumerate(job.mine):
            job.set_host_inputs(li, vol, None)
        job.run()
        ok = job.check()
        q.put({"rank": rank, "ok": ok, "mine": list(job.mine), "merged": job.merged_table().copy(),
               "labels": [job.labels[li].numpy().view(np.uint32).copy() for li in range(len(job.mine))]})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_rank_without_chunks_world2_gloo():
    """More ranks than chunks: the idle rank still takes part in every collective and gets the merge results."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_idle, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=100) for _ in range(2)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    vol = orc.synth_pred((24, 20, 16), seed=5, sigma=1.5, fg=0.15)
    want = orc.volume_job(vol, None, (24, 20, 16), 0.5, 6, 10)
    assert outs[0]["mine"] == [0] and outs[1]["mine"] == [] and outs[0]["ok"] and outs[1]["ok"]
    assert np.array_equal(outs[0]["labels"][0], want["final"])
    wm = {int(a): [int(x) for x in v] for (a, v) in want["merged"].items()}
    for o in outs:
        assert {int(x[0]): x[1:].tolist() for x in o["merged"]} == wm


def test_plan_and_grid_helpers():
    grid, chunks = multi.grid_chunks((50, 40, 72), (24, 16, 32))
    assert grid == (3, 3, 3) and len(chunks) == 27
    assert chunks[0] == ((0, 0, 0), (24, 16, 32)) and chunks[-1] == ((48, 32, 64), (2, 8, 8))
    plan, per_rank, npairs = multi.merge_plan((2, 1, 2), 3)
    assert per_rank == 2 and npairs == 4
    assert plan[:2].tolist() == [4, 4] and plan[2:6].tolist() == [0, 2, 4, 1]
    assert plan[6:].reshape(4, 4).tolist() == [[0, 0, 2, 1], [0, 4, 1, 5], [1, 0, 3, 1], [2, 4, 3, 5]]
