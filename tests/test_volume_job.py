"""
The whole-volume job (synaptor_b200.multi.VolumeJob: chunk_ccs -> merge_ccs -> remap_ids -> chunk_overlaps with the
label volume written once in final ids) against the oracle's restatement of the reference workflow
(proc/tasks.py:26-69, 72-122, 314-333, 290-297).  Bit-exact, merged centroids included.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from util import oracle_volume_job

pytestmark = pytest.mark.gpu


def run_job(vol, base_vol, cshape, thresh, dust, size_thr, graph, vol_offset=(0, 0, 0), world=1, rank=0):
    import torch
    from synaptor_b200 import multi
    dev = torch.device("cuda", torch.cuda.current_device())
    job = multi.VolumeJob(vol.shape, cshape, thresh, dust, size_thr, with_base=base_vol is not None, rank=rank,
                          world=world, device=dev, vol_offset=vol_offset, graph=graph)
    keep = []
    for li, ci in enumerate(job.mine):
        beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
        sl = tuple(slice(b, b + s) for (b, s) in zip(beg, shp))
        p = torch.from_numpy(np.ascontiguousarray(vol[sl])).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(base_vol[sl]).view(np.int64)).to(dev) if base_vol is not None else None
        keep.append((p, b))
        job.set_inputs(li, p, b)
    for _ in range(4):
        job.run()
        if job.check():
            break
    else:
        raise AssertionError("capacities never settled")
    return job


def compare(job, want, vol_shape, base_vol):
    final = np.zeros(vol_shape, dtype=np.uint32)
    for li, ci in enumerate(job.mine):
        gi, sl, beg = want["chunks"][ci]
        final[sl] = job.labels[li].cpu().numpy().view(np.uint32)
    assert np.array_equal(final, want["final"]), f"{np.count_nonzero(final != want['final'])} voxels differ"
    got = {int(r[0]): r[1:].tolist() for r in job.merged_table()}
    assert got == {int(k): [int(x) for x in v] for (k, v) in want["merged"].items()}
    for ci in range(job.nchunks):
        gi = want["chunks"][ci][0]
        assert job.chunk_id_map(ci) == {int(k): int(v) for (k, v) in want["id_maps"][gi].items()}
        info = want["infos"][gi]
        t = job.chunk_table(ci)
        assert np.array_equal(t[:, 0], np.asarray(info.index, dtype=np.int64))
        assert np.array_equal(t[:, 1:], info.to_numpy(dtype=np.int64).reshape(len(info), 10))
    if base_vol is not None:
        for li, ci in enumerate(job.mine):
            r, c, v, shape = job.overlaps(li)
            wr, wc, wv, wshape = want["overlaps"][ci]
            assert shape == wshape
            assert np.array_equal(r, wr) and np.array_equal(c, wc) and np.array_equal(v, wv)


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("vshape,cshape", [
    ((40, 36, 32), (20, 18, 16)),        # 2x2x2, generic (non-sector) path
    ((48, 24, 64), (16, 24, 32)),        # 3x1x2, sector path (nz % 8 == 0)
    ((32, 32, 256), (16, 16, 128)),      # 2x2x2, TMA stream path (nz % 128 == 0)
    ((50, 40, 72), (24, 16, 32)),        # 3x3x3 ragged: remainder chunks (chunking.py:47-69)
    ((24, 24, 24), (24, 24, 24)),        # a single chunk: no face pairs at all
])
def test_volume_job_vs_oracle(vshape, cshape, graph):
    vol = orc.synth_pred(vshape, seed=11, sigma=2.0, fg=0.10)
    base_vol = orc.synth_base(vshape, seed=12, block=5, bg=0.1)
    want = oracle_volume_job(vol, base_vol, cshape, 0.5, 10, 15, vol_offset=(100, 200, 300))
    job = run_job(vol, base_vol, cshape, 0.5, 10, 15, graph, vol_offset=(100, 200, 300))
    compare(job, want, vshape, base_vol)
    info = job.info_dict()
    assert info["flags"] == 0 and info["n_merged"] == len(want["merged"])
    # the merged volume is a 1:1 relabelling of whole-volume CCL restricted to the surviving segments
    final = want["final"]
    whole = orc.label_c(vol > 0.5, 6)[0]
    pairs = np.unique(np.stack([final[final != 0], whole[final != 0]], 1), axis=0)
    assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


def test_volume_job_no_base_dense_noise():
    rng = np.random.default_rng(3)
    vol = rng.random((36, 40, 64), dtype=np.float32)
    want = oracle_volume_job(vol, None, (12, 20, 32), 0.6, 3, 5)
    job = run_job(vol, None, (12, 20, 32), 0.6, 3, 5, graph=True)
    compare(job, want, vol.shape, None)


def test_volume_job_empty_and_full():
    for fill in (0.0, 1.0):
        vol = np.full((16, 16, 32), fill, dtype=np.float32)
        want = oracle_volume_job(vol, None, (8, 16, 16), 0.5, 10, 10)
        job = run_job(vol, None, (8, 16, 16), 0.5, 10, 10, graph=False)
        compare(job, want, vol.shape, None)


def test_volume_job_rerun_new_data_grows_capacities():
    """The same job object on denser data: check() reports the overflow, grows, and the re-run is exact."""
    import torch
    vshape, cshape = (32, 32, 64), (16, 16, 32)
    sparse = orc.synth_pred(vshape, seed=5, sigma=2.0, fg=0.02)
    job = run_job(sparse, None, cshape, 0.5, 10, 15, graph=True)
    dense = np.random.default_rng(9).random(vshape, dtype=np.float32)
    for li, ci in enumerate(job.mine):
        beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
        sl = tuple(slice(b, b + s) for (b, s) in zip(beg, shp))
        job.pred[li].copy_(torch.from_numpy(np.ascontiguousarray(dense[sl])))
    ok = False
    for _ in range(5):
        job.run()
        if job.check():
            ok = True
            break
    assert ok
    want = oracle_volume_job(dense, None, cshape, 0.5, 10, 15)
    compare(job, want, vshape, None)


@pytest.mark.parametrize("order", ["C", "F"])
def test_volume_tasks_host_arrays(order):
    """The public host-array call (proc.tasks.volume_tasks): C- and F-ordered chunk arrays (the F-ordered ones are
    transposed on the device, conncomps.py:21), results as the reference's task outputs."""
    import synaptor_b200 as syn
    vshape, cshape = (48, 40, 64), (24, 20, 32)
    vol = orc.synth_pred(vshape, seed=17, sigma=2.0, fg=0.10)
    base_vol = orc.synth_base(vshape, seed=18, block=6, bg=0.1)
    want = oracle_volume_job(vol, base_vol, cshape, 0.5, 10, 15, vol_offset=(5, 6, 7))
    conv = np.asfortranarray if order == "F" else np.ascontiguousarray
    preds = [conv(vol[sl]) for (_, sl, _) in want["chunks"]]
    bases = [conv(base_vol[sl]) for (_, sl, _) in want["chunks"]]
    for _ in range(2):          # second call: the cached job, steady state
        out = syn.proc.tasks.volume_tasks(preds, bases, vshape, cshape, 0.5, 10, 15, offset=(5, 6, 7), rank=0, world=1)
    for (ci, (gi, sl, beg)) in enumerate(want["chunks"]):
        assert out.ccs[ci].dtype == np.uint32 and np.array_equal(out.ccs[ci], want["final"][sl])
        assert out.chunk_id_maps[ci] == {int(k): int(v) for (k, v) in want["id_maps"][gi].items()}
        info = want["infos"][gi]
        assert list(out.chunk_infos[ci].columns) == list(info.columns)
        assert np.array_equal(out.chunk_infos[ci].to_numpy(), info.to_numpy())
        assert list(out.chunk_infos[ci].index) == list(info.index)
        wr, wc, wv, wshape = want["overlaps"][ci]
        m = out.overlaps[ci]
        assert m.shape == wshape and np.array_equal(m.row, wr) and np.array_equal(m.col, wc) and np.array_equal(m.data, wv)
    got = {int(i): [int(x) for x in row] for (i, row) in zip(out.merged_info.index, out.merged_info.to_numpy())}
    assert got == {int(k): [int(x) for x in v] for (k, v) in want["merged"].items()}
    assert out.merged_info.index.name == "cleft_segid"


def test_f_order_single_chunk_tasks():
    """cc_task / cc_overlap_task with an F-ordered chunk (what CloudVolume delivers) == the C-ordered result."""
    import synaptor_b200 as syn
    pred = orc.synth_pred((40, 24, 72), seed=4, sigma=1.5, fg=0.1)
    base = orc.synth_base(pred.shape, seed=5, block=8)
    want = orc.cc_task_c(pred, 0.5, 10)
    got = syn.proc.tasks.cc_overlap_task(np.asfortranarray(pred), np.asfortranarray(base), 0.5, 10)
    assert np.array_equal(got[0], want[0])
    r, c, v, shp = orc.count_overlaps_c(want[0], base)
    assert got[3].shape == shp and np.array_equal(got[3].row, r) and np.array_equal(got[3].col, c)


def _nccl_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        vshape, cshape = (64, 48, 96), (24, 24, 32)          # 3 x 2 x 3 = 18 chunks, ragged, uneven over 4 ranks
        vol = orc.synth_pred(vshape, seed=31, sigma=2.0, fg=0.10)
        base_vol = orc.synth_base(vshape, seed=32, block=6, bg=0.1)
        out = {"rank": rank}
        for graph in (False, True):
            job = run_job(vol, base_vol, cshape, 0.5, 10, 15, graph, world=world, rank=rank)
            out[graph] = {"labels": {ci: job.labels[li].cpu().numpy().view(np.uint32) for li, ci in enumerate(job.mine)},
                          "merged": job.merged_table(all_ranks=True), "own_rows": len(job.merged_table()),
                          "maps": [job.chunk_id_map(ci) for ci in range(job.nchunks)],
                          "ov": {ci: job.overlaps(li) for li, ci in enumerate(job.mine)}, "info": job.info_dict()}
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_volume_job_nccl_ranks():
    """The job over 2+ real GPUs (NCCL all_gather of the packed slots): every rank's chunks and the replicated merge
    results equal the oracle workflow.  Skipped on a single-GPU box."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vshape, cshape = (64, 48, 96), (24, 24, 32)
    vol = orc.synth_pred(vshape, seed=31, sigma=2.0, fg=0.10)
    base_vol = orc.synth_base(vshape, seed=32, block=6, bg=0.1)
    want = oracle_volume_job(vol, base_vol, cshape, 0.5, 10, 15)
    wm = {int(k): [int(x) for x in v] for (k, v) in want["merged"].items()}
    for graph in (False, True):
        seen = set()
        for o in outs:
            r = o[graph]
            assert {int(x[0]): x[1:].tolist() for x in r["merged"]} == wm      # the gathered shares
            assert r["info"]["flags"] == 0
            for ci in range(len(want["chunks"])):
                gi = want["chunks"][ci][0]
                assert r["maps"][ci] == {int(k): int(v) for (k, v) in want["id_maps"][gi].items()}
            for (ci, lab) in r["labels"].items():
                assert np.array_equal(lab, want["final"][want["chunks"][ci][1]])
                rr, cc, vv, shp = r["ov"][ci]
                wr, wc, wv, wshp = want["overlaps"][ci]
                assert shp == wshp and np.array_equal(rr, wr) and np.array_equal(cc, wc) and np.array_equal(vv, wv)
                seen.add(ci)
        assert seen == set(range(len(want["chunks"])))
        assert sum(o[graph]["own_rows"] for o in outs) == len(wm)          # every merged id on exactly one rank


def test_volume_tasks_stream_pipelined():
    """The generator form: several volumes (different data, same shape) in flight; each result == volume_tasks'."""
    import synaptor_b200 as syn
    vshape, cshape = (40, 40, 64), (20, 40, 32)
    vols = []
    for k in range(4):
        vol = orc.synth_pred(vshape, seed=40 + k, sigma=2.0, fg=0.05 + 0.04 * k)
        base_vol = orc.synth_base(vshape, seed=50 + k, block=6, bg=0.1)
        want = oracle_volume_job(vol, base_vol, cshape, 0.5, 10, 15)
        preds = [np.ascontiguousarray(vol[sl]) for (_, sl, _) in want["chunks"]]
        bases = [np.ascontiguousarray(base_vol[sl]) for (_, sl, _) in want["chunks"]]
        vols.append((preds, bases, want))
    outs = list(syn.proc.tasks.volume_tasks_stream(((p, b) for (p, b, _) in vols), vshape, cshape, 0.5, 10, 15,
                                                   rank=0, world=1))
    assert len(outs) == len(vols)
    for (out, (_, _, want)) in zip(outs, vols):
        for (ci, (gi, sl, beg)) in enumerate(want["chunks"]):
            assert np.array_equal(out.ccs[ci], want["final"][sl])
            assert out.chunk_id_maps[ci] == {int(k): int(v) for (k, v) in want["id_maps"][gi].items()}
            wr, wc, wv, wshape = want["overlaps"][ci]
            m = out.overlaps[ci]
            assert m.shape == wshape and np.array_equal(m.row, wr) and np.array_equal(m.col, wc) and np.array_equal(m.data, wv)
        got = {int(i): [int(x) for x in row] for (i, row) in zip(out.merged_info.index, out.merged_info.to_numpy())}
        assert got == {int(k): [int(x) for x in v] for (k, v) in want["merged"].items()}


def test_pageable_large_input_staged():
    """A pageable prediction / base chunk above the staging threshold (8 MB) goes through the threaded pinned staging
    path (C- and F-ordered): same result as the oracle."""
    import synaptor_b200 as syn
    pred = orc.synth_pred((96, 128, 224), seed=8, sigma=2.0, fg=0.05)          # 11 MB float32, 22 MB uint64
    base = orc.synth_base(pred.shape, seed=9, block=16)
    want = orc.cc_task_c(pred, 0.5, 30, with_continuations=False)
    r, c, v, shp = orc.count_overlaps_c(want[0], base)
    for conv in (np.ascontiguousarray, np.asfortranarray):
        got = syn.proc.tasks.cc_overlap_task(conv(pred), conv(base), 0.5, 30)
        assert np.array_equal(got[0], want[0])
        assert got[3].shape == shp and np.array_equal(got[3].row, r) and np.array_equal(got[3].col, c) \
            and np.array_equal(got[3].data, v)


def test_volume_job_deterministic_and_lanes():
    """Two runs of the same job, and a job whose chunk passes alternate over two streams, give identical bytes
    (the unions, the pair table and the merge use atomics in arbitrary order; the results are canonical)."""
    vshape, cshape = (48, 40, 128), (24, 20, 64)
    vol = orc.synth_pred(vshape, seed=61, sigma=1.5, fg=0.12)
    base_vol = orc.synth_base(vshape, seed=62, block=7, bg=0.1)
    import torch
    from synaptor_b200 import multi
    outs = []
    for lanes in (1, 2, 2):
        dev = torch.device("cuda", torch.cuda.current_device())
        job = multi.VolumeJob(vshape, cshape, 0.5, 10, 15, with_base=True, rank=0, world=1, device=dev, graph=True,
                              lanes=lanes)
        keep = []
        for li, ci in enumerate(job.mine):
            beg, shp = job.chunk_rel_begin[ci], job.chunk_shape[ci]
            sl = tuple(slice(b, b + s) for (b, s) in zip(beg, shp))
            keep.append((torch.from_numpy(np.ascontiguousarray(vol[sl])).to(dev),
                         torch.from_numpy(np.ascontiguousarray(base_vol[sl]).view(np.int64)).to(dev)))
            job.set_inputs(li, *keep[-1])
        for _ in range(3):
            job.run()
        assert job.check()
        outs.append(([job.labels[li].cpu().numpy().copy() for li in range(len(job.mine))], job.merged_table().copy(),
                     [tuple(np.asarray(a).copy() for a in job.overlaps(li)[:3]) for li in range(len(job.mine))]))
    want = oracle_volume_job(vol, base_vol, cshape, 0.5, 10, 15)
    for (labels, merged, ovs) in outs:
        for (a, b) in zip(labels, outs[0][0]):
            assert np.array_equal(a, b)
        assert np.array_equal(merged, outs[0][1])
        for (x, y) in zip(ovs, outs[0][2]):
            assert all(np.array_equal(p, q) for (p, q) in zip(x, y))
    for (ci, (gi, sl, beg)) in enumerate(want["chunks"]):
        assert np.array_equal(outs[0][0][ci].view(np.uint32), want["final"][sl])
