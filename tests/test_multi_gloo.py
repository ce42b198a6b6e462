"""
world_size-2 `gloo` test (CPU) of the multi-GPU merge plumbing in synaptor_b200/multi.py:
chunk ownership, global id bases, face / table all_gathers, interior face pairing, LUT
composition.  The CUDA kernels are replaced by a numpy stand-in (tests only); the result
must equal the oracle's merge_ccs_task + remap on the same chunks.
"""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from synaptor_b200 import multi


class NumpyOps:
    """CPU stand-in for multi.DeviceOps with the same contracts (include/synaptor_b200.h)."""

    def read_counts(self, r):
        return r.counts

    def extract_faces(self, labels):
        a = labels.numpy().view(np.uint32)
        faces = [np.take(a, -1 if hi else 0, axis=axis) for (axis, hi) in multi.FACE_ORDER]
        return torch.from_numpy(np.concatenate([f.ravel() for f in faces]).view(np.int32).copy())

    def ids_to_lut(self, table, n, base, lut_len, device):
        lut = np.zeros(lut_len, dtype=np.uint32)
        ids = table[:n, 0].numpy()
        lut[ids] = base + 1 + np.arange(n, dtype=np.uint32)
        return torch.from_numpy(lut.view(np.int32))

    def relabel(self, arr, lut):
        a = arr.numpy().view(np.uint32)
        l = lut.numpy().view(np.uint32)
        ok = a < len(l)
        a[ok] = l[a[ok]]
        return arr

    def face_pair_edges(self, fa, fb, edges, n_edges):
        a, b = fa.numpy().view(np.uint32), fb.numpy().view(np.uint32)
        m = (a != 0) & (b != 0)
        e = np.stack([a[m], b[m]], 1)
        k = int(n_edges[0])
        edges.numpy().view(np.uint32)[k:k + len(e)] = e
        n_edges[0] = k + len(e)

    def union_min_map(self, edges, n_edges, n_ids):
        par = np.arange(n_ids, dtype=np.uint32)

        def find(x):
            while par[x] != x:
                par[x] = par[par[x]]
                x = par[x]
            return x
        for (a, b) in edges.numpy().view(np.uint32)[:int(n_edges[0])]:
            ra, rb = find(a), find(b)
            if ra != rb:
                par[max(ra, rb)] = min(ra, rb)
        return torch.from_numpy(np.array([find(i) for i in range(n_ids)], dtype=np.uint32).view(np.int32))

    def merge_overlaps(self, rows, cols, counts, n_ids):
        r = rows.numpy().view(np.uint32).astype(np.int64)
        c = cols.numpy().view(np.uint64)
        v = counts.numpy()
        tot = {}
        for (a, b, n) in zip(r.tolist(), c.tolist(), v.tolist()):
            if a and n > 0:
                tot[(a, b)] = tot.get((a, b), 0) + n
        arg = np.full(max(n_ids, 1), np.iinfo(np.uint64).max, dtype=np.uint64)
        mx = np.zeros(max(n_ids, 1), dtype=np.int64)
        for ((a, b), n) in sorted(tot.items()):
            if n > mx[a]:
                mx[a], arg[a] = n, b
        return torch.from_numpy(arg.view(np.int64)), torch.from_numpy(mx)

    def merge_tables(self, rows, n_rows, gmap, size_thr):
        g = gmap.numpy().view(np.uint32)
        r = rows.numpy()
        out, fmap = [], np.zeros(n_rows + 1, dtype=np.uint32)
        for d in sorted(set(g[1:].tolist())):
            mem = np.nonzero(g[1:] == d)[0]
            S = int(r[mem, 1].sum())
            if S < size_thr:
                continue
            w = r[mem, 1].astype(np.float64) / np.float64(S)
            cent = [int(sum((r[m, 2 + a] * w[j] for (j, m) in enumerate(mem)), 0.0)) for a in range(3)]
            out.append([d, S, *cent, *r[mem, 5:8].min(0).tolist(), *r[mem, 8:11].max(0).tolist()])
            fmap[mem + 1] = d
        merged = torch.tensor(out, dtype=torch.int64).reshape(-1, 11)
        return merged, torch.from_numpy(fmap.view(np.int32)), torch.tensor([len(out)], dtype=torch.int64)


def make_chunks(vol, cshape, thresh, dust):
    grid = tuple(-(-s // c) for (s, c) in zip(vol.shape, cshape))
    chunks = []
    for gi in np.ndindex(grid):
        beg = tuple(a * b for (a, b) in zip(gi, cshape))
        sl = tuple(slice(b, b + c) for (b, c) in zip(beg, cshape))
        ccs, conts, info = orc.cc_task_c(np.ascontiguousarray(vol[sl]), thresh, dust, offset=beg)
        table = np.concatenate([np.asarray(info.index, dtype=np.int64)[:, None], info.to_numpy(dtype=np.int64)], 1)
        chunks.append((gi, sl, ccs, conts, info, table.reshape(-1, 11)))
    return grid, chunks


def to_result(ccs, table, base=None):
    res = SimpleNamespace(labels=torch.from_numpy(ccs.view(np.int32).copy()),
                          seg_table=torch.from_numpy(table.copy()),
                          counts={"kept": len(table), "max_label": int(ccs.max())})
    if base is not None:
        r, c, v, _ = orc.count_overlaps_c(ccs, base)
        res.pair_rows = torch.from_numpy(r.astype(np.uint32).view(np.int32).copy())
        res.pair_cols = torch.from_numpy(c.astype(np.uint64).view(np.int64).copy())
        res.pair_vals = torch.from_numpy(v.astype(np.int64).copy())
        res.counts["pairs"] = len(r)
    return res


def oracle_max_overlaps(vshape, chunks, id_maps, base):
    """merge_overlaps_task on the remapped chunk volumes: {merged id: (base id, count)}."""
    tot = {}
    for (gi, sl, ccs, _, _, _) in chunks:
        final = orc.remap_ids(ccs.copy(), id_maps[gi])
        r, c, v, _ = orc.count_overlaps_c(final, np.ascontiguousarray(base[sl]))
        for (a, b, n) in zip(r.tolist(), c.tolist(), v.tolist()):
            tot[(a, b)] = tot.get((a, b), 0) + n
    best = {}
    for ((a, b), n) in sorted(tot.items()):          # (row, col) order of the consolidated matrix; first strict max
        if a not in best or n > best[a][1]:
            best[a] = (b, n)
    return best


def oracle_merge(grid, chunks, size_thr, face_shape):
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    for (gi, _, _, conts, info, _) in chunks:
        cont_arr[gi], info_arr[gi] = conts, info
    merged, id_maps = orc.merge_ccs_task(cont_arr, info_arr, size_thr, face_shape)
    return merged, id_maps


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = orc.synth_pred((32, 24, 40), seed=21, sigma=2.0, fg=0.12)
        cshape = (16, 12, 20)
        grid, chunks = make_chunks(vol, cshape, 0.5, 6)
        owners = multi.chunk_owners(grid, world)
        mine = [i for i in range(len(chunks)) if owners[i] == rank]
        base = orc.synth_base(vol.shape, seed=22, block=6, bg=0.1)
        results = [to_result(chunks[i][2], chunks[i][5], np.ascontiguousarray(base[chunks[i][1]])) for i in mine]
        res = multi.merge_ccs_device(results, grid, rank=rank, world=world, size_thr=10, ops=NumpyOps())
        mo = multi.merge_overlaps_device(results, res, rank=rank, world=world, ops=NumpyOps())
        out = {"rank": rank, "merged": res.merged_table.numpy(), "labels": {i: r.labels.numpy().view(np.uint32)
                                                                            for (i, r) in zip(mine, results)},
               "base": res.id_base, "n_edges": res.n_edges, "max_overlaps": mo}
        q.put(out)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_merge_plumbing_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    vol = orc.synth_pred((32, 24, 40), seed=21, sigma=2.0, fg=0.12)
    grid, chunks = make_chunks(vol, (16, 12, 20), 0.5, 6)
    merged, id_maps = oracle_merge(grid, chunks, 10, (16, 20))
    for o in outs:
        got = {int(r[0]): r[1:].tolist() for r in o["merged"]}
        assert set(got) == set(merged)
        for (k, row) in merged.items():
            assert got[k][0] == row[0] and got[k][4:] == row[4:]
            assert all(abs(a - b) <= 1 for (a, b) in zip(got[k][1:4], row[1:4]))
    # both ranks computed the same table; every chunk volume carries the merged global ids
    assert np.array_equal(outs[0]["merged"], outs[1]["merged"])
    labels = {}
    for o in outs:
        labels.update(o["labels"])
    for (i, (gi, sl, ccs, _, _, _)) in enumerate(chunks):
        want = orc.remap_ids(ccs.copy(), id_maps[gi])
        assert np.array_equal(labels[i], want)
    # merge_overlaps: every rank holds the same arg-max table, equal to the oracle's on the remapped volumes
    base = orc.synth_base(vol.shape, seed=22, block=6, bg=0.1)
    want_mo = oracle_max_overlaps(vol.shape, chunks, id_maps, base)
    for o in outs:
        ids, cols, vals = o["max_overlaps"]
        got_mo = {int(i): (int(c), int(v)) for (i, c, v) in zip(ids, cols, vals)}
        assert got_mo == want_mo


def test_plan_helpers():
    assert multi.chunk_owners((2, 2, 2), 4) == [0, 1, 2, 3, 0, 1, 2, 3]
    base, total = multi.id_bases([3, 0, 5])
    assert base.tolist() == [0, 3, 3] and total == 8
    pairs = multi.interior_face_pairs((2, 1, 2))
    assert pairs == [(0, 0, 2, 1), (0, 4, 1, 5), (1, 0, 3, 1), (2, 4, 3, 5)]
    offs, sizes, shapes = multi.face_layout((4, 5, 6))
    assert sizes == [30, 30, 24, 24, 20, 20] and offs[2] == 60 and shapes[4] == (4, 5)
