#!/usr/bin/env python
"""
Randomised parity sweep against the oracle (not part of the test suite: run by hand on a GPU box).
Random shapes (biased to nz % 8 == 0 and nz % 128 == 0, the sector / TMA paths), densities, connectivities,
dust thresholds, with and without a base segmentation / dilation / overlap_seg.

  python tools/fuzz_parity.py [cases=200] [seed=0] [max voxels per case=300000]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as orc
import synaptor_b200 as syn

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
maxvox = int(sys.argv[3]) if len(sys.argv) > 3 else 300000      # bigger volumes span several tiles of every kernel
T = syn.proc.tasks
bad = 0
for case in range(ncases):
    nz = int(rng.choice([8, 16, 64, 96, 128, 128, 256, 384, 512, 40, 100, 131]))
    nx, ny = int(rng.integers(1, 40 if maxvox <= 300000 else 96)), int(rng.integers(1, 70 if maxvox <= 300000 else 160))
    while nx * ny * nz > maxvox:
        nx, ny = max(1, nx // 2), max(1, ny // 2)
    shape = (nx, ny, nz)
    p = float(rng.choice([0.01, 0.05, 0.2, 0.5, 0.8, 0.99]))
    smooth = rng.random() < 0.5
    pred = orc.synth_pred(shape, seed=int(rng.integers(1 << 30)), sigma=float(rng.choice([1.0, 2.0])), fg=p) if smooth \
        else rng.random(shape, dtype=np.float32) + np.float32(p - 0.5)
    dust = int(rng.choice([0, 1, 5, 50]))
    off = tuple(int(v) for v in rng.integers(-50, 50, size=3))
    mode = rng.choice(["cc", "cc26", "cc18", "fused", "dil", "ml", "host", "utils", "merge"],
                      p=[0.15, 0.08, 0.04, 0.25, 0.08, 0.1, 0.1, 0.1, 0.1])
    base = orc.synth_base(shape, seed=int(rng.integers(1 << 30)), block=int(rng.choice([1, 2, 4, 9])), bg=0.2)
    try:
        if mode == "fused":
            got = T.cc_overlap_task(pred, torch.from_numpy(base.view(np.int64)).cuda(), 0.5, dust, offset=off)
            want = orc.cc_task_c(pred, 0.5, dust, offset=off)
            r, c, v, shp = orc.count_overlaps_c(want[0], base)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[2].to_numpy(), want[2].to_numpy()) and \
                got[3].shape == shp and np.array_equal(got[3].row, r) and np.array_equal(got[3].col, c) and \
                np.array_equal(got[3].data, v)
        elif mode == "host":       # host arrays in: pipelined copies + standalone overlap counter; F-order prediction
            pf = np.asfortranarray(pred) if rng.random() < 0.5 else pred
            got = T.cc_overlap_task(pf, base, 0.5, dust, offset=off)
            want = orc.cc_task_c(pred, 0.5, dust, offset=off)
            r, c, v, shp = orc.count_overlaps_c(want[0], base)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[2].to_numpy(), want[2].to_numpy()) and \
                got[3].shape == shp and np.array_equal(got[3].row, r) and np.array_equal(got[3].col, c) and \
                np.array_equal(got[3].data, v)
        elif mode == "utils":      # seg_utils on an arbitrary label volume
            seg = orc.cc_task_c(pred, 0.5, 0)[0]
            seg[seg != 0] = (seg[seg != 0] * 7919) % 100003 + 1          # non-contiguous ids
            su = syn.seg_utils
            sz = su.segment_sizes(seg)
            wsz = orc.segment_sizes_np(seg)
            ok = {int(k): int(v) for (k, v) in sz.items()} == {int(k): int(v) for (k, v) in wsz.items()}
            cm, wcm = su.centers_of_mass(seg, offset=off), orc.centers_of_mass_np(seg, offset=off)
            ok = ok and {int(k): tuple(int(x) for x in v) for (k, v) in cm.items()} == \
                {int(k): tuple(int(x) for x in v) for (k, v) in wcm.items()}
            bb, wbb = su.bounding_boxes(seg, offset=off), orc.bounding_boxes_np(seg, offset=off)
            ok = ok and {int(k): tuple(v.astuple()) for (k, v) in bb.items()} == \
                {int(k): tuple(int(x) for x in v) for (k, v) in wbb.items()}
            fs, fsz = su.filter_segs_by_size(seg, max(dust, 1))
            wfs, wfsz = orc.filter_segs_by_size_np(seg, max(dust, 1))
            ok = ok and np.array_equal(fs, wfs)
        elif mode == "merge":      # chunked pipeline + device merge + remap vs the oracle's merge_ccs_task
            from synaptor_b200 import _device as D
            from synaptor_b200 import multi
            cs = tuple(max(1, s // int(rng.choice([1, 2]))) for s in shape)
            vs = tuple((s // c) * c for (s, c) in zip(shape, cs))
            vol = np.ascontiguousarray(pred[:vs[0], :vs[1], :vs[2]])
            grid = tuple(v // c for (v, c) in zip(vs, cs))
            cont_arr, info_arr = np.empty(grid, dtype=object), np.empty(grid, dtype=object)
            results, chunks = [], []
            for gi in np.ndindex(grid):
                beg = tuple(a * b for (a, b) in zip(gi, cs))
                sl = tuple(slice(b, b + c) for (b, c) in zip(beg, cs))
                sub = np.ascontiguousarray(vol[sl])
                results.append(D.chunk_ccs_device(torch.from_numpy(sub).cuda(), np.float32(0.5), dust, offset=beg))
                ccs, conts, info = orc.cc_task_c(sub, 0.5, dust, offset=beg)
                chunks.append((gi, sl, ccs))
                cont_arr[gi], info_arr[gi] = conts, info
            fs2 = tuple(sorted(cs)[1:])
            merged, id_maps = orc.merge_ccs_task(cont_arr, info_arr, dust + 3, (max(cs), max(cs)))
            out = multi.merge_ccs_device(results, grid, rank=0, world=1, size_thr=dust + 3)
            gotm = {int(r[0]): r[1:].tolist() for r in out.merged_table.cpu().numpy()}
            ok = set(gotm) == set(merged) and all(gotm[k][0] == merged[k][0] and gotm[k][4:] == merged[k][4:]
                                                  for k in merged)
            for ((gi, sl, ccs), res) in zip(chunks, results):
                ok = ok and np.array_equal(res.labels_numpy(), orc.remap_ids(ccs.copy(), id_maps[gi]))
        elif mode == "ml":
            got = T.cc_task(pred, 0.5, dust, offset=off, overlap_seg=base, verbose=False)
            want = orc.cc_task_c(pred, 0.5, dust, offset=off, overlap_seg=base)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[2].to_numpy().astype(np.int64),
                                                                    want[2].to_numpy().astype(np.int64))
        else:
            conn = {"cc26": 26, "cc18": 18}.get(mode, 6)
            dil = int(rng.choice([1, 2, 5])) if mode == "dil" else 0
            got = T.cc_task(pred, 0.5, dust, offset=off, connectivity=conn, dil_param=dil, verbose=False)
            want = orc.cc_task_c(pred, 0.5, dust, offset=off, connectivity=conn, dil_param=dil)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[2].to_numpy(), want[2].to_numpy())
    except Exception as e:      # noqa: BLE001
        ok = False
        print("EXC", repr(e)[:200])
    if not ok:
        bad += 1
        print("MISMATCH", case, mode, shape, p, smooth, dust, off)
print(f"fuzz: {ncases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
