"""
Parity of the CUDA path (through the C-ABI) against the oracle and the committed
golden fixtures of the executed reference.  Bit-exact: labels, sizes, bboxes,
integer centroids, overlap triples and their order.  Needs a B200.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from util import flatten_conts, gold, table_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def syn():
    import synaptor_b200
    return synaptor_b200


def first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    d = np.argwhere(a != b)
    if len(d) == 0:
        return "equal"
    i = tuple(d[0])
    return f"{len(d)} diffs; first at {i}: got {a[i]} want {b[i]}; max got {a.max()} want {b.max()}"


def check_cc_task(syn, pred, thresh, dust, offset=(0, 0, 0), conn=6, dil=0):
    want = orc.cc_task_c(pred, thresh, dust, offset=offset, connectivity=conn, dil_param=dil)
    got = syn.proc.tasks.cc_task(pred, thresh, dust, offset=offset, connectivity=conn, dil_param=dil, verbose=False)
    assert got[0].dtype == np.uint32 and got[0].flags.c_contiguous
    assert np.array_equal(got[0], want[0]), first_diff(got[0], want[0])
    gm, gc = flatten_conts(got[1])
    wm, wc = flatten_conts(want[1])
    assert np.array_equal(gm, wm), first_diff(gm, wm)
    assert np.array_equal(gc, wc)
    gi, gt = table_of(got[2])
    wi, wt = table_of(want[2])
    assert list(got[2].columns) == list(want[2].columns) and got[2].index.name == "cleft_segid"
    assert np.array_equal(gi, wi), first_diff(gi, wi)
    assert np.array_equal(gt, wt), first_diff(gt, wt)
    return got


# ----------------------------------------------------------------------------- goldens
def test_g1_cc_task_golden(syn):
    g = gold("g1_cc_task.npz")
    ccs, conts, info = syn.proc.tasks.cc_task(g["pred"], float(g["thresh"]), int(g["dust"]),
                                              offset=tuple(g["offset"].tolist()), verbose=False)
    assert np.array_equal(ccs, g["ccs"]), first_diff(ccs, g["ccs"])
    meta, coords = flatten_conts(conts)
    assert np.array_equal(meta, g["cont_meta"]) and np.array_equal(coords, g["cont_coords"])
    idx, tab = table_of(info)
    assert list(info.columns) == g["info_columns"].tolist()
    assert np.array_equal(idx, g["info_index"].astype(np.int64))
    assert np.array_equal(tab, g["info_table"].astype(np.int64)), first_diff(tab, g["info_table"].astype(np.int64))


def test_g2_overlap_golden(syn):
    g = gold("g2_overlap_task.npz")
    ov = syn.proc.tasks.overlap_task(g["ccs"], g["base"])
    assert ov.shape == tuple(g["shape"].tolist())
    assert np.array_equal(ov.row, g["row"]), first_diff(ov.row, g["row"])
    assert np.array_equal(ov.col, g["col"]) and np.array_equal(ov.data, g["data"])
    assert ov.data.dtype == np.int64
    m = gold("g2_max_overlaps.npz")
    mo = syn.proc.overlap.find_max_overlaps(ov)
    assert np.array_equal(mo.row, m["row"]) and np.array_equal(mo.col, m["col"]) and np.array_equal(mo.data, m["data"])
    mat, ids1, ids2 = syn.seg_utils.count_overlaps(g["ccs"], g["base"], orig_ids=False)
    assert np.array_equal(ids1, np.unique(g["ccs"])[1:]) and np.array_equal(ids2, np.unique(g["base"])[1:])
    assert np.array_equal(ids1[mat.row], g["row"]) and np.array_equal(ids2[mat.col].astype(np.int64), g["col"])


def test_g3_dilated_golden(syn):
    g = gold("g3_dilated.npz")
    for k in (1, 2, 5):
        ccs = syn.dilated_components(g["pred"], k, float(g["thresh"]))
        assert np.array_equal(ccs, g[f"ccs_k{k}"]), (k, first_diff(ccs, g[f"ccs_k{k}"]))
        dil = syn.seg_utils.dilate_by_k(g["pred"] > float(g["thresh"]), k)
        assert dil.dtype == np.uint32 and np.array_equal(dil, g[f"dil_k{k}"]), (k, first_diff(dil, g[f"dil_k{k}"]))


def test_g4_merge_golden(syn):
    g = gold("g4_merge_ccs.npz")
    vol = g["vol"]
    cs = tuple(g["chunk_shape"].tolist())
    boxes = syn.chunk_bboxes(vol.shape, cs)
    grid = (2, 2, 2)
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    segs = np.empty(grid, dtype=object)
    for (i, (gi, bb)) in enumerate(zip(np.ndindex(grid), boxes)):
        ccs, conts, info = syn.proc.tasks.cc_task(np.ascontiguousarray(vol[bb.index()]), float(g["thresh"]),
                                                  int(g["dust"]), offset=tuple(bb.min()), verbose=False)
        t = g[f"chunk_table_{i}"]
        idx, tab = table_of(info)
        assert np.array_equal(idx, t[:, 0]) and np.array_equal(tab, t[:, 1:])
        assert np.array_equal(ccs, g[f"chunk_seg_{i}"])
        cont_arr[gi], info_arr[gi], segs[gi] = conts, info, ccs
    merged, id_maps = syn.proc.tasks.merge_ccs_task(cont_arr, info_arr, int(g["size_thr"]),
                                                    tuple(g["max_face_shape"].tolist()))
    gold_rows = {int(i): r.astype(np.int64) for (i, r) in zip(g["merged_index"], g["merged_table"])}
    midx, mtab = table_of(merged)
    assert sorted(midx.tolist()) == sorted(gold_rows)
    for (k, row) in zip(midx.tolist(), mtab):
        gr = gold_rows[k]
        assert row[0] == gr[0] and np.array_equal(row[4:], gr[4:])
        assert np.all(np.abs(row[1:4] - gr[1:4]) <= 1)   # float64 weighted mean, truncated (1e-6 rel. tolerance)
    final = np.zeros(vol.shape, dtype=np.uint32)
    for (i, (gi, bb)) in enumerate(zip(np.ndindex(grid), boxes)):
        assert sorted((int(a), int(b)) for (a, b) in id_maps[gi].items()) == [tuple(x) for x in g[f"id_map_{i}"].tolist()]
        final[bb.index()] = syn.proc.tasks.remap_ids_task(segs[gi].copy(), id_maps[gi], copy=False)
    assert np.array_equal(final, g["final"]), first_diff(final, g["final"])


def test_g4_merge_through_files(syn, tmp_path):
    """The same merge with a processing directory between the stages, as tasks_w_io.py:54-65, 112-138 runs it: chunk
    tasks write seg-info CSVs and HDF5 continuation files, the merge task reads the directory back."""
    g = gold("g4_merge_ccs.npz")
    vol = g["vol"]
    boxes = syn.chunk_bboxes(vol.shape, tuple(g["chunk_shape"].tolist()))
    taskio = syn.proc.io
    d = str(tmp_path)
    for bb in boxes:
        _, conts, info = syn.proc.tasks.cc_task(np.ascontiguousarray(vol[bb.index()]), float(g["thresh"]),
                                                int(g["dust"]), offset=tuple(bb.min()), verbose=False)
        taskio.write_chunk_continuations(conts, d, bb)
        taskio.write_chunk_seg_info(info, d, bb)
    cont_arr, _ = taskio.read_all_continuations(d)
    info_arr, _ = taskio.read_all_chunk_seg_infos(d)
    assert cont_arr.shape == info_arr.shape == (2, 2, 2)
    merged, id_maps = syn.proc.tasks.merge_ccs_task(cont_arr, info_arr, int(g["size_thr"]),
                                                    tuple(g["max_face_shape"].tolist()))
    gold_rows = {int(i): r.astype(np.int64) for (i, r) in zip(g["merged_index"], g["merged_table"])}
    midx, mtab = table_of(merged)
    assert sorted(midx.tolist()) == sorted(gold_rows)
    for (k, row) in zip(midx.tolist(), mtab):
        assert row[0] == gold_rows[k][0] and np.array_equal(row[4:], gold_rows[k][4:])
    for (i, gi) in enumerate(np.ndindex(2, 2, 2)):
        assert sorted((int(a), int(b)) for (a, b) in id_maps[gi].items()) == [tuple(x) for x in g[f"id_map_{i}"].tolist()]


def test_tasks_w_io_local_workflow(syn, tmp_path, capsys):
    """The five task wrappers of proc/tasks_w_io.py (mirror of synaptor/proc/tasks_w_io.py:22-143, 503-594) run as the
    reference's workflow runs them -- cc_task per chunk, merge_ccs_task, remap_ids_task + overlap_task per chunk,
    merge_overlaps_task -- with nothing but files between the stages (F-ordered .npy volumes, a processing
    directory): final volume, merged table, id maps, chunk overlap matrices and max overlaps equal the executed
    reference's (fixtures g4, g13)."""
    import os
    import pandas as pd
    g, g13 = gold("g4_merge_ccs.npz"), gold("g13_merge_overlaps.npz")
    W, taskio = syn.proc.tasks_w_io, syn.proc.io
    d = str(tmp_path / "proc")
    pred_p, seg_p, out_p, base_p = (str(tmp_path / n) for n in ("pred.npy", "seg.npy", "seg_final.npy", "base.npy"))
    np.save(pred_p, np.asfortranarray(g["vol"]))              # what CloudVolume hands out is F-ordered
    np.save(base_p, np.asfortranarray(g13["base"]))
    shape = g["vol"].shape
    syn.io.init_seg_volume(seg_p, vol_size=shape)
    syn.io.init_seg_volume(out_p, vol_size=shape)
    boxes = syn.chunk_bboxes(shape, tuple(g["chunk_shape"].tolist()))
    for bb in boxes:
        W.cc_task(pred_p, seg_p, d, float(g["thresh"]), int(g["dust"]), tuple(bb.min()), tuple(bb.max()), timing_tag="t")
    assert "Running connected components" in capsys.readouterr().out
    for (i, bb) in enumerate(boxes):
        assert np.array_equal(syn.io.read_cloud_volume_chunk(seg_p, bb), g[f"chunk_seg_{i}"])
    assert len(os.listdir(os.path.join(d, "continuations"))) == 6 * len(boxes)
    W.merge_ccs_task(d, int(g["size_thr"]), tuple(g["max_face_shape"].tolist()))
    merged = taskio.read_merged_seg_info(d)
    gold_rows = {int(i): r.astype(np.int64) for (i, r) in zip(g["merged_index"], g["merged_table"])}
    midx, mtab = table_of(merged)
    assert sorted(midx.tolist()) == sorted(gold_rows)
    for (k, row) in zip(midx.tolist(), mtab):
        assert row[0] == gold_rows[k][0] and np.array_equal(row[4:], gold_rows[k][4:])
    for (i, bb) in enumerate(boxes):
        idm = taskio.read_chunk_id_map(d, bb)
        assert sorted((int(a), int(b)) for (a, b) in idm.items()) == [tuple(x) for x in g[f"id_map_{i}"].tolist()]
        W.remap_ids_task(seg_p, out_p, tuple(bb.min()), tuple(bb.max()), d)
        W.overlap_task(out_p, base_p, tuple(bb.min()), tuple(bb.max()), d)
    final = np.load(out_p)
    assert np.array_equal(final, g["final"]), first_diff(final, g["final"])
    for (i, bb) in enumerate(boxes):
        m = taskio.read_chunk_overlap_mat(taskio.chunk_overlap_fname(d, bb))
        got = sorted(zip(m.row.tolist(), m.col.tolist(), m.data.tolist()))
        assert got == sorted(tuple(int(v) for v in t) for t in g13[f"ov_{i}"]), i
    W.merge_overlaps_task(d)
    mo = pd.read_csv(os.path.join(d, "max_overlaps.df"), index_col=0)
    cn = syn.proc.colnames
    got = sorted(zip(mo[cn.rows].tolist(), mo[cn.cols].tolist(), mo[cn.vals].tolist()))
    assert got == sorted(zip(g13["max_row"].tolist(), g13["max_col"].tolist(), g13["max_val"].tolist()))
    assert taskio.read_task_timing(d, "ccs", "t") > 0


def test_g9_hashed_merge_golden(syn):
    """Merge sharded by hash (tasks.py:125-180) against the executed reference: same edges, map, buckets, tables."""
    g, g4 = gold("g9_hashed_merge.npz"), gold("g4_merge_ccs.npz")
    cn, tasks, Face = syn.proc.colnames, syn.proc.tasks, syn.proc.seg.Face
    vol = g4["vol"]
    boxes = syn.chunk_bboxes(vol.shape, tuple(g4["chunk_shape"].tolist()))
    grid = (2, 2, 2)
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    for (gi, bb) in zip(np.ndindex(grid), boxes):
        _, cont_arr[gi], info_arr[gi] = tasks.cc_task(np.ascontiguousarray(vol[bb.index()]), 0.5, 10,
                                                      offset=tuple(bb.min()), verbose=False)
    full, maps = syn.proc.seg.merge.assign_unique_ids_serial(info_arr)
    mfs = tuple(g4["max_face_shape"].tolist())
    edges = []
    for gi in np.ndindex(grid):
        for axis in range(3):
            if gi[axis] == grid[axis] - 1:
                continue
            gj = tuple(v + (a == axis) for (a, v) in enumerate(gi))
            edges.extend(tasks.match_continuations_task(cont_arr[gi][Face(axis, True)], cont_arr[gj][Face(axis, False)],
                                                        max_face_shape=mfs, id_map1=maps[gi], id_map2=maps[gj]))
    assert sorted(set((int(a), int(b)) for (a, b) in edges)) == [tuple(e) for e in g["edges"].tolist()]
    hashmax = int(g["hashmax"])
    mdf = tasks.seg_graph_cc_task(edges, hashmax, list(full.index))
    assert list(mdf.columns) == [cn.src_id, cn.dst_id, cn.dst_id_hash]
    got = {int(a): (int(b), int(h)) for (a, b, h) in zip(mdf[cn.src_id], mdf[cn.dst_id], mdf[cn.dst_id_hash])}
    want = {int(a): (int(b), int(h)) for (a, b, h) in zip(g["map_src"], g["map_dst"], g["map_hash"])}
    assert got == want
    full[cn.dst_id] = full.index.map({k: v[0] for (k, v) in got.items()})
    gold_rows = {int(i): (int(h), r.astype(np.int64)) for (i, h, r) in
                 zip(g["merged_index"], g["merged_hash"], g["merged_table"])}
    seen, dropped = set(), []
    for h in range(hashmax):
        part = full[[got[int(i)][1] == h for i in full.index]].copy()
        if len(part) == 0:
            continue
        merged, szmap = tasks.merge_seginfo_task(part, szthresh=int(g["size_thr"]))
        assert all(v == 0 for v in szmap.values())
        dropped.extend(int(k) for k in szmap)
        midx, mtab = table_of(merged)
        for (k, row) in zip(midx.tolist(), mtab):
            gh, gr = gold_rows[k]
            assert gh == h and row[0] == gr[0] and np.array_equal(row[4:], gr[4:])
            assert np.all(np.abs(row[1:4] - gr[1:4]) <= 1)
            seen.add(k)
    assert seen == set(gold_rows) and sorted(dropped) == g["dropped"].tolist()
    assert tasks.merge_seginfo_task(full.copy())[1] is None


def test_g10_seg_utils_golden(syn):
    """The standalone seg_utils calls of the path against outputs of the executed reference."""
    g = gold("g10_seg_utils.npz")
    seg, base = g["seg"], g["base"]
    su = syn.seg_utils
    ids = g["ids"].tolist()
    szs = su.segment_sizes(seg)
    assert szs == dict(zip(ids, g["sizes"].tolist())) and list(szs) == ids
    assert su.centers_of_mass(seg, offset=(3, 4, 5)) == {i: tuple(c) for (i, c) in zip(ids, g["coms"].tolist())}
    bb = su.bounding_boxes(seg, offset=(3, 4, 5))
    assert {k: v.astuple() for (k, v) in bb.items()} == {i: tuple(b) for (i, b) in zip(ids, g["bboxes"].tolist())}
    assert np.array_equal(su.nonzero_unique_ids(seg), g["uniq"])
    filt, rem = su.filter_segs_by_size(seg, 30, to_ignore=[1, 900])
    assert np.array_equal(filt, g["filt"]) and rem == dict(zip(g["filt_ids"].tolist(), g["filt_sizes"].tolist()))
    assert filt is not seg and np.array_equal(su.filter_segs_by_id(seg, [2, 900, 5], copy=True), g["filt_by_id"])
    mapping = dict(zip(g["map_keys"].tolist(), g["map_vals"].tolist()))
    assert np.array_equal(su.relabel_data(seg, mapping, copy=True), g["relabelled"])
    mat, r_ids, c_ids = su.count_overlaps(seg, base)
    assert np.array_equal(r_ids, g["ov_row_ids"]) and np.array_equal(c_ids, g["ov_col_ids"])
    assert mat.shape == tuple(g["ov_shape"].tolist())
    assert np.array_equal(mat.row, g["ov_row"]) and np.array_equal(mat.col, g["ov_col"])
    assert np.array_equal(mat.data, g["ov_val"])
    mo, _, _ = su.count_overlaps(seg, base, orig_ids=True)
    assert mo.shape == tuple(int(x) for x in g["ovo_shape"])
    assert np.array_equal(mo.row, g["ovo_row"]) and np.array_equal(np.asarray(mo.col).astype(np.uint64), g["ovo_col"])
    assert np.array_equal(mo.data, g["ovo_val"])


def test_g5_centroids_golden(syn):
    g = gold("g5_centroids.npz")
    c = syn.centers_of_mass(g["seg"])
    assert sorted(c) == g["ids"].tolist()
    assert np.array_equal(np.array([c[i] for i in sorted(c)]), g["cents"])
    big = np.ones(tuple(g["big_shape"].tolist()), dtype=np.uint32)
    assert syn.centers_of_mass(big)[1] == tuple(g["big_cent"].tolist())
    assert syn.centers_of_mass(g["seg"], offset=(1, 2, 3))[4] == tuple(int(a + b) for (a, b) in zip(c[4], (1, 2, 3)))


def test_g6_threshold_golden(syn):
    g = gold("g6_threshold.npz")
    ccs = syn.connected_components(g["d"], float(g["thresh"]))
    assert np.array_equal(ccs, g["ccs"]), (ccs, g["ccs"])


def test_g7_cc_task_overlap_seg_golden(syn):
    g = gold("g7_cc_task_overlap_seg.npz")
    ccs, conts, info = syn.proc.tasks.cc_task(g["pred"], float(g["thresh"]), int(g["dust"]),
                                              offset=tuple(g["offset"].tolist()), overlap_seg=g["base"], verbose=False)
    assert np.array_equal(ccs, g["ccs"]), first_diff(ccs, g["ccs"])
    meta, coords = flatten_conts(conts)
    assert np.array_equal(meta, g["cont_meta"]) and np.array_equal(coords, g["cont_coords"])
    assert list(info.columns) == g["info_columns"].tolist() and info.index.name == "cleft_segid"
    assert np.array_equal(np.asarray(info.index, dtype=np.int64), g["info_index"])
    assert np.array_equal(info.to_numpy().astype(np.int64), g["info_table"])
    lab = syn.proc.seg.connected_components(g["pred"], float(g["thresh"]), overlap_seg=g["base"])
    want = orc.connected_components(g["pred"], float(g["thresh"]), overlap_seg=g["base"])
    assert np.array_equal(lab, want), first_diff(lab, want)


@pytest.mark.parametrize("shape", [(1, 1, 70), (3, 5, 7), (9, 11, 64), (16, 8, 129), (20, 24, 128), (12, 10, 260)])
@pytest.mark.parametrize("block", [1, 3, 16])
def test_multilabel_noise_vs_oracle(syn, shape, block):
    """Multi-label labelling (overlap_seg) on noise at several densities; block = 1 makes nearly every voxel its
    own base segment, block = 16 leaves long runs that break only rarely."""
    rng = np.random.default_rng(hash((shape, block)) % (2 ** 32))
    for p in (0.1, 0.5, 0.9, 1.0):
        pred = rng.random(shape, dtype=np.float32)
        base = orc.synth_base(shape, seed=int(p * 10) + block, block=block, bg=0.2)
        want = orc.cc_task_c(pred, 1.0 - p, 3, offset=(1, 2, 3), overlap_seg=base)
        got = syn.proc.tasks.cc_task(pred, 1.0 - p, 3, offset=(1, 2, 3), overlap_seg=base, verbose=False)
        assert np.array_equal(got[0], want[0]), first_diff(got[0], want[0])
        assert list(got[2].columns) == list(want[2].columns)
        assert np.array_equal(np.asarray(got[2].index, dtype=np.int64), np.asarray(want[2].index, dtype=np.int64))
        assert np.array_equal(got[2].to_numpy().astype(np.int64), want[2].to_numpy().astype(np.int64))


# ----------------------------------------------------------------------------- random volumes vs the oracle
SHAPES = [(1, 1, 1), (1, 1, 40), (3, 5, 7), (2, 3, 33), (9, 11, 64), (16, 8, 129), (7, 13, 131),
          (20, 24, 128), (33, 31, 100), (12, 10, 260), (40, 40, 36)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("conn", [6, 18, 26])
def test_random_noise_vs_oracle(syn, shape, conn):
    rng = np.random.default_rng(hash((shape, conn)) % (2 ** 32))
    for p in (0.05, 0.3, 0.5, 0.8, 1.0):
        pred = rng.random(shape, dtype=np.float32)
        check_cc_task(syn, pred, 1.0 - p, dust=3, conn=conn, offset=(7, -2, 100))


# shapes with nz % 8 == 0 take the streaming / sector path (k_stream, k_sector_labels); (16, 32, 512) holds
# exactly one 2048-word tile of the shared-memory labelling pass, the larger ones several tiles and tile borders
# nz % 128 == 0 additionally takes the TMA stage ring of k_stream; (3, 7, 128) and (5, 13, 384) end in a partial
# quarter tile (21 and 195 chunks), (20, 40, 256) in a half tile
SECTOR_SHAPES = [(16, 32, 512), (20, 40, 256), (9, 70, 96), (33, 17, 64), (5, 6, 8), (40, 24, 1032), (3, 7, 128),
                 (5, 13, 384)]


@pytest.mark.parametrize("shape", SECTOR_SHAPES)
@pytest.mark.parametrize("conn", [6, 26])
def test_sector_path_noise_vs_oracle(syn, shape, conn):
    """0.5 noise on a full tile gives > 4096 runs per tile (tile left to the global union pass), 0.06 noise gives
    > 512 tile-local components (statistics left to the global pass): both fallbacks of k_tile_ccl are exercised
    and reported through the counts."""
    from synaptor_b200 import _device as D
    import torch
    rng = np.random.default_rng(hash((shape, conn, "sector")) % (2 ** 32))
    flagged = 0
    for p in (0.02, 0.06, 0.25, 0.5, 0.97):
        pred = rng.random(shape, dtype=np.float32)
        check_cc_task(syn, pred, 1.0 - p, dust=2, conn=conn, offset=(3, 5, -7))
        res = D.chunk_ccs_device(torch.from_numpy(pred).cuda(), np.float32(1.0 - p), 2, connectivity=conn)
        flagged += res.counts["flagged_tiles"]
    if shape == (16, 32, 512):
        assert flagged > 0, "the dense / many-roots fallbacks were never taken"


@pytest.mark.parametrize("shape", [(16, 32, 512), (24, 40, 128), (3, 7, 128), (5, 13, 384), (9, 70, 96)])
def test_sector_path_fused_overlap_noise(syn, shape):
    rng = np.random.default_rng(99)
    for p in (0.06, 0.5):
        pred = rng.random(shape, dtype=np.float32)
        base = orc.synth_base(shape, seed=3, block=4)
        want_ccs, _, want_info = orc.cc_task_c(pred, 1.0 - p, 2)
        r, c, v, shp = orc.count_overlaps_c(want_ccs, base)
        import torch
        base_dev = torch.from_numpy(base.view(np.int64)).cuda()       # device-resident base: the fused device pass
        for b in (base_dev, base):                                    # host base: pipelined two-call path
            ccs, conts, info, mat = syn.proc.tasks.cc_overlap_task(pred, b, 1.0 - p, 2)
            assert np.array_equal(ccs, want_ccs), first_diff(ccs, want_ccs)
            assert np.array_equal(info.to_numpy(), want_info.to_numpy())
            assert mat.shape == shp and np.array_equal(mat.row, r) and np.array_equal(mat.col, c) \
                and np.array_equal(mat.data, v)


@pytest.mark.parametrize("conn", [6, 26])
@pytest.mark.parametrize("sigma,fg", [(1.0, 0.1), (2.0, 0.03), (3.0, 0.2)])
def test_smooth_vs_oracle(syn, conn, sigma, fg):
    pred = orc.synth_pred((64, 72, 96), seed=int(sigma * 10), sigma=sigma, fg=fg)
    check_cc_task(syn, pred, 0.5, dust=50, conn=conn, offset=(1000, 2000, 3000))


@pytest.mark.parametrize("k", [1, 2, 5, 9])
@pytest.mark.parametrize("shape", [(6, 40, 44), (4, 33, 70), (3, 10, 200)])
def test_dilated_vs_oracle(syn, k, shape):
    pred = orc.synth_pred(shape, seed=k, sigma=1.2, fg=0.06)
    check_cc_task(syn, pred, 0.5, dust=5, dil=k)
    rng = np.random.default_rng(k)
    check_cc_task(syn, rng.random(shape, dtype=np.float32), 0.9, dust=0, dil=k)


def test_f_order_and_dtypes(syn):
    pred = orc.synth_pred((24, 20, 28), seed=4, sigma=1.5, fg=0.1)
    a = syn.connected_components(pred, 0.5)
    b = syn.connected_components(np.asfortranarray(pred), 0.5)
    assert np.array_equal(a, b) and b.flags.c_contiguous
    u8 = (pred * 255).astype(np.uint8)
    assert np.array_equal(syn.connected_components(u8, 127.5), orc.connected_components(u8, 127.5))
    assert syn.connected_components(pred, 0.5, dtype=np.uint64).dtype == np.uint64
    with pytest.raises(TypeError):
        syn.connected_components(pred.astype(np.float64), 0.5)
    with pytest.raises(ValueError):
        syn.connected_components(pred[0], 0.5)


def test_empty_and_full(syn):
    z = np.zeros((8, 9, 10), dtype=np.float32)
    ccs, conts, info = syn.proc.tasks.cc_task(z, 0.5, 10, verbose=False)
    assert ccs.max() == 0 and len(info) == 0 and list(info.columns) == orc.SEG_INFO_COLUMNS
    assert all(len(v) == 0 for v in conts.values()) and len(conts) == 6
    check_cc_task(syn, z + 1, 0.5, 10)
    e = syn.proc.tasks.cc_task(np.zeros((0, 4, 4), np.float32), 0.5, 1, verbose=False)
    assert e[0].shape == (0, 4, 4) and len(e[2]) == 0
    nan = np.full((4, 4, 8), np.nan, dtype=np.float32)
    assert syn.connected_components(nan, 0.0).max() == 0


def test_known_answers(syn):
    def cc(m, conn=6):
        return syn.connected_components(m.astype(np.float32), 0.5, connectivity=conn)
    m = np.zeros((3, 3, 3), bool)
    m[0, 0, 0] = m[1, 1, 1] = True                       # corner contact
    assert cc(m, 6).max() == 2 and cc(m, 18).max() == 2 and cc(m, 26).max() == 1
    m[:] = False
    m[0, 0, 0] = m[0, 1, 1] = True                       # edge contact
    assert cc(m, 6).max() == 2 and cc(m, 18).max() == 1 and cc(m, 26).max() == 1
    m[:] = False
    m[0, 0, 0] = m[0, 0, 1] = True                       # face contact
    assert cc(m, 6).max() == 1
    # U shape: the two arms get separate provisional labels and must merge late
    u = np.zeros((1, 5, 40), bool)
    u[0, :, 2] = u[0, :, 37] = True
    u[0, 4, 2:38] = True
    out = cc(u)
    assert out.max() == 1 and out.sum() == u.sum()
    # first voxel of component 1 is not its min-x voxel; numbering is by raster-first voxel
    v = np.zeros((4, 4, 4), bool)
    v[0, 3, 3] = True
    v[1, 0, 0] = True
    v[0:3, 2, 1] = True
    o = cc(v)
    assert o[0, 2, 1] == 1 and o[0, 3, 3] == 2 and o[1, 0, 0] == 3
    # snake through the volume: deep union chains
    s = np.zeros((6, 6, 70), bool)
    for x in range(6):
        for y in range(6):
            if (x * 6 + y) % 2 == 0:
                s[x, y, :] = True
            else:
                s[x, y, 0 if ((x * 6 + y) // 2) % 2 else 69] = True
    want = orc.label_c(s, 6)[0]
    assert np.array_equal(cc(s), want)


def test_dust_and_face_exemption(syn):
    m = np.zeros((10, 10, 40), np.float32)
    m[5, 5, 10:14] = 1      # interior, size 4
    m[0, 2, 3:5] = 1        # touches x-lo face, size 2
    m[4, 4, 39] = 1         # touches z-hi face, size 1
    m[7, 7, 20:30] = 1      # interior size 10
    ccs, conts, info = syn.proc.tasks.cc_task(m, 0.5, 5, verbose=False)
    assert sorted(info.index.tolist()) == [1, 2, 4]      # ids keep their raw numbers (gaps stay)
    assert ccs[5, 5, 10] == 0 and ccs[0, 2, 3] == 1 and ccs[4, 4, 39] == 2 and ccs[7, 7, 25] == 4
    assert info.loc[4, "size"] == 10 and tuple(info.loc[4, ["bbox_bz", "bbox_ez"]]) == (20, 30)
    check_cc_task(syn, m, 0.5, 5)
    check_cc_task(syn, m, 0.5, 0)
    check_cc_task(syn, m, 0.5, 11)


# ----------------------------------------------------------------------------- overlaps
@pytest.mark.parametrize("shape", [(5, 6, 7), (16, 16, 33), (24, 20, 64), (10, 12, 130)])
def test_overlaps_vs_oracle(syn, shape):
    rng = np.random.default_rng(shape[2])
    seg1 = rng.integers(0, 6, size=shape).astype(np.uint32) * rng.integers(0, 2, size=shape).astype(np.uint32)
    seg2 = orc.synth_base(shape, seed=3, block=4, bg=0.2)
    seg2[0, 0, 0] = np.uint64(2 ** 62 + 12345)
    r, c, v, sh = orc.count_overlaps_c(seg1, seg2)
    ov = syn.proc.tasks.overlap_task(seg1, seg2)
    assert ov.shape == sh
    assert np.array_equal(ov.row, r) and np.array_equal(ov.col, c) and np.array_equal(ov.data, v)


def test_overlaps_empty(syn):
    z = np.zeros((4, 5, 6), dtype=np.uint32)
    b = np.ones((4, 5, 6), dtype=np.uint64)
    assert syn.proc.tasks.overlap_task(z, b).shape == (0, 0)
    assert syn.proc.tasks.overlap_task(z + 1, b * 0).shape == (0, 0)
    one = syn.proc.tasks.overlap_task(z + 3, b * 7)
    assert one.shape == (4, 8) and one.nnz == 1 and one.data[0] == 120


@pytest.mark.parametrize("shape", [(40, 36, 64), (17, 19, 50)])
def test_fused_cc_overlap(syn, shape):
    pred = orc.synth_pred(shape, seed=9, sigma=1.5, fg=0.12)
    base = orc.synth_base(shape, seed=2, block=8, bg=0.1)
    ccs, conts, info, mat = syn.proc.tasks.cc_overlap_task(pred, base, 0.5, 20, offset=(5, 5, 5))
    want = orc.cc_task_c(pred, 0.5, 20, offset=(5, 5, 5))
    assert np.array_equal(ccs, want[0])
    r, c, v, sh = orc.count_overlaps_c(want[0], base)
    assert mat.shape == sh
    assert np.array_equal(mat.row, r) and np.array_equal(mat.col, c) and np.array_equal(mat.data, v)


# ----------------------------------------------------------------------------- seg_utils + capacity growth
def test_seg_utils_describe(syn):
    pred = orc.synth_pred((30, 28, 40), seed=6, sigma=1.5, fg=0.15)
    seg = orc.connected_components(pred, 0.5)
    seg[seg == 3] = 900       # non-dense ids
    szs = syn.seg_utils.segment_sizes(seg)
    assert szs == orc.segment_sizes_np(seg)
    assert syn.seg_utils.centers_of_mass(seg, offset=(3, 4, 5)) == orc.centers_of_mass_np(seg, offset=(3, 4, 5))
    bb = syn.seg_utils.bounding_boxes(seg, offset=(3, 4, 5))
    assert {k: v.astuple() for (k, v) in bb.items()} == orc.bounding_boxes_np(seg, offset=(3, 4, 5))
    assert np.array_equal(syn.seg_utils.nonzero_unique_ids(seg), np.unique(seg)[1:])
    f, rem = syn.seg_utils.filter_segs_by_size(seg, 30, to_ignore=[1])
    fo, remo = orc.filter_segs_by_size_np(seg, 30, to_ignore=[1])
    assert np.array_equal(f, fo) and rem == remo
    same, _ = syn.seg_utils.filter_segs_by_size(seg, 0)
    assert same is seg
    r = syn.seg_utils.relabel_data(seg, {1: 77, 900: 5, 123456: 9}, copy=True)
    assert np.array_equal(r, orc.relabel_data_np(seg, {1: 77, 900: 5, 123456: 9}))
    inplace = seg.copy()
    out = syn.seg_utils.relabel_data(inplace, {1: 0}, copy=False)
    assert out is inplace and (inplace == 1).sum() == 0


def test_capacity_retry(syn):
    # dense noise: far more runs / segments / pairs than the default capacities
    rng = np.random.default_rng(0)
    pred = rng.random((64, 64, 256), dtype=np.float32)
    base = rng.integers(1, 2 ** 40, size=pred.shape, dtype=np.uint64)
    want = orc.cc_task_c(pred, 0.5, 0, with_continuations=False)
    ccs, _, info, mat = syn.proc.tasks.cc_overlap_task(pred, base, 0.5, 0)
    assert np.array_equal(ccs, want[0])
    assert np.array_equal(table_of(info)[1], table_of(want[2])[1])
    assert mat.nnz == int((ccs != 0).sum())


def test_determinism(syn):
    pred = orc.synth_pred((96, 96, 128), seed=1, sigma=1.0, fg=0.1)
    base = orc.synth_base(pred.shape, seed=1, block=16)
    a = syn.proc.tasks.cc_overlap_task(pred, base, 0.5, 10)
    b = syn.proc.tasks.cc_overlap_task(pred, base, 0.5, 10)
    assert np.array_equal(a[0], b[0]) and a[2].equals(b[2])
    assert np.array_equal(a[3].row, b[3].row) and np.array_equal(a[3].col, b[3].col)
    assert np.array_equal(a[3].data, b[3].data)


# ----------------------------------------------------------------------------- BASELINE configs (scaled to seconds)
def test_config1_256cubed(syn):
    pred = orc.synth_pred((256, 256, 256), seed=0, sigma=2.0, fg=0.03)
    for conn in (6, 26):
        check_cc_task(syn, pred, 0.5, 100, conn=conn)


def test_config3_dense_small(syn):
    pred = orc.synth_pred((256, 256, 128), seed=0, sigma=1.0, fg=0.10)
    base = orc.synth_base(pred.shape, seed=1)
    got = check_cc_task(syn, pred, 0.5, 100)
    r, c, v, sh = orc.count_overlaps_c(got[0], base)
    mat = syn.proc.tasks.cc_overlap_task(pred, base, 0.5, 100)[3]
    assert mat.shape == sh and np.array_equal(mat.row, r) and np.array_equal(mat.col, c) and np.array_equal(mat.data, v)


def test_full_size_properties(syn):
    """512^3 (BASELINE config 2): size-independent properties + oracle on the labels."""
    import torch
    from synaptor_b200 import _device as D
    shape = (512, 512, 512)
    tile = orc.synth_pred((128, 128, 128), seed=0, sigma=2.0, fg=0.03)   # seamless (mode="wrap")
    pred = np.tile(tile, (4, 4, 4))
    base = orc.synth_base(shape, seed=1)
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.from_numpy(pred).to(dev)
    b = torch.from_numpy(base.view(np.int64)).to(dev)
    res = D.chunk_ccs_device(p, np.float32(0.5), 100, base=b)
    labels = res.labels_numpy()
    table = res.table_numpy()
    mask = pred > 0.5
    assert res.counts["fg_voxels"] == int(mask.sum())
    # kept labels are exactly the table ids, ascending; sizes add up to the labelled voxels
    assert np.all(np.diff(table[:, 0]) > 0)
    assert int(table[:, 1].sum()) == int((labels != 0).sum())
    assert not np.any(labels[~mask])
    # idempotence: relabelling the label volume as a mask reproduces the partition
    want, n = orc.label_c(mask, 6)
    assert n == res.counts["components"]
    kept = np.zeros(n + 1, dtype=np.uint32)
    kept[table[:, 0]] = table[:, 0]
    assert np.array_equal(labels, kept[want])
    # overlaps: total count = voxels with both nonzero; checksum against the oracle
    rows, cols, vals = res.pairs_numpy()
    assert int(vals.sum()) == int(((labels != 0) & (base != 0)).sum())
    r, c, v, sh = orc.count_overlaps_c(labels, base)
    assert np.array_equal(rows, r) and np.array_equal(cols.astype(np.int64), c) and np.array_equal(vals, v)
    assert (res.counts["max_label"] + 1, res.counts["max_base"] + 1) == sh


# ----------------------------------------------------------------------------- device-resident cross-chunk merge
@pytest.mark.parametrize("vshape,cshape", [((40, 36, 32), (20, 18, 16)), ((48, 24, 64), (16, 24, 32))])
def test_merge_ccs_device_vs_oracle(syn, vshape, cshape):
    import torch
    from synaptor_b200 import _device as D
    from synaptor_b200 import multi
    vol = orc.synth_pred(vshape, seed=11, sigma=2.0, fg=0.10)
    base_vol = orc.synth_base(vshape, seed=12, block=5, bg=0.1)
    grid = tuple(v // c for (v, c) in zip(vshape, cshape))
    dev = torch.device("cuda", torch.cuda.current_device())
    results, chunks = [], []
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    for gi in np.ndindex(grid):
        beg = tuple(a * b for (a, b) in zip(gi, cshape))
        sl = tuple(slice(b, b + c) for (b, c) in zip(beg, cshape))
        sub = np.ascontiguousarray(vol[sl])
        bsub = np.ascontiguousarray(base_vol[sl])
        res = D.chunk_ccs_device(torch.from_numpy(sub).to(dev), np.float32(0.5), 10, offset=beg,
                                 base=torch.from_numpy(bsub.view(np.int64)).to(dev))
        ccs, conts, info = orc.cc_task_c(sub, 0.5, 10, offset=beg)
        assert np.array_equal(res.labels_numpy(), ccs)
        results.append(res)
        chunks.append((gi, sl, ccs))
        cont_arr[gi], info_arr[gi] = conts, info
    face_shape = tuple(sorted(cshape)[1:])
    # the oracle pads faces to (two largest dims); axis order inside a face is (smaller axis index, larger)
    merged, id_maps = orc.merge_ccs_task(cont_arr, info_arr, 15, (max(cshape), max(cshape)))
    out = multi.merge_ccs_device(results, grid, rank=0, world=1, size_thr=15)
    # merge_overlaps_task on the device (proc/tasks.py:300-311): arg-max base segment per merged cleft
    ids, cols, vals = multi.merge_overlaps_device(results, out, rank=0, world=1)
    tot = {}
    for (gi, sl, ccs) in chunks:
        fin = orc.remap_ids(ccs.copy(), id_maps[gi])
        r, c, v, _ = orc.count_overlaps_c(fin, np.ascontiguousarray(base_vol[sl]))
        for (a, b, n) in zip(r.tolist(), c.tolist(), v.tolist()):
            tot[(a, b)] = tot.get((a, b), 0) + n
    best = {}
    for ((a, b), n) in sorted(tot.items()):
        if a not in best or n > best[a][1]:
            best[a] = (b, n)
    assert {int(i): (int(c), int(v)) for (i, c, v) in zip(ids, cols, vals)} == best
    got = {int(r[0]): r[1:].tolist() for r in out.merged_table.cpu().numpy()}
    assert set(got) == set(merged) and out.n_merged == len(merged)
    for (k, row) in merged.items():
        assert got[k][0] == row[0] and got[k][4:] == row[4:], (k, got[k], row)
        assert all(abs(a - b) <= 1 for (a, b) in zip(got[k][1:4], row[1:4]))   # float64 mean, truncated
    final = np.zeros(vshape, dtype=np.uint32)
    want = np.zeros(vshape, dtype=np.uint32)
    for ((gi, sl, ccs), res) in zip(chunks, results):
        final[sl] = res.labels_numpy()
        want[sl] = orc.remap_ids(ccs.copy(), id_maps[gi])
    assert np.array_equal(final, want), first_diff(final, want)
    # the merged volume is a 1:1 relabelling of whole-volume CCL restricted to the surviving segments
    whole = orc.label_c(vol > 0.5, 6)[0]
    pairs = np.unique(np.stack([final[final != 0], whole[final != 0]], 1), axis=0)
    assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


def test_cc_overlap_tasks_pipelined(syn):
    """The multi-chunk entry point gives, in order, exactly what cc_overlap_task gives per chunk."""
    rng = np.random.default_rng(5)
    chunks = []
    for i in range(5):
        shape = (24 + 8 * (i % 2), 32, 64)
        pred = orc.synth_pred(shape, seed=20 + i, sigma=1.5, fg=0.08)
        base = orc.synth_base(shape, seed=30 + i, block=8)
        chunks.append((pred, base, (i, 2 * i, 3 * i)))
    outs = list(syn.proc.tasks.cc_overlap_tasks(chunks, 0.5, 10))
    assert len(outs) == len(chunks)
    for ((pred, base, off), (ccs, conts, info, mat)) in zip(chunks, outs):
        want_ccs, _, want_info = orc.cc_task_c(pred, 0.5, 10, offset=off)
        assert np.array_equal(ccs, want_ccs), first_diff(ccs, want_ccs)
        assert np.array_equal(info.to_numpy(), want_info.to_numpy())
        r, c, v, shp = orc.count_overlaps_c(want_ccs, base)
        assert mat.shape == shp and np.array_equal(mat.row, r) and np.array_equal(mat.col, c) and np.array_equal(mat.data, v)


def test_unique_ids_many_distinct(syn):
    """More distinct uint64 ids than the first table guess (the hash table fills up and the call must grow it
    sensibly: the overflow count it reports is not the number of distinct ids)."""
    rng = np.random.default_rng(8)
    seg = rng.integers(1, 2 ** 40, size=(30, 40, 200), dtype=np.uint64)
    seg[rng.random(seg.shape) < 0.1] = 0
    got = syn.seg_utils.nonzero_unique_ids(seg)
    want = np.unique(seg)
    assert np.array_equal(got, want[want != 0])


def test_chunk_pipeline_graph_replay(syn):
    """ChunkPipeline: the captured CUDA graph gives, chunk after chunk, what the eager call gives -- including
    after a chunk that overflows the capacities sized on the first one."""
    import torch
    from synaptor_b200 import _device as D
    shape = (16, 32, 512)
    pipe = D.ChunkPipeline(shape, dust_thresh=5, offset=(1, 2, 3))
    rng = np.random.default_rng(4)
    for (i, p) in enumerate((0.02, 0.05, 0.6, 0.03)):          # 0.6: far more runs / pairs than the first chunk
        pred = rng.random(shape, dtype=np.float32) + np.float32(p - 0.5)      # foreground fraction p at threshold 0.5
        base = orc.synth_base(shape, seed=40 + i, block=8)
        pipe.pred.copy_(torch.from_numpy(pred))
        pipe.base.copy_(torch.from_numpy(base.view(np.int64)))
        res = pipe.run()
        c = pipe.counts()
        res = pipe.result
        want_ccs, _, want_info = orc.cc_task_c(pred, 0.5, 5, offset=(1, 2, 3))
        assert np.array_equal(res.labels_numpy(), want_ccs), first_diff(res.labels_numpy(), want_ccs)
        assert c["kept"] == len(want_info)
        assert np.array_equal(res.table_numpy()[:, 1:], want_info.to_numpy())
        r, cc, v, _ = orc.count_overlaps_c(want_ccs, base)
        gr, gc, gv = res.pairs_numpy()
        assert np.array_equal(gr, r) and np.array_equal(gc, cc.astype(np.uint64)) and np.array_equal(gv, v)
    assert pipe.launches_per_run >= 10


def test_g13_merge_overlaps_device_golden(syn):
    """multi.merge_overlaps_device (syn_merge_overlaps) == the executed reference's merge_overlaps_task
    (tasks.py:300-311) on the remapped g4 chunks and on matrices with exact ties."""
    import torch
    from types import SimpleNamespace
    from synaptor_b200 import multi
    g = gold("g13_merge_overlaps.npz")
    dev = torch.device("cuda", torch.cuda.current_device())

    def run(tris, n_ids):
        results, luts = [], []
        for t in tris:
            r = SimpleNamespace(labels=torch.zeros(1, device=dev),
                                pair_rows=torch.from_numpy(t[:, 0].astype(np.uint32).view(np.int32).copy()).to(dev),
                                pair_cols=torch.from_numpy(t[:, 1].astype(np.int64).copy()).to(dev),
                                pair_vals=torch.from_numpy(t[:, 2].astype(np.int64).copy()).to(dev),
                                counts={"pairs": len(t)})
            results.append(r)
            luts.append(torch.arange(n_ids, dtype=torch.int32, device=dev))        # rows are final ids already
        merge = SimpleNamespace(chunk_luts=luts, fmap=torch.zeros(n_ids, dtype=torch.int32, device=dev))
        ids, cols, vals = multi.merge_overlaps_device(results, merge, rank=0, world=1)
        return {int(i): (int(c), int(v)) for (i, c, v) in zip(ids, cols, vals)}

    tris = [g[f"ov_{i}"] for i in range(8)]
    n_ids = int(max(int(t[:, 0].max()) for t in tris if len(t))) + 1
    want = {int(r): (int(c), int(v)) for (r, c, v) in zip(g["max_row"], g["max_col"], g["max_val"])}
    assert run(tris, n_ids) == want
    want_t = {int(r): (int(c), int(v)) for (r, c, v) in g["tie_out"]}
    assert run([g["tie_in_a"], g["tie_in_b"]], 12) == want_t


@pytest.mark.parametrize("name,shape,sigma,fg", [("c2", (512, 512, 512), 2.0, 0.03), ("c3", (1024, 1024, 128), 1.0, 0.10)])
def test_full_size_bench_inputs_vs_oracle(syn, name, shape, sigma, fg):
    """BASELINE configs 2 and 3 at their FULL size on the bench's own inputs (oracle.synth_pred seed 0, base seed
    1000; not a tiled pattern): labels, table, overlap triples and their order bit-exact against the C oracle.
    Config 3 is the dense case: 6.2 M runs, 577 k components, 100 k pairs."""
    import torch
    from synaptor_b200 import _device as D
    pred = orc.synth_pred(shape, seed=0, sigma=sigma, fg=fg)
    base = orc.synth_base(shape, seed=1000)
    dev = torch.device("cuda", torch.cuda.current_device())
    res = D.chunk_ccs_device(torch.from_numpy(pred).to(dev), np.float32(0.5), 100,
                             base=torch.from_numpy(base.view(np.int64)).to(dev))
    want_ccs, _, want_info = orc.cc_task_c(pred, 0.5, 100, with_continuations=False)
    labels = res.labels_numpy()
    assert np.array_equal(labels, want_ccs), first_diff(labels, want_ccs)
    table = res.table_numpy()
    assert np.array_equal(table[:, 0], np.asarray(want_info.index, dtype=np.int64))
    assert np.array_equal(table[:, 1:], want_info.to_numpy(dtype=np.int64).reshape(-1, 10))
    r, c, v, sh = orc.count_overlaps_c(want_ccs, base)
    rows, cols, vals = res.pairs_numpy()
    assert np.array_equal(rows, r) and np.array_equal(cols.astype(np.int64), c) and np.array_equal(vals, v)
    assert (res.counts["max_label"] + 1, res.counts["max_base"] + 1) == sh
    assert res.counts["flagged_tiles"] == 0
    if name == "c3":
        assert res.counts["components"] > 500000 and res.counts["runs"] > 6000000
