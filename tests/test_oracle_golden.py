"""
Pins the oracle (oracle/oracle.py + oracle_c.c) against fixtures produced by the
EXECUTED reference (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from util import flatten_conts, gold, table_of


@pytest.mark.parametrize("flavour", ["np", "c"])
def test_g1_cc_task(flavour):
    g = gold("g1_cc_task.npz")
    fn = orc.cc_task_np if flavour == "np" else orc.cc_task_c
    ccs, conts, info = fn(g["pred"], float(g["thresh"]), int(g["dust"]), offset=tuple(g["offset"].tolist()))
    assert ccs.dtype == np.uint32
    assert np.array_equal(ccs, g["ccs"])
    meta, coords = flatten_conts(conts)
    assert np.array_equal(meta, g["cont_meta"])
    assert np.array_equal(coords, g["cont_coords"])
    idx, tab = table_of(info)
    assert list(info.columns) == g["info_columns"].tolist()
    assert np.array_equal(idx, g["info_index"].astype(np.int64))
    assert np.array_equal(tab, g["info_table"].astype(np.int64))


@pytest.mark.parametrize("flavour", ["np", "c"])
def test_g2_overlaps(flavour):
    g = gold("g2_overlap_task.npz")
    fn = orc.count_overlaps_np if flavour == "np" else orc.count_overlaps_c
    r, c, v, shape = fn(g["ccs"], g["base"])
    # raw order == Counter insertion order == first-occurrence raster order
    assert np.array_equal(r, g["row"]) and np.array_equal(c, g["col"]) and np.array_equal(v, g["data"])
    assert tuple(shape) == tuple(g["shape"].tolist())
    m = gold("g2_max_overlaps.npz")
    mr, mc, mv = orc.find_max_overlaps(r, c, v)
    assert np.array_equal(mr, m["row"]) and np.array_equal(mc, m["col"]) and np.array_equal(mv, m["data"])


def test_overlaps_empty():
    z = np.zeros((3, 4, 5), dtype=np.uint32)
    b = np.ones((3, 4, 5), dtype=np.uint64)
    for fn in (orc.count_overlaps_np, orc.count_overlaps_c):
        assert fn(z, b)[3] == (0, 0)
        assert fn(z + 1, b * 0)[3] == (0, 0)


@pytest.mark.parametrize("backend", ["c", "scipy"])
def test_g3_dilated(backend):
    g = gold("g3_dilated.npz")
    for k in (1, 2, 5):
        mask = g["pred"] > float(g["thresh"])
        assert np.array_equal(orc.dilate_by_k(mask, k, backend="c"), g[f"dil_k{k}"])
        ccs = orc.dilated_components(g["pred"], k, float(g["thresh"]), backend=backend)
        assert np.array_equal(ccs, g[f"ccs_k{k}"])


def test_g4_merge_ccs():
    g = gold("g4_merge_ccs.npz")
    vol = g["vol"]
    cs = tuple(g["chunk_shape"].tolist())
    grid = tuple(-(-s // c) for (s, c) in zip(vol.shape, cs))
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    segs = np.empty(grid, dtype=object)
    for (i, gi) in enumerate(np.ndindex(grid)):
        beg = tuple(a * b for (a, b) in zip(gi, cs))
        sl = tuple(slice(b, b + c) for (b, c) in zip(beg, cs))
        ccs, conts, info = orc.cc_task_c(np.ascontiguousarray(vol[sl]), float(g["thresh"]), int(g["dust"]), offset=beg)
        t = g[f"chunk_table_{i}"]
        idx, tab = table_of(info)
        assert np.array_equal(idx, t[:, 0]) and np.array_equal(tab, t[:, 1:])
        assert np.array_equal(ccs, g[f"chunk_seg_{i}"])
        cont_arr[gi], info_arr[gi], segs[gi] = conts, info, ccs
    merged, id_maps = orc.merge_ccs_task(cont_arr, info_arr, int(g["size_thr"]), tuple(g["max_face_shape"].tolist()))
    gold_rows = {int(i): r for (i, r) in zip(g["merged_index"], g["merged_table"])}
    assert set(merged) == set(gold_rows)
    for (k, row) in merged.items():
        gr = gold_rows[k].astype(np.int64)
        assert row[0] == gr[0] and row[4:] == gr[4:].tolist()
        # centroid: float64 weighted mean, truncated (merge_df.py:46-61); summation order may differ by 1 ulp
        assert np.all(np.abs(np.array(row[1:4]) - gr[1:4]) <= 1)
    final = np.zeros(vol.shape, dtype=np.uint32)
    for (i, gi) in enumerate(np.ndindex(grid)):
        assert sorted(id_maps[gi].items()) == [tuple(x) for x in g[f"id_map_{i}"].tolist()]
        beg = tuple(a * b for (a, b) in zip(gi, cs))
        sl = tuple(slice(b, b + c) for (b, c) in zip(beg, cs))
        final[sl] = orc.remap_ids(segs[gi].copy(), id_maps[gi])
    assert np.array_equal(final, g["final"])


def test_g5_centroids():
    g = gold("g5_centroids.npz")
    for fn in (orc.centers_of_mass_c, lambda s: orc.centers_of_mass_np(s)):
        c = fn(g["seg"])
        assert sorted(c) == g["ids"].tolist()
        assert np.array_equal(np.array([c[i] for i in sorted(c)]), g["cents"])
    big = np.ones(tuple(g["big_shape"].tolist()), dtype=np.uint32)
    size, sums, _ = orc.segment_stats_c(big)
    assert np.array_equal(orc.centroid_f32(sums[1:2], size[1:2, None])[0], g["big_cent"])
    assert orc.centers_of_mass_c(big)[1] == tuple(g["big_cent"].tolist())


def test_g6_threshold():
    g = gold("g6_threshold.npz")
    ccs = orc.connected_components(g["d"], float(g["thresh"]))
    assert np.array_equal(ccs, g["ccs"])
    assert np.array_equal(orc.threshold(g["d"], float(g["thresh"])), g["ccs"] != 0)


@pytest.mark.parametrize("conn", [6, 18, 26])
def test_c_labeler_matches_scipy(conn):
    rng = np.random.default_rng(conn)
    for shape in [(1, 1, 1), (3, 1, 7), (9, 11, 13), (16, 16, 33), (5, 40, 64)]:
        for p in (0.1, 0.35, 0.6, 1.0):
            m = rng.random(shape) < p
            a, na = orc.label_c(m, conn)
            b, nb = orc.label_scipy(m, conn)
            assert na == nb and np.array_equal(a, b)


def test_c_stats_match_np_flavour():
    pred = orc.synth_pred((40, 33, 29), seed=2, sigma=1.3, fg=0.1)
    a = orc.cc_task_np(pred, 0.5, 12, offset=(5, 6, 7))
    b = orc.cc_task_c(pred, 0.5, 12, offset=(5, 6, 7))
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(flatten_conts(a[1])[0], flatten_conts(b[1])[0])
    assert np.array_equal(flatten_conts(a[1])[1], flatten_conts(b[1])[1])
    ia, ta = table_of(a[2])
    ib, tb = table_of(b[2])
    assert np.array_equal(ia, ib) and np.array_equal(ta, tb)


def test_g7_cc_task_overlap_seg():
    """cc_task(overlap_seg=...) of the executed reference (multi-label labelling + max_overlap column)."""
    g = gold("g7_cc_task_overlap_seg.npz")
    ccs, conts, info = orc.cc_task_c(g["pred"], float(g["thresh"]), int(g["dust"]), offset=tuple(g["offset"].tolist()),
                                     overlap_seg=g["base"])
    assert int(g["plain_ncomp"]) < int(g["ccs"].max())          # the base segmentation really splits components
    assert np.array_equal(ccs, g["ccs"])
    meta, coords = flatten_conts(conts)
    assert np.array_equal(meta, g["cont_meta"]) and np.array_equal(coords, g["cont_coords"])
    assert list(info.columns) == g["info_columns"].tolist()
    assert np.array_equal(np.asarray(info.index, dtype=np.int64), g["info_index"])
    assert np.array_equal(info.to_numpy().astype(np.int64), g["info_table"])


def test_g10_seg_utils():
    """describe / filtering / relabel / overlap restatements against the reference's standalone seg_utils calls."""
    g = gold("g10_seg_utils.npz")
    seg, base = g["seg"], g["base"]
    ids = g["ids"].tolist()
    assert orc.segment_sizes_np(seg) == dict(zip(ids, g["sizes"].tolist()))
    want_com = {i: tuple(c) for (i, c) in zip(ids, g["coms"].tolist())}
    assert orc.centers_of_mass_np(seg, offset=(3, 4, 5)) == want_com
    sums_c = orc.centers_of_mass_c(seg)                       # the plain-C flavour (no offset)
    assert {i: (c[0] + 3, c[1] + 4, c[2] + 5) for (i, c) in sums_c.items()} == want_com
    assert orc.bounding_boxes_np(seg, offset=(3, 4, 5)) == {i: tuple(b) for (i, b) in zip(ids, g["bboxes"].tolist())}
    filt, rem = orc.filter_segs_by_size_np(seg, 30, to_ignore=[1, 900])
    assert np.array_equal(filt, g["filt"]) and rem == dict(zip(g["filt_ids"].tolist(), g["filt_sizes"].tolist()))
    mapping = dict(zip(g["map_keys"].tolist(), g["map_vals"].tolist()))
    assert np.array_equal(orc.relabel_data_np(seg, mapping), g["relabelled"])
    for fn in (orc.count_overlaps_np, orc.count_overlaps_c):
        r, c, v, shape = fn(seg, base)
        assert np.array_equal(r, g["ovo_row"]) and np.array_equal(c.astype(np.uint64), g["ovo_col"])
        assert np.array_equal(v, g["ovo_val"]) and tuple(shape) == tuple(int(x) for x in g["ovo_shape"])
        # orig_ids=False: the same entries as indices into the sorted id arrays
        assert np.array_equal(np.searchsorted(g["ov_row_ids"], r), g["ov_row"])
        assert np.array_equal(np.searchsorted(g["ov_col_ids"], c.astype(np.uint64)), g["ov_col"])


# ---- G13: merge_overlaps_task + merge_seginfo_df with extra columns, pinned to the executed reference ----------
def _g13_chunk_mats(g):
    import scipy.sparse as sp
    mats = []
    for i in range(8):
        t, shp = g[f"ov_{i}"], tuple(int(x) for x in g[f"ov_shape_{i}"])
        mats.append(sp.coo_matrix((t[:, 2], (t[:, 0], t[:, 1])), shape=shp) if shp != (0, 0)
                    else sp.coo_matrix((0, 0), dtype=np.int64))
    arr = np.empty((2, 2, 2), dtype=object)
    for (i, gi) in enumerate(np.ndindex(2, 2, 2)):
        arr[gi] = mats[i]
    return arr


def test_g13_merge_overlaps_task_host():
    """proc.tasks.merge_overlaps_task (host objects in, as the reference: tasks.py:300-311) == executed reference,
    on the remapped g4 chunks and on hand-made matrices with exact ties (the smallest base id wins a tie)."""
    import io
    from contextlib import redirect_stdout
    import scipy.sparse as sp
    from synaptor_b200.proc import tasks
    g = gold("g13_merge_overlaps.npz")
    with redirect_stdout(io.StringIO()):
        mo = tasks.merge_overlaps_task(_g13_chunk_mats(g))
    assert np.array_equal(np.asarray(mo.row, dtype=np.int64), g["max_row"])
    assert np.array_equal(np.asarray(mo.col, dtype=np.int64), g["max_col"])
    assert np.array_equal(np.asarray(mo.data, dtype=np.int64), g["max_val"])
    ta, tb = g["tie_in_a"], g["tie_in_b"]
    arr = np.empty((2, 1, 1), dtype=object)
    arr[0, 0, 0] = sp.coo_matrix((ta[:, 2], (ta[:, 0], ta[:, 1])), shape=(12, 50))
    arr[1, 0, 0] = sp.coo_matrix((tb[:, 2], (tb[:, 0], tb[:, 1])), shape=(12, 50))
    with redirect_stdout(io.StringIO()):
        tmo = tasks.merge_overlaps_task(arr)
    got = np.stack([tmo.row, tmo.col, tmo.data], 1).astype(np.int64)
    assert np.array_equal(got, g["tie_out"])
    # the oracle's restatement agrees too
    want = {int(r): (int(c), int(v)) for (r, c, v) in g["tie_out"]}
    rows = np.concatenate([ta[:, 0], tb[:, 0]]); cols = np.concatenate([ta[:, 1], tb[:, 1]]); vals = np.concatenate([ta[:, 2], tb[:, 2]])
    tot = {}
    for (a, b, n) in zip(rows.tolist(), cols.tolist(), vals.tolist()):
        tot[(a, b)] = tot.get((a, b), 0) + n
    best = {}
    for ((a, b), n) in sorted(tot.items()):
        if a not in best or n > best[a][1]:
            best[a] = (b, n)
    assert best == want


def test_g13_merge_seginfo_df_extra_columns():
    """merge_seginfo_df keeps columns beyond the ten seg-info ones, from the smallest fragment of each merged id
    (merge_df.py:30-31), and sums like pandas: identical to the executed reference's frame, compared by id."""
    import pandas as pd
    from synaptor_b200.proc import colnames as cn
    from synaptor_b200.proc.seg.merge import merge_seginfo_df
    g = gold("g13_merge_overlaps.npz")
    df = pd.DataFrame(g["df_table"], index=g["df_index"], columns=cn.seg_info_cols)
    df.index.name = cn.seg_id
    df[cn.dst_id] = g["df_dst"]
    df["max_overlap"] = g["df_extra"]
    got = merge_seginfo_df(df, new_id_colname=cn.dst_id)
    assert list(got.columns) == list(g["merged_columns"])
    want = {int(i): row for (i, row) in zip(g["merged_index"], g["merged_table"])}
    assert set(int(i) for i in got.index) == set(want)
    for (i, row) in zip(got.index, got.to_numpy(dtype=np.float64)):
        assert np.array_equal(row, want[int(i)]), (i, row, want[int(i)])


def test_kahan_sum_is_pandas_group_sum():
    """The oracle's (and the device's) ordered Kahan sum is what pandas' groupby().sum() computes."""
    import pandas as pd
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 40, size=4000)
    vals = rng.random(4000) * 10.0 ** rng.integers(-3, 9, size=4000)
    want = pd.DataFrame({"k": keys, "v": vals}).groupby("k")["v"].sum()
    for k in want.index:
        assert orc.kahan_sum(vals[keys == k]) == want[k]
