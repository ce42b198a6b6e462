"""
On-disk formats of the chunk tasks: the files synaptor_b200.proc.io writes must be byte-identical
to the ones the reference's own writers produced (tests/golden/g8_formats.npz, made by make_golden.py from the G1 /
G2 outputs of the executed reference), and its readers must give the tables back.  CPU only.
"""
import os

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from util import gold


def _golden_files():
    g = gold("g8_formats.npz")
    return {str(n): str(t) for (n, t) in zip(g["names"], g["texts"])}, g


def _g1_info():
    g1 = gold("g1_cc_task.npz")
    info = pd.DataFrame(g1["info_table"].astype(np.int64), index=g1["info_index"].astype(np.int64),
                        columns=[str(c) for c in g1["info_columns"]])
    info.index.name = "cleft_segid"
    return info


def test_writers_reproduce_reference_files(tmp_path):
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.types.bbox import BBox3d
    files, g = _golden_files()
    bb = BBox3d(tuple(g["bbox"][:3].tolist()), tuple(g["bbox"][3:].tolist()))
    assert taskio.fname_chunk_tag(bb) == "100_200_300-148_240_336"
    assert taskio.bbox_from_fname("x/seg_info_100_200_300-148_240_336.df") == bb
    taskio.write_chunk_seg_info(_g1_info(), str(tmp_path), bb)
    g2 = gold("g2_overlap_task.npz")
    ov = sp.coo_matrix((g2["data"], (g2["row"], g2["col"])), shape=tuple(g2["shape"].tolist()))
    taskio.write_chunk_overlap_mat(ov, bb, str(tmp_path))
    taskio.write_chunk_id_map(dict(zip(g["id_map_keys"].tolist(), g["id_map_vals"].tolist())), str(tmp_path), bb)
    for (rel, want) in files.items():
        got = open(os.path.join(tmp_path, rel)).read()
        assert got == want, f"{rel} differs from the file the reference wrote"


def test_readers_parse_reference_files(tmp_path):
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.types.bbox import BBox3d
    files, g = _golden_files()
    for (rel, text) in files.items():
        os.makedirs(os.path.dirname(os.path.join(tmp_path, rel)), exist_ok=True)
        open(os.path.join(tmp_path, rel), "w").write(text)
    bb = BBox3d(tuple(g["bbox"][:3].tolist()), tuple(g["bbox"][3:].tolist()))
    info = taskio.read_chunk_seg_info(str(tmp_path), bb)
    want = _g1_info()
    assert info.index.name == "cleft_segid" and list(info.columns) == list(want.columns)
    assert np.array_equal(info.to_numpy(), want.to_numpy()) and np.array_equal(info.index, want.index)
    ov = taskio.read_chunk_overlap_mat(taskio.chunk_overlap_fname(str(tmp_path), bb))
    g2 = gold("g2_overlap_task.npz")
    ref = sp.coo_matrix((g2["data"], (g2["row"], g2["col"])), shape=tuple(g2["shape"].tolist()))
    assert (ov.tocsr() != ref.tocsr()[:ov.shape[0], :ov.shape[1]]).nnz == 0 and ov.nnz == ref.nnz
    assert taskio.read_chunk_id_map(str(tmp_path), bb) == dict(zip(g["id_map_keys"].tolist(), g["id_map_vals"].tolist()))
    arr, _ = taskio.read_all_chunk_seg_infos(str(tmp_path))
    assert arr.shape == (1, 1, 1) and len(arr[0, 0, 0]) == len(want)
    mats, _ = taskio.read_all_overlap_mats(str(tmp_path))
    assert mats.shape == (1, 1, 1) and mats[0, 0, 0].nnz == ref.nnz


def test_g9_hashing_golden():
    """hashval / hash_chunk_faces / add_hashed_index against values produced by the reference (proc/hashing.py)."""
    from synaptor_b200.proc import hashing
    from synaptor_b200.proc.seg.continuation import Face, hash_chunk_faces
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "g9_hashed_merge.npz"))
    for (text, want) in zip(g["hash_in"].tolist(), g["hash_out"].tolist()):
        assert hashing.hashval(tuple(int(x) for x in text.split(",")), 1000) == want
    for (v, want) in zip(g["hash_scalar_in"].tolist(), g["hash_scalar_out"].tolist()):
        assert hashing.hashval(v, 7) == want
    for (b, e, want) in zip(g["chunk_begin"].tolist(), g["chunk_end"].tolist(), g["face_hashes"].tolist()):
        d = hash_chunk_faces(tuple(b), tuple(e), 64)
        assert [d[f] for f in Face.all_faces()] == want
    # faces on either side of an interior boundary share a bucket
    lo = hash_chunk_faces((0, 0, 0), (20, 18, 16), 1 << 20)
    hi = hash_chunk_faces((20, 0, 0), (40, 18, 16), 1 << 20)
    assert lo[Face(0, True)] == hi[Face(0, False)]
    import pandas as pd
    df = pd.DataFrame({"dst_id": g["map_dst"]})
    df = hashing.add_hashed_index(df, ["dst_id"], int(g["hashmax"]), indexname="dst_id_hash")
    assert df["dst_id_hash"].tolist() == g["map_hash"].tolist()
    df2 = pd.DataFrame({"a": [1.0, np.nan], "b": [2, 3]})
    assert hashing.add_hashed_index(df2, ["a"], 9, null_fillval=-1)["hashed_index"].tolist()[1] == -1


def test_g9_merge_seginfo_task_host():
    """merge_seginfo_task (host-only part of the merge sharded by hash, tasks.py:166-180) per dst-hash bucket."""
    from synaptor_b200.proc import tasks, colnames as cn
    from synaptor_b200.proc.seg import merge
    g, g4 = gold("g9_hashed_merge.npz"), gold("g4_merge_ccs.npz")
    grid = (2, 2, 2)
    info = np.empty(grid, dtype=object)
    for (i, gi) in enumerate(np.ndindex(grid)):
        t = g4[f"chunk_table_{i}"]
        df = pd.DataFrame(t[:, 1:], index=t[:, 0], columns=cn.seg_info_cols)
        df.index.name = cn.seg_id
        info[gi] = df
    full, _ = merge.assign_unique_ids_serial(info)
    dst = dict(zip(g["map_src"].tolist(), g["map_dst"].tolist()))
    bucket = dict(zip(g["map_src"].tolist(), g["map_hash"].tolist()))
    full[cn.dst_id] = full.index.map(dst)
    want = {int(i): (int(h), r.astype(np.int64)) for (i, h, r) in zip(g["merged_index"], g["merged_hash"], g["merged_table"])}
    seen, dropped = set(), []
    for h in range(int(g["hashmax"])):
        part = full[[bucket[int(i)] == h for i in full.index]].copy()
        if len(part) == 0:
            continue
        merged, szmap = tasks.merge_seginfo_task(part, szthresh=int(g["size_thr"]))
        dropped += [int(k) for k in szmap]
        for (k, row) in zip(merged.index.tolist(), merged.to_numpy(dtype=np.int64)):
            gh, gr = want[int(k)]
            assert gh == h and row[0] == gr[0] and np.array_equal(row[4:], gr[4:])
            assert np.all(np.abs(row[1:4] - gr[1:4]) <= 1)     # float64 weighted mean, truncated
            seen.add(int(k))
    assert seen == set(want) and sorted(dropped) == g["dropped"].tolist()
    # a frame that carries cleft_segid as a column (read back from a store) merges the same way
    part = full.reset_index()
    m2, none = tasks.merge_seginfo_task(part)
    assert none is None and len(m2) == len(set(dst.values()))


def test_g11_hashed_merge_files(tmp_path):
    """unique-id maps, duplicate maps and per-bucket merged tables: byte-identical to the reference's local writers."""
    from synaptor_b200.proc import io as taskio, colnames as cn
    from synaptor_b200.types.bbox import BBox3d
    g = gold("g11_formats_hashed.npz")
    want = dict(zip(g["names"].tolist(), g["texts"].tolist()))
    bb = BBox3d(tuple(g["bbox"][:3].tolist()), tuple(g["bbox"][3:].tolist()))
    root = str(tmp_path)
    uniq = dict(zip(g["unique_keys"].tolist(), g["unique_vals"].tolist()))
    taskio.write_chunk_unique_ids(uniq, root, bb)
    dup = {k: 0 for k in g["dup_keys"].tolist()}
    taskio.write_dup_id_map(dup, root)
    taskio.write_dup_id_map(dup, root, hash_index=3)
    merged = pd.DataFrame(g["merged_table"].astype(np.int64), index=g["merged_index"].astype(np.int64),
                          columns=g["merged_columns"].tolist())
    merged.index.name = str(g["merged_index_name"])
    taskio.write_merged_seg_info(merged, root, hash_tag=4)
    taskio.write_merged_seg_info(merged, root)
    got = {}
    for (d, _, names) in os.walk(root):
        for n in names:
            got[os.path.relpath(os.path.join(d, n), root)] = open(os.path.join(d, n)).read()
    assert got == want
    back = taskio.read_chunk_unique_ids(root, bb)
    assert back == dict(zip(g["readback_keys"].tolist(), g["readback_vals"].tolist())) == uniq
    assert taskio.read_unique_ids(taskio.unique_ids_fname(root, bb)) == uniq
    assert taskio.read_dup_id_map(root) == dup
    assert taskio.read_dup_id_map(os.path.join(root, "nowhere")) == {}
    assert taskio.read_merged_seg_info(root).equals(merged)


def test_g12_continuation_file_names_and_pairing():
    """face tags, continuation file names, the face-hash table and pair_continuation_files against the reference."""
    from synaptor_b200.proc import io as taskio, colnames as cn
    from synaptor_b200.proc.seg import ContinFile, Face, hash_chunk_faces
    from synaptor_b200.proc.seg.merge import pair_continuation_files
    from synaptor_b200.types.bbox import BBox3d
    g = gold("g12_contin_files.npz")
    boxes = [BBox3d(tuple(b[:3]), tuple(b[3:])) for b in g["boxes"].tolist()]
    assert [taskio.fname_face_tag(f) for f in Face.all_faces()] == g["face_tags"].tolist()
    for f in Face.all_faces():
        name = taskio.face_filename("procdir", boxes[2], f)
        assert taskio.face_from_filename(name) == f and taskio.bbox_from_fname(name) == boxes[2]
    assert taskio.face_filename("x", boxes[0], Face(1, False), local=True) == os.path.join(
        "x", os.path.basename(taskio.face_filename("x", boxes[0], Face(1, False))))
    df, table = taskio.prep_face_hashes(hash_chunk_faces(tuple(boxes[0].min()), tuple(boxes[0].max()), 64),
                                        boxes[0], "procdir")
    assert table == str(g["table_name"]) and list(df.columns) == g["columns"].tolist()
    assert df[cn.contin_filename].tolist() == g["filenames"].tolist()
    assert df[cn.facehash].tolist() == g["hashes"].tolist()
    cfs = [ContinFile(taskio.face_filename("procdir", bb, f), bb, f) for bb in boxes for f in Face.all_faces()]
    groups = pair_continuation_files(cfs + cfs[:5])
    assert sorted("|".join(sorted(c.filename for c in grp)) for grp in groups) == g["groups"].tolist()
    # the two files of an interior boundary land in one group, and in one hash bucket
    assert sum(len(grp) == 2 for grp in groups) == 12 and sum(len(grp) == 1 for grp in groups) == 24


def test_continuation_files_roundtrip(tmp_path):
    """The continuation payload (proc/io/continuation.py:94-118) under the reference's file names, as HDF5 written by
    h5min: same content back in h5py's key order (link names sorted as byte strings), empty faces included; the
    .npz stop-gap of earlier versions still reads."""
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.proc.seg.continuation import Continuation, Face
    from synaptor_b200.types.bbox import BBox3d
    bb = BBox3d((0, 18, 16), (20, 36, 32))
    conts = {f: [] for f in Face.all_faces()}
    f0 = Face(0, True)
    conts[f0] = [Continuation(7, f0, np.array([[1, 2], [1, 3]], dtype=np.int64)),
                 Continuation(12, f0, np.array([[5, 5]], dtype=np.int64)),
                 Continuation(100, f0, np.array([[0, 0]], dtype=np.int64))]
    taskio.write_chunk_continuations(conts, str(tmp_path), bb)
    name = taskio.face_filename(str(tmp_path), bb, f0)
    assert name.endswith("continuations/continuations_0_18_16-20_36_32_f0_hi.h5") and os.path.exists(name)
    back = taskio.read_chunk_continuations(str(tmp_path), bb)
    assert [c.segid for c in back[f0]] == [100, 12, 7]            # "100" < "12" < "7": what group.keys() yields
    assert back[f0][0].face.axis == 0 and bool(back[f0][0].face.hi_index) is True and back[f0][0].face == f0
    got = {c.segid: c.face_coords for c in back[f0]}
    assert np.array_equal(got[7], [[1, 2], [1, 3]]) and np.array_equal(got[12], [[5, 5]]) and got[7].dtype == np.int64
    assert all(back[f] == [] for f in Face.all_faces() if not (f.axis == 0 and f.hi_index))
    assert [c.segid for c in taskio.read_face_continuations(str(tmp_path), bb, f0)] == [100, 12, 7]
    with pytest.raises(Exception, match="Need a face argument"):
        taskio._write_face_file([], str(tmp_path / "x.h5"))
    # .npz files of earlier versions
    taskio.write_face_file_npz(conts[f0], taskio.npz_face_filename(str(tmp_path), bb, f0), f0)
    assert [c.segid for c in taskio.read_face_file_npz(taskio.npz_face_filename(str(tmp_path), bb, f0))] == [7, 12, 100]


def test_legacy_continuation_file(tmp_path):
    """continuation.py:72-91: one file per chunk with groups {axis}/{high|low}/{segid}"""
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.proc.io import h5min
    from synaptor_b200.proc.seg.continuation import Face
    p = str(tmp_path / "legacy.h5")
    h5min.write_file(p, {"0": {"high": {"4": np.array([[1, 2]], dtype=np.int64), "15": np.array([[3, 3], [3, 4]], dtype=np.int64)}},
                         "2": {"low": {"9": np.array([[0, 0]], dtype=np.int64)}, "high": {}}})
    c = taskio._read_legacy_file(p)
    assert set(c) == set(Face.all_faces())
    assert [(x.segid, x.face_coords.tolist()) for x in c[Face(0, True)]] == [(15, [[3, 3], [3, 4]]), (4, [[1, 2]])]
    assert [(x.segid, x.face_coords.tolist()) for x in c[Face(2, False)]] == [(9, [[0, 0]])]
    assert c[Face(2, True)] == [] and c[Face(1, True)] == [] and c[Face(0, False)] == []
