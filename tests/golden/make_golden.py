"""
Generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (build container only).

The reference (/root/reference, read-only) is imported unmodified.  Its three C++
extensions are the ones oracle/Makefile compiles into oracle/_ref from the
reference's own sources.  Dependencies that are not installed here and are not
on the chunk_ccs / chunk_overlaps / merge_ccs path are replaced by MagicMock;
two that ARE on the path are absent too and get minimal stand-ins:

* cc3d.connected_components  -> scipy.ndimage.label (same partition; same 1..N
  raster-first numbering -- see oracle/oracle.py header)
* igraph.Graph               -> scipy.sparse.csgraph.connected_components
  (proc/utils.py:39-74 only needs Graph(n), .vs["ids"], len(.vs), .add_edges,
  .components(); the result is order-free because make_id_map maps to min())

Run:  python tests/golden/make_golden.py      (writes next to this file)
The fixtures travel to the GPU box; this script and /root/reference do not need to.
"""
import importlib.util
import io
import os
import sys
import types
from contextlib import redirect_stdout
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _install_reference():
    from scipy import ndimage
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components as cg_cc

    for name in ["h5py", "cloudvolume", "taskqueue", "sqlalchemy", "sqlalchemy.sql", "psycopg2", "boto3",
                 "fastremap", "zstandard", "future", "future.standard_library", "builtins_stub"]:
        sys.modules.setdefault(name, mock.MagicMock())
    import google.cloud  # namespace package exists in this image
    storage = mock.MagicMock()
    sys.modules["google.cloud.storage"] = storage
    google.cloud.storage = storage

    cc3d = types.ModuleType("cc3d")

    def connected_components(arr, connectivity=26, **kw):
        rank = {6: 1, 18: 2, 26: 3}[connectivity]
        struct = ndimage.generate_binary_structure(3, rank)
        arr = np.asarray(arr)
        if arr.dtype == bool:
            return ndimage.label(arr, structure=struct)[0].astype(np.uint32)
        # multi-label: components of equal nonzero value, numbered by first voxel
        out = np.zeros(arr.shape, dtype=np.int64)
        firsts = []
        for v in np.unique(arr):
            if v == 0:
                continue
            lab, n = ndimage.label(arr == v, structure=struct)
            for k in range(1, n + 1):
                firsts.append((np.flatnonzero(lab.ravel() == k)[0], v, k))
            out[lab > 0] = -1
        firsts.sort()
        out[:] = 0
        cache = {}
        for (new, (_, v, k)) in enumerate(firsts, start=1):
            if v not in cache:
                cache[v] = ndimage.label(arr == v, structure=struct)[0]
            out[cache[v] == k] = new
        return out.astype(np.uint32)

    cc3d.connected_components = connected_components
    sys.modules["cc3d"] = cc3d

    igraph = types.ModuleType("igraph")

    class _VS(dict):
        def __init__(self, n):
            super().__init__()
            self._n = n

        def __len__(self):
            return self._n

    class Graph:
        def __init__(self, n):
            self.n = n
            self.vs = _VS(n)
            self.edges = []

        def add_edges(self, es):
            self.edges.extend(list(es))

        def components(self):
            if self.n == 0:
                return []
            if self.edges:
                e = np.array(self.edges, dtype=np.int64)
                m = coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(self.n, self.n))
            else:
                m = coo_matrix((self.n, self.n))
            _, lab = cg_cc(m, directed=False)
            out = {}
            for (i, l) in enumerate(lab.tolist()):
                out.setdefault(l, []).append(i)
            return list(out.values())

    igraph.Graph = Graph
    sys.modules["igraph"] = igraph

    # the reference's own extensions, compiled from its sources into oracle/_ref
    refdir = os.path.join(ROOT, "oracle", "_ref")
    for name in ["_describe", "_relabel", "_seg_utils"]:
        f = [x for x in os.listdir(refdir) if x.startswith(name + ".")][0]
        full = "synaptor.seg_utils." + name
        spec = importlib.util.spec_from_file_location(full, os.path.join(refdir, f))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[full] = mod

    sys.path.insert(0, REF)
    import synaptor  # noqa
    return synaptor


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def flatten_conts(conts):
    """{Face: [Continuation]} -> arrays (face_axis, face_hi, segid, n) + concatenated coords."""
    meta, coords = [], []
    for (face, lst) in conts.items():
        for c in lst:
            meta.append((face.axis, int(face.hi_index), int(c.segid), len(c.face_coords)))
            coords.append(np.asarray(c.face_coords, dtype=np.int64).reshape(-1, 2))
    meta = np.array(meta, dtype=np.int64).reshape(-1, 4)
    coords = np.concatenate(coords) if coords else np.zeros((0, 2), dtype=np.int64)
    return meta, coords


def df_to_arrays(df):
    return np.asarray(df.index, dtype=np.float64), df.to_numpy(dtype=np.float64)


def main():
    syn = _install_reference()
    from oracle import oracle as orc
    tasks = syn.proc.tasks
    out = {}

    # ---- G1: cc_task, ragged shape, offset, dust ---------------------------------
    pred = orc.synth_pred((48, 40, 36), seed=3, sigma=1.5, fg=0.08)
    ccs, conts, info = quiet(tasks.cc_task, pred, 0.5, 20, offset=(100, 200, 300))
    meta, coords = flatten_conts(conts)
    idx, tab = df_to_arrays(info)
    np.savez_compressed(os.path.join(HERE, "g1_cc_task.npz"), pred=pred, thresh=0.5, dust=20,
                        offset=np.array([100, 200, 300]), ccs=ccs, cont_meta=meta, cont_coords=coords,
                        info_index=idx, info_table=tab, info_columns=np.array(list(info.columns)))
    out["g1"] = (int(ccs.max()), len(info))

    # ---- G2: overlap_task vs uint64 base with ids > 2**32 -------------------------
    base = orc.synth_base((48, 40, 36), seed=5, block=8, bg=0.1)
    ov = quiet(tasks.overlap_task, ccs, base)
    np.savez_compressed(os.path.join(HERE, "g2_overlap_task.npz"), ccs=ccs, base=base,
                        row=np.asarray(ov.row, dtype=np.int64), col=np.asarray(ov.col, dtype=np.int64),
                        data=np.asarray(ov.data, dtype=np.int64), shape=np.array(ov.shape, dtype=np.int64))
    mo = syn.proc.overlap.find_max_overlaps(ov)
    np.savez_compressed(os.path.join(HERE, "g2_max_overlaps.npz"), row=np.asarray(mo.row, dtype=np.int64),
                        col=np.asarray(mo.col, dtype=np.int64), data=np.asarray(mo.data, dtype=np.int64))
    out["g2"] = ov.nnz
    # empty cases
    e1 = quiet(tasks.overlap_task, np.zeros((4, 4, 4), np.uint32), base[:4, :4, :4])
    assert e1.shape == (0, 0)

    # ---- G3: dilated_components ---------------------------------------------------
    pred3 = orc.synth_pred((12, 40, 44), seed=7, sigma=1.2, fg=0.06)
    d = {"pred": pred3, "thresh": 0.5}
    for k in (1, 2, 5):
        d[f"ccs_k{k}"] = syn.proc.seg.dilated_components(pred3, k, 0.5)
        mask = pred3 > 0.5
        d[f"dil_k{k}"] = syn.seg_utils.dilate_by_k(mask, k)
    np.savez_compressed(os.path.join(HERE, "g3_dilated.npz"), **d)
    out["g3"] = [int(d[f"ccs_k{k}"].max()) for k in (1, 2, 5)]

    # ---- G4: chunked cc_task -> merge_ccs_task -> remap_ids_task -------------------
    vol = orc.synth_pred((40, 36, 32), seed=11, sigma=2.0, fg=0.10)
    from synaptor.types.bbox import chunk_bboxes
    cshape = (20, 18, 16)
    boxes = chunk_bboxes(vol.shape, cshape)
    grid = (2, 2, 2)
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    seg_arr = np.empty(grid, dtype=object)
    for (gi, bb) in zip(np.ndindex(grid), boxes):
        sl = bb.index()
        c, ct, inf = quiet(tasks.cc_task, np.ascontiguousarray(vol[sl]), 0.5, 10, offset=tuple(bb.min()))
        cont_arr[gi], info_arr[gi], seg_arr[gi] = ct, inf, c
    chunk_tables = {f"chunk_table_{i}": np.concatenate(
        [np.asarray(info_arr[gi].index, dtype=np.int64)[:, None], info_arr[gi].to_numpy(dtype=np.int64)], axis=1)
        for (i, gi) in enumerate(np.ndindex(grid))}
    chunk_segs = {f"chunk_seg_{i}": seg_arr[gi] for (i, gi) in enumerate(np.ndindex(grid))}
    merged, id_maps = quiet(tasks.merge_ccs_task, cont_arr, info_arr, 15, (20, 18))
    final = np.zeros(vol.shape, dtype=np.uint32)
    for (gi, bb) in zip(np.ndindex(grid), boxes):
        final[bb.index()] = quiet(tasks.remap_ids_task, seg_arr[gi].copy(), id_maps[gi], copy=False)
    midx, mtab = df_to_arrays(merged)
    maps = {f"id_map_{i}": np.array(sorted(id_maps[gi].items()), dtype=np.int64).reshape(-1, 2)
            for (i, gi) in enumerate(np.ndindex(grid))}
    np.savez_compressed(os.path.join(HERE, "g4_merge_ccs.npz"), vol=vol, thresh=0.5, dust=10, size_thr=15,
                        chunk_shape=np.array(cshape), max_face_shape=np.array([20, 18]),
                        merged_index=midx, merged_table=mtab, merged_columns=np.array(list(merged.columns)),
                        final=final, **maps, **chunk_tables, **chunk_segs)
    out["g4"] = (len(merged), int(final.max()))

    # ---- G5: centers_of_mass rounding / float32 path ------------------------------
    seg = np.zeros((6, 5, 700), dtype=np.uint32)
    seg[0, 0, 0:2] = 1           # mean z = 0.5 -> rounds away to 1
    seg[1:4, 1, 3] = 2           # x mean 2
    seg[0:2, 2, 7] = 3           # x mean 0.5 -> 1
    seg[5, 4, :] = 4             # sum z = 244650
    seg[2, 0:4, 10:12] = 5
    big = np.ones((300, 300, 300), dtype=np.uint32)  # sums > 2**24: float32 rounding of the sums
    c_small = syn.seg_utils.centers_of_mass(seg)
    c_big = syn.seg_utils.centers_of_mass(big)
    assert c_big[1] == (150, 150, 150), c_big
    np.savez_compressed(os.path.join(HERE, "g5_centroids.npz"), seg=seg,
                        ids=np.array(sorted(c_small), dtype=np.int64),
                        cents=np.array([c_small[i] for i in sorted(c_small)], dtype=np.int64),
                        big_shape=np.array(big.shape), big_cent=np.array(c_big[1], dtype=np.int64))

    # ---- G6: threshold semantics (float32 compare, NaN) ---------------------------
    t = np.array([0.3, np.float32(0.3), np.nextafter(np.float32(0.3), np.float32(1)), np.nan, 0.5, -0.0, 1.0],
                 dtype=np.float32).reshape(1, 1, -1)
    np.savez_compressed(os.path.join(HERE, "g6_threshold.npz"), d=t, thresh=0.3,
                        ccs=syn.proc.seg.connected_components(t, 0.3))
    # ---- G7: cc_task(overlap_seg=...): multi-label labelling + max_overlap column (tasks.py:41,64-67) ----
    pred7 = orc.synth_pred((24, 28, 40), seed=13, sigma=2.5, fg=0.25)
    base7 = orc.synth_base((24, 28, 40), seed=14, block=5, bg=0.15)
    ccs7, conts7, info7 = quiet(tasks.cc_task, pred7, 0.5, 6, offset=(5, 6, 7), overlap_seg=base7)
    meta7, coords7 = flatten_conts(conts7)
    np.savez_compressed(os.path.join(HERE, "g7_cc_task_overlap_seg.npz"), pred=pred7, base=base7, thresh=0.5, dust=6,
                        offset=np.array([5, 6, 7]), ccs=ccs7, cont_meta=meta7, cont_coords=coords7,
                        info_index=np.asarray(info7.index, dtype=np.int64),
                        info_table=info7.to_numpy().astype(np.int64), info_columns=np.array(list(info7.columns)),
                        plain_ncomp=int(syn.proc.seg.connected_components(pred7, 0.5).max()))
    out["g7"] = (int(ccs7.max()), len(info7))
    # ---- G8: on-disk formats of the chunk tasks written by the reference's own local writers -----------
    import tempfile
    import scipy.sparse as sp
    from synaptor.proc import io as taskio
    from synaptor.types.bbox import BBox3d
    sys.modules.pop("zstandard", None)      # the stand-in for the absent module would confuse pandas' to_csv
    with tempfile.TemporaryDirectory() as tmp:
        bb = BBox3d((100, 200, 300), (148, 240, 336))
        for sub in ("seg_infos", "overlaps", "id_maps"):
            os.makedirs(os.path.join(tmp, sub))
        taskio.write_chunk_seg_info(info, tmp, bb)
        taskio.write_chunk_overlap_mat(ov, bb, tmp)
        id_map = {int(k): int(k) + 1000 for k in list(info.index)[:7]}
        taskio.write_chunk_id_map(id_map, tmp, bb)
        files = {}
        for (root, _, names) in os.walk(tmp):
            for n in names:
                files[os.path.relpath(os.path.join(root, n), tmp)] = open(os.path.join(root, n)).read()
    np.savez_compressed(os.path.join(HERE, "g8_formats.npz"), names=np.array(sorted(files)),
                        texts=np.array([files[k] for k in sorted(files)]), bbox=np.array(bb.astuple()),
                        id_map_keys=np.array(list(id_map.keys())), id_map_vals=np.array(list(id_map.values())))
    out["g8"] = sorted(files)
    # ---- G9: merge sharded by hash: hashval / hash_chunk_faces, match_continuations_task per face pair,
    #          seg_graph_cc_task, merge_seginfo_task per dst-hash bucket (tasks.py:125-180) -----------------
    import collections
    import collections.abc
    collections.Iterable = collections.abc.Iterable      # proc/hashing.py:33 predates its removal (python 3.10)
    import pandas as pd
    from synaptor.proc import hashing as ref_hashing
    from synaptor.proc import colnames as rcn
    from synaptor.proc.seg.continuation import Face, hash_chunk_faces
    hash_in = [(0, 0, 0, 0), (1024, 2048, 256, 2), (20, 18, 16, 1), (7,), (123456789012, 5, 3, 0)]
    hash_scalar_in = [0, 1, 17, 99999, 2 ** 40 + 3]
    hashmax9 = 5
    hv = [ref_hashing.hashval(v, 1000) for v in hash_in]
    hs = [ref_hashing.hashval(v, 7) for v in hash_scalar_in]
    fh = {}
    for (gi, bb) in zip(np.ndindex(grid), boxes):
        d9 = hash_chunk_faces(tuple(bb.min()), tuple(bb.max()), 64)
        fh[gi] = [d9[f] for f in Face.all_faces()]
    cont9 = np.empty(grid, dtype=object)
    info9 = np.empty(grid, dtype=object)
    for (gi, bb) in zip(np.ndindex(grid), boxes):
        _, ct, inf = quiet(tasks.cc_task, np.ascontiguousarray(vol[bb.index()]), 0.5, 10, offset=tuple(bb.min()))
        cont9[gi], info9[gi] = ct, inf
    full9, maps9 = syn.proc.seg.merge.assign_unique_ids_serial(info9)
    edges9 = []
    for gi in np.ndindex(grid):
        for axis in range(3):
            if gi[axis] == grid[axis] - 1:
                continue
            gj = list(gi)
            gj[axis] += 1
            gj = tuple(gj)
            edges9.extend(quiet(tasks.match_continuations_task, cont9[gi][Face(axis, True)],
                                cont9[gj][Face(axis, False)], max_face_shape=(20, 18),
                                id_map1=maps9[gi], id_map2=maps9[gj]))
    all_ids9 = list(full9.index)
    mdf9 = quiet(tasks.seg_graph_cc_task, edges9, hashmax9, all_ids9)
    mapping9 = dict(zip(mdf9[rcn.src_id], mdf9[rcn.dst_id]))
    hashes9 = dict(zip(mdf9[rcn.src_id], mdf9[rcn.dst_id_hash]))
    full9[rcn.dst_id] = full9.index.map(mapping9)
    b_idx, b_tab, b_hash, dropped = [], [], [], []
    for h in range(hashmax9):
        part = full9[[hashes9[i] == h for i in full9.index]].copy()
        if len(part) == 0:
            continue
        mg, szm = quiet(tasks.merge_seginfo_task, part, szthresh=15)
        i9, t9 = df_to_arrays(mg)
        b_idx.append(i9); b_tab.append(t9); b_hash.append(np.full(len(i9), h)); dropped.extend(szm.keys())
    np.savez_compressed(os.path.join(HERE, "g9_hashed_merge.npz"),
                        hash_in=np.array([",".join(map(str, v)) for v in hash_in]), hash_out=np.array(hv),
                        hash_scalar_in=np.array(hash_scalar_in, dtype=np.int64), hash_scalar_out=np.array(hs),
                        face_hashes=np.array([fh[gi] for gi in np.ndindex(grid)]),
                        chunk_begin=np.array([bb.min() for bb in boxes]), chunk_end=np.array([bb.max() for bb in boxes]),
                        edges=np.array(sorted(set((int(a), int(b)) for (a, b) in edges9)), dtype=np.int64).reshape(-1, 2),
                        map_src=mdf9[rcn.src_id].to_numpy(dtype=np.int64), map_dst=mdf9[rcn.dst_id].to_numpy(dtype=np.int64),
                        map_hash=mdf9[rcn.dst_id_hash].to_numpy(dtype=np.int64), hashmax=hashmax9,
                        merged_index=np.concatenate(b_idx), merged_table=np.concatenate(b_tab),
                        merged_hash=np.concatenate(b_hash), merged_columns=np.array(list(mg.columns)),
                        dropped=np.array(sorted(int(d) for d in dropped), dtype=np.int64), size_thr=15)
    out["g9"] = (len(edges9), len(mdf9), int(sum(len(x) for x in b_idx)), len(dropped))
    # ---- G10: the seg_utils functions of the path, executed standalone (describe.py / filtering.py / relabel.py /
    #           overlap.py with orig_ids False and True) on a volume with non-dense ids ----------------------------
    su = syn.seg_utils
    seg10 = quiet(syn.proc.seg.connected_components, orc.synth_pred((26, 30, 44), seed=31, sigma=1.5, fg=0.15), 0.5)
    seg10[seg10 == 3] = 900
    seg10[seg10 == 7] = 70000
    base10 = orc.synth_base(seg10.shape, seed=32, block=7, bg=0.2)
    sz10 = su.segment_sizes(seg10)
    com10 = su.centers_of_mass(seg10, offset=(3, 4, 5))
    bb10 = su.bounding_boxes(seg10, offset=(3, 4, 5))
    ids10 = sorted(sz10)
    filt10, rem10 = su.filter_segs_by_size(seg10, 30, to_ignore=[1, 900])
    filt_id10 = su.filter_segs_by_id(seg10, [2, 900, 5], copy=True)
    map10 = {1: 77, 900: 5, 123456: 9, 2: 0}
    rel10 = su.relabel_data(seg10, map10, copy=True)
    ov10, r10, c10 = su.count_overlaps(seg10, base10)
    ovo10, _, _ = su.count_overlaps(seg10, base10, orig_ids=True)
    np.savez_compressed(os.path.join(HERE, "g10_seg_utils.npz"), seg=seg10, base=base10,
                        ids=np.array(ids10, dtype=np.int64), sizes=np.array([sz10[i] for i in ids10], dtype=np.int64),
                        coms=np.array([com10[i] for i in ids10], dtype=np.int64),
                        bboxes=np.array([bb10[i].astuple() for i in ids10], dtype=np.int64),
                        uniq=np.asarray(su.nonzero_unique_ids(seg10), dtype=np.int64),
                        filt=filt10, filt_ids=np.array(sorted(rem10), dtype=np.int64),
                        filt_sizes=np.array([rem10[i] for i in sorted(rem10)], dtype=np.int64), filt_by_id=filt_id10,
                        map_keys=np.array(list(map10.keys()), dtype=np.int64),
                        map_vals=np.array(list(map10.values()), dtype=np.int64), relabelled=rel10,
                        ov_row=np.asarray(ov10.row, dtype=np.int64), ov_col=np.asarray(ov10.col, dtype=np.int64),
                        ov_val=np.asarray(ov10.data, dtype=np.int64), ov_shape=np.array(ov10.shape),
                        ov_row_ids=np.asarray(r10, dtype=np.int64), ov_col_ids=np.asarray(c10).astype(np.uint64),
                        ovo_row=np.asarray(ovo10.row, dtype=np.int64), ovo_col=np.asarray(ovo10.col).astype(np.uint64),
                        ovo_val=np.asarray(ovo10.data, dtype=np.int64), ovo_shape=np.array(ovo10.shape, dtype=np.uint64))
    out["g10"] = (len(ids10), int(ov10.nnz), len(rem10))
    # ---- G11: files of the merge sharded by hash written by the reference's local writers (idmap.py, seginfo.py) ----
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "unique_ids"))
        bb11 = boxes[3]
        taskio.write_chunk_unique_ids({int(k): int(v) for (k, v) in maps9[(0, 1, 1)].items()}, tmp, bb11)
        dup11 = {int(d): 0 for d in dropped}
        taskio.write_dup_id_map(dup11, tmp)
        taskio.write_dup_id_map(dup11, tmp, hash_index=3)
        taskio.write_merged_seg_info(mg, tmp, hash_tag=4)
        taskio.write_merged_seg_info(mg, tmp)
        files11 = {}
        for (root, _, names) in os.walk(tmp):
            for n in names:
                files11[os.path.relpath(os.path.join(root, n), tmp)] = open(os.path.join(root, n)).read()
        back11 = taskio.read_chunk_unique_ids(tmp, bb11)
    np.savez_compressed(os.path.join(HERE, "g11_formats_hashed.npz"), names=np.array(sorted(files11)),
                        texts=np.array([files11[k] for k in sorted(files11)]), bbox=np.array(bb11.astuple()),
                        unique_keys=np.array(list(maps9[(0, 1, 1)].keys()), dtype=np.int64),
                        unique_vals=np.array(list(maps9[(0, 1, 1)].values()), dtype=np.int64),
                        dup_keys=np.array(list(dup11.keys()), dtype=np.int64),
                        merged_index=np.asarray(mg.index, dtype=np.float64), merged_table=mg.to_numpy(dtype=np.float64),
                        merged_columns=np.array(list(mg.columns)), merged_index_name=str(mg.index.name),
                        readback_keys=np.array(list(back11.keys()), dtype=np.int64),
                        readback_vals=np.array(list(back11.values()), dtype=np.int64))
    out["g11"] = sorted(files11)
    # ---- G12: continuation file names, the face-hash table and the pairing of the files of a boundary -----------
    from synaptor.proc.seg.continuation import ContinFile
    df12, tn12 = taskio.prep_face_hashes(hash_chunk_faces(tuple(boxes[0].min()), tuple(boxes[0].max()), 64),
                                          boxes[0], "procdir")
    cfs12 = [ContinFile(taskio.continuation.face_filename("procdir", bb, f), bb, f)
             for bb in boxes for f in Face.all_faces()]
    groups12 = syn.proc.seg.merge.pair_continuation_files(cfs12 + cfs12[:5])
    np.savez_compressed(os.path.join(HERE, "g12_contin_files.npz"), table_name=tn12,
                        columns=np.array(list(df12.columns)), filenames=np.array(df12[rcn.contin_filename].tolist()),
                        hashes=np.array(df12[rcn.facehash].tolist(), dtype=np.int64),
                        boxes=np.array([bb.astuple() for bb in boxes]),
                        groups=np.array(sorted("|".join(sorted(c.filename for c in grp)) for grp in groups12)),
                        face_tags=np.array([taskio.continuation.fname_face_tag(f) for f in Face.all_faces()]))
    out["g12"] = (len(df12), len(groups12))
    # ---- G13: merge_overlaps_task (tasks.py:300-311) over the overlap matrices of the REMAPPED g4 chunks (what the
    #          kube workflow feeds it), ties included; merge_seginfo_df with an extra (max_overlap) column ----------
    from scipy.sparse import coo_matrix as coo_matrix13
    base13 = (orc.synth_base(vol.shape, seed=12, block=5, bg=0.1) % np.uint64(7)) * np.uint64(1 << 33)   # few ids: ties
    ovs13 = np.empty(grid, dtype=object)
    tri13 = {}
    for (i, (gi, bb)) in enumerate(zip(np.ndindex(grid), boxes)):
        ovs13[gi] = quiet(tasks.overlap_task, np.ascontiguousarray(final[bb.index()]),
                          np.ascontiguousarray(base13[bb.index()]))
        tri13[f"ov_{i}"] = np.stack([np.asarray(ovs13[gi].row, dtype=np.int64), np.asarray(ovs13[gi].col, dtype=np.int64),
                                     np.asarray(ovs13[gi].data, dtype=np.int64)], 1)
        tri13[f"ov_shape_{i}"] = np.array(ovs13[gi].shape, dtype=np.int64)
    mo13 = quiet(tasks.merge_overlaps_task, ovs13)
    cons13 = syn.proc.overlap.merge.consolidate_overlaps(ovs13)
    vals13 = np.asarray(cons13.data, dtype=np.int64)
    nties13 = 0
    for r in np.unique(cons13.row):
        v = vals13[cons13.row == r]
        nties13 += int((v == v.max()).sum() > 1)
    # merge_seginfo_df with an extra column: the dst ids of g9's full table, plus a per-fragment marker column
    df13 = full9.copy()
    df13["max_overlap"] = np.arange(len(df13), dtype=np.int64) * 3 + 1000
    merged13 = syn.proc.seg.merge.merge_seginfo_df(df13, new_id_colname=rcn.dst_id)
    # hand-made chunk matrices with exact ties after the duplicates are summed (rows 5 and 9), entry order scrambled
    tA = coo_matrix13(([3, 2, 4, 1, 7], ([5, 5, 9, 9, 2], [20, 10, 40, 30, 8])), shape=(12, 50))
    tB = coo_matrix13(([1, 3, 6, 7], ([5, 9, 2, 11], [10, 30, 8, 3])), shape=(12, 50))
    tarr = np.empty((2, 1, 1), dtype=object)
    tarr[0, 0, 0], tarr[1, 0, 0] = tA, tB
    tmo = quiet(tasks.merge_overlaps_task, tarr)
    tri13["tie_in_a"] = np.stack([tA.row, tA.col, tA.data], 1).astype(np.int64)
    tri13["tie_in_b"] = np.stack([tB.row, tB.col, tB.data], 1).astype(np.int64)
    tri13["tie_out"] = np.stack([tmo.row, tmo.col, tmo.data], 1).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "g13_merge_overlaps.npz"), base=base13, **tri13,
                        max_row=np.asarray(mo13.row, dtype=np.int64), max_col=np.asarray(mo13.col, dtype=np.int64),
                        max_val=np.asarray(mo13.data, dtype=np.int64), n_rows_with_ties=nties13,
                        cons_row=np.asarray(cons13.row, dtype=np.int64), cons_col=np.asarray(cons13.col, dtype=np.int64),
                        cons_val=vals13,
                        df_index=np.asarray(df13.index, dtype=np.int64), df_table=df13[[rcn.size] + rcn.centroid_cols + rcn.bbox_cols].to_numpy(dtype=np.int64),
                        df_dst=df13[rcn.dst_id].to_numpy(dtype=np.float64), df_extra=df13["max_overlap"].to_numpy(dtype=np.int64),
                        merged_index=np.asarray(merged13.index, dtype=np.float64),
                        merged_columns=np.array(list(merged13.columns)),
                        merged_table=merged13.to_numpy(dtype=np.float64))
    out["g13"] = (int(mo13.nnz), nties13, len(merged13), list(merged13.columns))
    print("golden written:", out)


if __name__ == "__main__":
    main()
