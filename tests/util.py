"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def flatten_conts(conts):
    """
    Canonical form of a continuation dict, accepting either the oracle's
    {(axis, hi): [(segid, coords)]} or the product's {Face: [Continuation]}.
    Returns (meta (m,4) int64 = axis, hi, segid, n ; coords (sum n, 2) int64).
    """
    meta, coords = [], []
    for (face, lst) in conts.items():
        if isinstance(face, tuple):
            axis, hi = face
        else:
            axis, hi = face.axis, face.hi_index
        for c in lst:
            if isinstance(c, tuple):
                segid, fc = c
            else:
                segid, fc = c.segid, c.face_coords
            meta.append((axis, int(hi), int(segid), len(fc)))
            coords.append(np.asarray(fc, dtype=np.int64).reshape(-1, 2))
    meta = np.array(meta, dtype=np.int64).reshape(-1, 4)
    coords = np.concatenate(coords) if coords else np.zeros((0, 2), dtype=np.int64)
    return meta, coords


def table_of(df):
    """(index int64, (n,10) int64 table) of a seg-info DataFrame."""
    return np.asarray(df.index, dtype=np.int64), df.to_numpy(dtype=np.int64).reshape(len(df), len(df.columns))


def sorted_triples(r, c, v):
    a = np.stack([np.asarray(r, dtype=np.int64), np.asarray(c, dtype=np.int64), np.asarray(v, dtype=np.int64)], 1)
    if len(a) == 0:
        return a
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def oracle_volume_job(vol, base_vol, cshape, thresh, dust, size_thr, vol_offset=(0, 0, 0)):
    """
    The reference workflow on a chunked volume, all on the CPU oracle: cc_task per chunk (ragged last chunks),
    merge_ccs_task, remap_ids_task, overlap_task on the remapped chunks.
    Returns dict(grid, chunks=[(grid index, slices, begin)], final (whole remapped volume), merged {id: row},
    id_maps (object array), chunk_infos, overlaps=[(rows, cols, vals, shape)]).
    """
    from oracle import oracle as orc
    grid = tuple(-(-v // c) for (v, c) in zip(vol.shape, cshape))
    cont_arr = np.empty(grid, dtype=object)
    info_arr = np.empty(grid, dtype=object)
    chunks, ccs_all = [], []
    for gi in np.ndindex(grid):
        beg = tuple(a * b for (a, b) in zip(gi, cshape))
        sl = tuple(slice(b, min(b + c, v)) for (b, c, v) in zip(beg, cshape, vol.shape))
        off = tuple(b + o for (b, o) in zip(beg, vol_offset))
        ccs, conts, info = orc.cc_task_c(np.ascontiguousarray(vol[sl]), thresh, dust, offset=off)
        cont_arr[gi], info_arr[gi] = conts, info
        chunks.append((gi, sl, beg))
        ccs_all.append(ccs)
    face = (max(cshape), max(cshape))
    merged, id_maps = orc.merge_ccs_task(cont_arr, info_arr, size_thr, face)
    final = np.zeros(vol.shape, dtype=np.uint32)
    overlaps = []
    for ((gi, sl, beg), ccs) in zip(chunks, ccs_all):
        fin = orc.remap_ids(ccs.copy(), id_maps[gi])
        final[sl] = fin
        if base_vol is not None:
            overlaps.append(orc.count_overlaps_c(fin, np.ascontiguousarray(base_vol[sl])))
    return dict(grid=grid, chunks=chunks, final=final, merged=merged, id_maps=id_maps, infos=info_arr,
                overlaps=overlaps, chunk_ccs=ccs_all)
