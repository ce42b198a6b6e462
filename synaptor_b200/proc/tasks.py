"""
Pure (no-IO) chunk tasks with the signatures of synaptor/proc/tasks.py:26-122,
290-333 -- the drop-in seam.  numpy in, numpy / pandas / scipy out; all array
work happens on the GPU through libsynaptor_b200.so.  No CPU fallback.
"""
import time

import numpy as np

from .. import _device as D
from .. import seg_utils
from . import colnames as cn
from . import overlap
from . import seg
from .seg import continuation as cont_mod
from .seg.merge import misc as merge_misc


def timed(fn_desc, fn, *args, **kwargs):
    """Same stdout trace as the reference's timed() (tasks.py:16-23)."""
    print(fn_desc)
    start = time.time()
    result = fn(*args, **kwargs)
    print(f"{fn_desc} complete in {time.time() - start:.3f}s")
    return result


def cc_task_device(pred, cc_thresh, sz_thresh, offset=(0, 0, 0), base=None, connectivity=6, dil_param=0, **kw):
    """Device-resident variant: cuda tensors in, ChunkResult (cuda tensors) out; see _device.chunk_ccs_device."""
    return D.chunk_ccs_device(pred, np.float32(cc_thresh), int(sz_thresh), connectivity=connectivity,
                              dil_k=int(dil_param), offset=offset, base=base, **kw)


def _materialise_cc(res, shape):
    """ChunkResult -> (ccs ndarray, {Face: [Continuation]}, seg-info DataFrame)."""
    ccs = res.labels_numpy().reshape(shape)
    faces = cont_mod.faces_from_device(res.labels)
    continuations, _ = cont_mod.continuations_from_faces(faces)
    table = res.table_numpy()
    info = seg.misc.seg_info_from_table(table[:, 0], table[:, 1:])
    return ccs, continuations, info


def cc_task(desc_vol, cc_thresh, sz_thresh, offset=(0, 0, 0), overlap_seg=None, connectivity=6, dil_param=0,
            verbose=True):
    """
    - connected components of `desc_vol > cc_thresh` (ids 1..N, raster order of first voxel)
    - continuations: segments touching a chunk face (never filtered)
    - complete segments under sz_thresh voxels removed (ids are NOT renumbered)
    - centroid / size / bounding box of the survivors in a DataFrame

    Returns (ccs uint32 ndarray, {Face: [Continuation]}, DataFrame) exactly like the
    reference's cc_task (tasks.py:26-69).  `connectivity` and `dil_param` are
    extensions (the reference hard-codes 6 and has no dilation hook here).
    """
    if overlap_seg is not None:
        raise NotImplementedError("cc_task(overlap_seg=...) (multi-label CCL, tasks.py:64-67) is not built yet")
    say = print if verbose else (lambda *a, **k: None)
    start = time.time()
    say("Running connected components")
    torch = D._torch()
    d = desc_vol if isinstance(desc_vol, torch.Tensor) else np.asarray(desc_vol)
    shape = tuple(int(s) for s in d.shape)
    if len(shape) != 3:
        raise ValueError(f"expected a 3-D volume, got shape {shape}")
    if 0 in shape:
        return (np.zeros(shape, dtype=np.uint32), {f: [] for f in cont_mod.Face.all_faces()},
                seg.misc.empty_cleft_df())
    t32 = np.float32(cc_thresh) if isinstance(d, torch.Tensor) else D.float32_threshold(d.dtype, cc_thresh)
    device = torch.device("cuda", torch.cuda.current_device())
    pred = D.to_device_f32(torch, d, device)
    res = D.chunk_ccs_device(pred, t32, int(sz_thresh), connectivity=connectivity, dil_k=int(dil_param),
                             offset=tuple(int(o) for o in offset))
    say(f"Running connected components complete in {time.time() - start:.3f}s")
    t1 = time.time()
    say("Making seg info DataFrame")
    out = _materialise_cc(res, shape)
    say(f"Making seg info DataFrame complete in {time.time() - t1:.3f}s")
    return out


def overlap_task(segs, base_segs, orig_ids=True):
    """Overlap matrix between segments and a base segmentation (tasks.py:290-297; orig_ids is always True there)."""
    return timed("Counting overlaps", overlap.count_overlaps, segs, base_segs, orig_ids=True)[0]


def cc_overlap_task(desc_vol, base_segs, cc_thresh, sz_thresh, offset=(0, 0, 0), connectivity=6, dil_param=0):
    """
    chunk_ccs and chunk_overlaps fused into ONE device pass (the labels never leave
    HBM between the two tasks).  Returns (ccs, continuations, seg_info, overlap_mat)
    == cc_task(...) + (overlap_task(ccs, base_segs),).
    """
    torch = D._torch()
    d = np.asarray(desc_vol)
    shape = d.shape
    if 0 in shape:
        return cc_task(d, cc_thresh, sz_thresh, offset, verbose=False) + (seg_utils.overlap.make_coo([], [], [], 0, 0),)
    device = torch.device("cuda", torch.cuda.current_device())
    pred = D.to_device_f32(torch, d, device)
    base = D.to_device_u64(torch, base_segs, device)
    res = D.chunk_ccs_device(pred, D.float32_threshold(d.dtype, cc_thresh), int(sz_thresh),
                             connectivity=connectivity, dil_k=int(dil_param),
                             offset=tuple(int(o) for o in offset), base=base)
    ccs, continuations, info = _materialise_cc(res, shape)
    rows, cols, vals = res.pairs_numpy()
    if res.counts["max_label"] == 0 or res.counts["max_base"] == 0:
        mat = seg_utils.overlap.make_coo([], [], [], 0, 0)
    else:
        mat = seg_utils.overlap.make_coo(rows, cols, vals, res.counts["max_label"] + 1, res.counts["max_base"] + 1)
    return ccs, continuations, info, mat


def merge_ccs_task(cont_info_arr, cleft_info_arr, size_thr, max_face_shape, enforce_overlaps=False):
    """
    Global merge of the chunk results (tasks.py:72-122): global ids, face matching,
    merged id maps, merged seg-info table, size threshold over merged segments.
    Returns (DataFrame of merged clefts, object ndarray of {chunk id: global id}).
    """
    cons_cleft_info, chunk_id_maps = timed("Assigning new cleft ids", seg.merge.assign_unique_ids_serial,
                                           cleft_info_arr)
    cont_info_arr = timed("Applying chunk_id_maps to continuations", seg.merge.apply_chunk_id_maps,
                          cont_info_arr, chunk_id_maps)
    cont_id_map = timed("Merging connected continuations", seg.merge.merge_continuations, cont_info_arr,
                        max_face_shape=max_face_shape, overlap_df=(cons_cleft_info if enforce_overlaps else None))
    chunk_id_maps = timed("Updating chunk id maps", seg.merge.update_chunk_id_maps, chunk_id_maps, cont_id_map)
    seg.merge.add_new_ids(cons_cleft_info, cont_id_map, new_id_colname=cn.dst_id)
    cons_cleft_info = timed("Merging cleft dataframes", seg.merge.merge_seginfo_df, cons_cleft_info,
                            new_id_colname=cn.dst_id)
    size_thr_map = timed("Enforcing size threshold over merged ccs", seg.merge.enforce_size_threshold,
                         cons_cleft_info, size_thr)
    chunk_id_maps = timed("Updating chunk id maps (for size thresholding)", seg.merge.update_chunk_id_maps,
                          chunk_id_maps, size_thr_map)
    return cons_cleft_info, chunk_id_maps


def merge_overlaps_task(overlaps_arr):
    """Sums the chunk overlap matrices and takes the arg-max base segment per row (tasks.py:300-311)."""
    full_overlap = timed("Merging overlap matrices", overlap.merge.consolidate_overlaps, overlaps_arr)
    return timed("Finding segments with maximal overlap", overlap.find_max_overlaps, full_overlap)


def remap_ids_task(clefts, *id_maps, copy=False):
    """Maps the ids of `clefts` through the id maps, applied in the order listed (tasks.py:314-333)."""
    id_map = id_maps[0]
    for (i, next_map) in enumerate(id_maps[1:]):
        id_map = timed("Updating id map: round {i}".format(i=i + 1), merge_misc.update_id_map, id_map, next_map,
                       reused_ids=True)
    return timed("Relabeling data by id map", seg_utils.relabel_data, clefts, id_map, copy=copy)
