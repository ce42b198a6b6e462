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
from . import hashing
from . import utils
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


def _materialise_cc(res, shape, between=None, say=None):
    """
    ChunkResult -> (ccs ndarray, {Face: [Continuation]}, seg-info DataFrame).  The 4 B/voxel D2H of the
    label volume runs on a copy stream while the host groups the face voxels into continuations and
    builds the table; `between()` (optional) is called after the copy was queued (more device work that
    can overlap it) and its result is appended.  `say`: prints the reference's stage lines (tasks.py:41-60).
    """
    torch = D._torch()
    say = say or (lambda *a, **k: None)
    # the small copies first: the D2H engine serves requests in order, so behind the 0.5 GB label copy the faces
    # would arrive ~10 ms later and the host work below would end up after the copies instead of under them
    t0 = time.time()
    say("Extracting continuations")
    faces = cont_mod.faces_from_device(res.labels)
    table = res.table_numpy()
    host, ev = res.labels_numpy_async(D.side_stream(torch, res.labels.device, "out"))
    continuations, _ = cont_mod.continuations_from_faces(faces)     # ~10 ms of host work, hidden under the copies
    say(f"Extracting continuations complete in {time.time() - t0:.3f}s")
    # size filter, centroids and bounding boxes were computed by the fused device pass above (one kernel chain:
    # the stages are printed for the reference's trace, their time is inside "Running connected components")
    for name in ("Filtering complete segments by size", "Computing seg centroids", "Computing seg bounding boxes"):
        say(name)
        say(f"{name} complete in 0.000s")
    t1 = time.time()
    say("Making seg info DataFrame")
    info = seg.misc.seg_info_from_table(table[:, 0], table[:, 1:])
    say(f"Making seg info DataFrame complete in {time.time() - t1:.3f}s")
    extra = between() if between is not None else None
    ev.synchronize()
    ccs = D.pinned_numpy(host, np.uint32).reshape(shape)
    if between is not None:
        return ccs, continuations, info, extra
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
    With `overlap_seg` the labelling is the multi-label one of conncomps.py:24-28 (voxels join only
    inside one non-zero base segment) and the table gets the `max_overlap` column (tasks.py:64-67).
    """
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
    ovl_dev = None
    if overlap_seg is not None:
        if tuple(int(x) for x in overlap_seg.shape) != shape:
            raise ValueError(f"overlap_seg shape {tuple(overlap_seg.shape)} != prediction shape {shape}")
        ovl_dev = D.to_device_u64(torch, overlap_seg, device)
    res = D.chunk_ccs_device(pred, t32, int(sz_thresh), connectivity=connectivity, dil_k=int(dil_param),
                             offset=tuple(int(o) for o in offset), overlap_seg=ovl_dev)
    torch.cuda.current_stream().synchronize()
    say(f"Running connected components complete in {time.time() - start:.3f}s")
    out = _materialise_cc(res, shape, say=say)
    if overlap_seg is not None:
        ccs, continuations, info = out
        say("Adding overlapping seg to DataFrame")
        t2 = time.time()
        ovl_host = overlap_seg.cpu().numpy() if isinstance(overlap_seg, torch.Tensor) else np.asarray(overlap_seg)
        info = overlap.add_overlapping_seg(info, ccs, ovl_host)
        say(f"Adding overlapping seg to DataFrame complete in {time.time() - t2:.3f}s")
        return ccs, continuations, info
    return out


def overlap_task(segs, base_segs, orig_ids=True):
    """Overlap matrix between segments and a base segmentation (tasks.py:290-297; orig_ids is always True there)."""
    return timed("Counting overlaps", overlap.count_overlaps, segs, base_segs, orig_ids=True)[0]


def cc_overlap_task(desc_vol, base_segs, cc_thresh, sz_thresh, offset=(0, 0, 0), connectivity=6, dil_param=0):
    """
    chunk_ccs followed by chunk_overlaps on the same chunk without the label volume making a round
    trip through the host.  Returns (ccs, continuations, seg_info, overlap_mat)
    == cc_task(...) + (overlap_task(ccs, base_segs),).
    CUDA-tensor base: one fused device pass.  Host arrays: copy / compute / copy pipelined over
    three streams (see below).
    """
    torch = D._torch()
    d = desc_vol if isinstance(desc_vol, torch.Tensor) else np.asarray(desc_vol)
    shape = tuple(int(x) for x in d.shape)
    if 0 in shape:
        return cc_task(d, cc_thresh, sz_thresh, offset, verbose=False) + (seg_utils.overlap.make_coo([], [], [], 0, 0),)
    device = torch.device("cuda", torch.cuda.current_device())
    t32 = np.float32(cc_thresh) if isinstance(d, torch.Tensor) else D.float32_threshold(d.dtype, cc_thresh)
    off = tuple(int(o) for o in offset)
    if isinstance(base_segs, torch.Tensor) and base_segs.is_cuda:
        # device-resident base: one fused pass (labels never leave HBM between the two tasks)
        pred = D.to_device_f32(torch, d, device)
        res = D.chunk_ccs_device(pred, t32, int(sz_thresh), connectivity=connectivity, dil_k=int(dil_param),
                                 offset=off, base=D.to_device_u64(torch, base_segs, device))
        ccs, continuations, info = _materialise_cc(res, shape)
        return ccs, continuations, info, _coo_from_result(res)

    # host arrays: the PCIe transfers dominate (1.6 GB in, 0.5 GB out for 512^3), so the work is pipelined:
    #   copy-in stream : pred H2D, then base H2D
    #   main stream    : chunk_ccs as soon as pred has arrived  ->  overlap counting once base has arrived
    #   copy-out stream: label volume D2H, concurrent with the base H2D (PCIe is full duplex)
    ctx = _start_chunk(torch, device, d, base_segs)
    return _finish_chunk(torch, ctx, shape, t32, sz_thresh, off, connectivity, dil_param)


def _start_chunk(torch, device, d, base_segs):
    """Queues the two input copies of a chunk on the copy-in stream; returns the device tensors and their events."""
    s_main = torch.cuda.current_stream()
    s_in = D.side_stream(torch, device, "in")
    s_in.wait_stream(s_main)
    with torch.cuda.stream(s_in):
        pred = D.to_device_f32(torch, d, device)
        ev_pred = torch.cuda.Event()
        ev_pred.record(s_in)
        base = D.to_device_u64(torch, base_segs, device)
        ev_base = torch.cuda.Event()
        ev_base.record(s_in)
    pred.record_stream(s_main)
    base.record_stream(s_main)
    return pred, base, ev_pred, ev_base


def _finish_chunk(torch, ctx, shape, t32, sz_thresh, off, connectivity, dil_param):
    pred, base, ev_pred, ev_base = ctx
    s_main = torch.cuda.current_stream()
    s_main.wait_event(ev_pred)
    res = D.chunk_ccs_device(pred, t32, int(sz_thresh), connectivity=connectivity, dil_k=int(dil_param), offset=off)

    def overlaps():
        s_main.wait_event(ev_base)
        return D.count_overlaps_device(res.labels, base)

    ccs, continuations, info, ov = _materialise_cc(res, shape, between=overlaps)
    return ccs, continuations, info, _coo_from_result(ov)


def cc_overlap_tasks(chunks, cc_thresh, sz_thresh, connectivity=6, dil_param=0):
    """
    cc_overlap_task over a sequence of chunks, software-pipelined: `chunks` yields (desc_vol, base_segs) or
    (desc_vol, base_segs, offset) with host arrays; results come back in order as
    (ccs, continuations, seg_info, overlap_mat), each identical to cc_overlap_task's.

    One chunk is bound by its own host-to-device copy (1.6 GB for 512^3).  Here the input copies of chunk i+1 are
    queued (asynchronously, on the copy-in stream) BEFORE chunk i is finished, so chunk i's device pipeline, its
    label volume coming back and the host-side construction of continuations / DataFrame / COO all run under the
    next chunk's copy: the PCIe link carries input copies back to back.  Single host thread; two chunks' device
    buffers alive at a time.
    """
    torch = D._torch()
    device = torch.device("cuda", torch.cuda.current_device())

    def start(item):
        off = tuple(int(o) for o in item[2]) if len(item) > 2 else (0, 0, 0)
        if isinstance(item[0], torch.Tensor) or isinstance(item[1], torch.Tensor):
            return ("plain", item, off)             # tensors (possibly resident already): the per-chunk call
        d = np.asarray(item[0])
        shape = tuple(int(x) for x in d.shape)
        if 0 in shape:
            return ("plain", item, off)
        t32 = D.float32_threshold(d.dtype, cc_thresh)
        return ("piped", _start_chunk(torch, device, d, item[1]), shape, t32, off)

    def finish(st):
        if st[0] == "plain":
            return cc_overlap_task(st[1][0], st[1][1], cc_thresh, sz_thresh, offset=st[2], connectivity=connectivity,
                                   dil_param=dil_param)
        _, ctx, shape, t32, off = st
        return _finish_chunk(torch, ctx, shape, t32, sz_thresh, off, connectivity, dil_param)

    prev = None
    reserved = False
    for item in chunks:
        if not reserved:                  # pinned blocks for the label volumes that will be alive at once
            reserved = True
            shp = tuple(int(x) for x in item[0].shape)
            if len(shp) == 3 and 0 not in shp and not isinstance(item[0], torch.Tensor):
                D.pinned_reserve(torch, shp, torch.int32, 4)
        cur = start(item)                 # queue this chunk's input copies first ...
        if prev is not None:
            yield finish(prev)            # ... then finish the previous chunk under them
        prev = cur
    if prev is not None:
        yield finish(prev)


def _coo_from_result(res):
    rows, cols, vals = res.pairs_numpy()
    if res.counts["max_label"] == 0 or res.counts["max_base"] == 0:
        return seg_utils.overlap.make_coo([], [], [], 0, 0)
    return seg_utils.overlap.make_coo(rows, cols, vals, res.counts["max_label"] + 1, res.counts["max_base"] + 1)


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


def match_continuations_task(contins1, contins2, max_face_shape=(1024, 1024), id_map1=None, id_map2=None):
    """
    One unit of the merge sharded by face hash (tasks.py:125-138): the continuations of the two
    facing faces of a chunk boundary -> list of (id here, id there) graph edges.  The id maps
    (chunk id -> global id) are applied to the continuations in place first, as in the reference.
    """
    if id_map1 is not None:
        seg.merge.apply_id_map(contins1, id_map1)
    if id_map2 is not None:
        seg.merge.apply_id_map(contins2, id_map2)
    return timed("Matching continuations", seg.merge.match_continuations, contins1, contins2,
                 face_shape=max_face_shape)


def seg_graph_cc_task(graph_edges, hashmax, all_ids):
    """
    Components of the gathered edge graph -> DataFrame (src_id, dst_id, dst_id_hash): every id of
    all_ids mapped to the smallest id of its component, and the bucket of the target id that
    shards the table merge (tasks.py:141-163).
    """
    ccs = timed("Finding connected components", utils.find_connected_components, graph_edges)
    seg_merge_map = timed("Making seg merge_map", utils.make_id_map, ccs)
    seg_merge_map = timed("Expanding mapping to include all ids", seg.merge.expand_id_map, seg_merge_map, all_ids)
    seg_merge_df = timed("Making map dataframe", merge_misc.make_map_dframe, seg_merge_map)
    return timed("Hashing dst id", hashing.add_hashed_index, seg_merge_df, [cn.dst_id], hashmax,
                 indexname=cn.dst_id_hash)


def merge_seginfo_task(seginfo_dframe, szthresh=None):
    """
    Table merge of one dst-hash bucket (tasks.py:166-180): rows carrying a dst_id column ->
    (merged table, {dropped id: 0} or None).
    """
    merged_df = timed("Merging seginfo dataframe", seg.merge.merge_seginfo_df, seginfo_dframe,
                      new_id_colname=cn.dst_id)
    szthresh_map = None
    if szthresh is not None:
        szthresh_map = timed("Enforcing size threshold over merged ccs", seg.merge.enforce_size_threshold,
                             merged_df, szthresh)
    return merged_df, szthresh_map


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


# --------------------------------------------------------------------------------------------------
# The four task waves of a chunked volume in one call (single box, one process per GPU)
# --------------------------------------------------------------------------------------------------
class VolumeResult:
    """Outputs of volume_tasks for the chunks of THIS rank (np.ndindex order) plus the global merge results."""
    __slots__ = ("ccs", "merged_info", "chunk_id_maps", "chunk_infos", "overlaps", "chunk_indices", "d2h_small_bytes")


_volume_jobs = {}


def _volume_job(vol_shape, chunk_shape, cc_thresh, sz_thresh, size_thr, offset, connectivity, rank, world, group,
                with_base):
    from .. import multi
    torch = D._torch()
    if world is None:
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
    key = (tuple(vol_shape), tuple(chunk_shape), float(cc_thresh), int(sz_thresh), int(size_thr), tuple(offset),
           int(connectivity), int(rank), int(world), with_base, torch.cuda.current_device(), id(group))
    job = _volume_jobs.get(key)
    if job is None:
        _volume_jobs.clear()                      # one resident job at a time: its buffers are volume sized
        t32 = D.float32_threshold(np.float32, cc_thresh)
        job = multi.VolumeJob(vol_shape, chunk_shape, t32, sz_thresh, size_thr, connectivity=connectivity,
                              with_base=with_base, rank=rank, world=world, group=group, vol_offset=offset, graph=False)
        _volume_jobs[key] = job
    return job


def _volume_result(job, with_base, labels, merged, tables, maps, ovs):
    out = VolumeResult()
    out.ccs = labels
    out.merged_info = seg.misc.seg_info_from_table(merged[:, 0], merged[:, 1:])
    out.chunk_id_maps = maps
    out.chunk_infos = [seg.misc.seg_info_from_table(t[:, 0], t[:, 1:]) for t in tables]
    out.overlaps = None
    small = merged.nbytes + sum(t.nbytes for t in tables)
    if with_base:
        out.overlaps = []
        for (r, c, v, shp) in ovs:
            out.overlaps.append(seg_utils.overlap.make_coo(r, c, v, shp[0], shp[1]) if shp != (0, 0)
                                else seg_utils.overlap.make_coo([], [], [], 0, 0))
            small += len(r) * 20
    out.chunk_indices = list(job.mine)
    out.d2h_small_bytes = int(small)
    return out


def volume_tasks(desc_chunks, base_chunks, vol_shape, chunk_shape, cc_thresh, sz_thresh, size_thr,
                 offset=(0, 0, 0), connectivity=6, rank=None, world=None, group=None):
    """
    chunk_ccs -> merge_ccs -> remap_ids -> chunk_overlaps of a chunked volume (tasks.py:26-69, 72-122, 314-333,
    290-297; the reference runs them as four waves of tasks with files in between).  `desc_chunks` / `base_chunks`:
    the float32 prediction and uint64 base-segmentation arrays of the chunks THIS rank owns (chunk i of the
    np.ndindex chunk grid lives on rank i % world), host arrays in any memory order (or None for base_chunks).
    Collective over `group` when world > 1 (torch.distributed, NCCL).

    Returns a VolumeResult:
      ccs            [uint32 ndarray per own chunk]  final (merged, size-thresholded) global ids, as remap_ids_task
      merged_info    DataFrame of the merged segments (merge_ccs_task's first output; int64 index, ascending id);
                     with several ranks: THIS rank's share, the merged ids that belong to its own chunks
      chunk_id_maps  [{chunk-local id: final id} per own chunk]   (merge_ccs_task's second output, own entries)
      chunk_infos    [seg-info DataFrame per own chunk]           (cc_task's third output)
      overlaps       [scipy COO per own chunk]                    overlap_task on the remapped chunk
    """
    with_base = base_chunks is not None
    job = _volume_job(vol_shape, chunk_shape, cc_thresh, sz_thresh, size_thr, offset, connectivity, rank, world, group,
                      with_base)
    return _volume_result(job, with_base, *job.run_host(desc_chunks, base_chunks))


def volume_tasks_stream(volumes, vol_shape, chunk_shape, cc_thresh, sz_thresh, size_thr, offset=(0, 0, 0),
                        connectivity=6, rank=None, world=None, group=None, with_base=True):
    """
    volume_tasks over a sequence of equally shaped volumes, software-pipelined: `volumes` yields
    (desc_chunks, base_chunks); VolumeResults come back in order, each identical to volume_tasks'.  The input copies
    of volume k + 1 are queued before volume k is collected, so its label volumes going back and the host-side
    construction of its DataFrames / id maps / COO matrices run under the next volume's input copies.
    """
    job = _volume_job(vol_shape, chunk_shape, cc_thresh, sz_thresh, size_thr, offset, connectivity, rank, world, group,
                      with_base)
    for res in job.run_host_stream(volumes):
        yield _volume_result(job, with_base, *res)
