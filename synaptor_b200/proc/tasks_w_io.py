"""
The chunk tasks with their IO -- the callers of the hot path, mirror of synaptor/proc/tasks_w_io.py for LOCAL storage
(a processing directory on disk, volumes as .npy files through synaptor_b200.io):

  cc_task             tasks_w_io.py:22-109    read prediction chunk -> tasks.cc_task -> label chunk, six HDF5
                                              continuation files, seg-info CSV
  merge_ccs_task      tasks_w_io.py:112-143   read all continuations + seg infos -> tasks.merge_ccs_task -> merged
                                              seg info, chunk id maps
  remap_ids_task      tasks_w_io.py:558-594   chunk id map (+ duplicate map) -> tasks.remap_ids_task -> output volume
  overlap_task        tasks_w_io.py:503-534   remapped chunk + base-segmentation chunk -> tasks.overlap_task -> CSV
  merge_overlaps_task tasks_w_io.py:537-555   all overlap CSVs -> tasks.merge_overlaps_task -> max_overlaps.df

Same function names, argument names and order, same stage lines on stdout.  The database branches of the reference
(`io.is_db_url(storagestr)`) raise NotImplementedError: database and cloud back ends are out of scope (DESIGN.md
section 7), and so are the wrappers of tasks that are not on this path (edges, anchors, duplicates).
"""
import time

from .. import io
from .. import types
from . import io as taskio
from . import tasks
from .tasks import timed


def _local_only(storagestr):
    if io.is_db_url(storagestr):
        raise NotImplementedError("database storage is outside this package's scope; use a local processing directory")


def cc_task(desc_cvname, seg_cvname, storagestr,
            cc_thresh, sz_thresh, chunk_begin, chunk_end,
            mip=0, parallel=1, storagedir=None, hashmax=100,
            timing_tag=None):
    _local_only(storagestr)
    start_time = time.time()
    storagedir = storagestr if storagedir is None else storagedir
    chunk_bounds = types.BBox3d(chunk_begin, chunk_end)

    # a chunk that already has a unique-id file was completed before; re-running would assign ids twice
    try:
        unique_ids = timed(f"Testing task completion for chunk: {chunk_bounds}",
                           taskio.read_chunk_unique_ids, storagestr, chunk_bounds)
    except Exception:
        unique_ids = []
    if len(unique_ids) > 0:
        print("Task already completed.")
        timed("(Re)writing unique ids to cache just in case",
              taskio.write_chunk_unique_ids, unique_ids, storagedir, chunk_bounds)
        return

    desc_vol = timed(f"Reading network output chunk: {chunk_bounds}",
                     io.read_cloud_volume_chunk, desc_cvname, chunk_bounds, mip=mip, parallel=parallel)

    ccs, continuations, seg_info = tasks.cc_task(desc_vol, cc_thresh, sz_thresh, offset=chunk_begin)

    timed(f"Writing seg chunk: {chunk_bounds}",
          io.write_cloud_volume_chunk, ccs, seg_cvname, chunk_bounds, mip=mip, parallel=parallel)
    timed("Writing chunk_continuations",
          taskio.write_chunk_continuations, continuations, storagedir, chunk_bounds)
    timed("Writing seg info to storage",
          taskio.write_chunk_seg_info, seg_info, storagestr, chunk_bounds)

    if timing_tag is not None:
        timed("Writing total task time",
              taskio.write_task_timing, time.time() - start_time, "ccs", timing_tag, storagestr)


def merge_ccs_task(storagestr, size_thr, max_face_shape, timing_tag=None):
    _local_only(storagestr)
    start_time = time.time()

    cont_info_arr, _ = timed("Reading continuations", taskio.read_all_continuations, storagestr)
    cleft_info_arr, local_dir = timed("Reading cleft infos", taskio.read_all_chunk_seg_infos, storagestr)
    chunk_bounds = io.extract_sorted_bboxes(local_dir)

    cons_cleft_info, chunk_id_maps = tasks.merge_ccs_task(cont_info_arr, cleft_info_arr, size_thr, max_face_shape)

    timed("Writing merged cleft info", taskio.write_merged_seg_info, cons_cleft_info, storagestr)
    timed("Writing chunk id maps", taskio.write_chunk_id_maps, chunk_id_maps, chunk_bounds, storagestr)

    if timing_tag is not None:
        timed("Writing total task time",
              taskio.write_task_timing, time.time() - start_time, "merge_ccs", timing_tag, storagestr)


def overlap_task(seg_cvname, base_seg_cvname,
                 chunk_begin, chunk_end,
                 storagedir, mip=0, seg_mip=None,
                 parallel=1, timing_tag=None):
    _local_only(storagedir)
    start_time = time.time()
    seg_mip = mip if seg_mip is None else seg_mip
    chunk_bounds = types.BBox3d(chunk_begin, chunk_end)

    seg_chunk = timed("Reading seg chunk",
                      io.read_cloud_volume_chunk, seg_cvname, chunk_bounds, mip=mip, parallel=parallel)
    base_seg_chunk = timed("Reading base seg chunk",
                           io.read_cloud_volume_chunk, base_seg_cvname, chunk_bounds, mip=seg_mip, parallel=parallel)

    overlap_matrix = tasks.overlap_task(seg_chunk, base_seg_chunk)

    timed("Writing overlap matrix", taskio.write_chunk_overlap_mat, overlap_matrix, chunk_bounds, storagedir)

    if timing_tag is not None:
        timed("Writing total task time",
              taskio.write_task_timing, time.time() - start_time, "chunk_overlap", timing_tag, storagedir)


def merge_overlaps_task(storagestr, timing_tag=None):
    _local_only(storagestr)
    start_time = time.time()

    overlap_arr = timed("Reading overlap matrices", taskio.read_all_overlap_mats, storagestr)
    if isinstance(overlap_arr, tuple):             # (grid of matrices, directory)
        overlap_arr = overlap_arr[0]

    max_overlaps = tasks.merge_overlaps_task(overlap_arr)

    timed("Writing max overlaps", taskio.write_max_overlaps, max_overlaps, storagestr)

    if timing_tag is not None:
        timed("Writing total task time",
              taskio.write_task_timing, time.time() - start_time, "merge_overlap", timing_tag, storagestr)


def remap_ids_task(seg_in_cvname, seg_out_cvname,
                   chunk_begin, chunk_end, storagestr,
                   dup_map_storagestr=None,
                   mip=0, parallel=1, timing_tag=None):
    _local_only(storagestr)
    dup_map_storagestr = storagestr if dup_map_storagestr is None else dup_map_storagestr
    start_time = time.time()
    chunk_bounds = types.BBox3d(chunk_begin, chunk_end)

    chunk_id_map = timed("Reading chunk id map", taskio.read_chunk_id_map, storagestr, chunk_bounds)
    dup_id_map = timed("Reading duplicate id map",
                       taskio.read_filtered_dup_id_map, dup_map_storagestr, chunk_id_map.values())
    seg = timed("Reading cleft chunk",
                io.read_cloud_volume_chunk, seg_in_cvname, chunk_bounds, mip=mip, parallel=parallel)

    seg = tasks.remap_ids_task(seg, chunk_id_map, dup_id_map, copy=False)

    timed("Writing results",
          io.write_cloud_volume_chunk, seg, seg_out_cvname, chunk_bounds, mip=mip, parallel=parallel)

    if timing_tag is not None:
        timed("Writing total task time",
              taskio.write_task_timing, time.time() - start_time, "remap", timing_tag, storagestr)
