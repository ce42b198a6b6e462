"""
The chunk tasks with their IO -- the callers of the hot path, for LOCAL storage (a processing directory on disk,
volumes as .npy files through synaptor_b200.io).  Interface of synaptor/proc/tasks_w_io.py for the five tasks of
this path: same function names, same argument names and order, same stage lines on stdout (`timed`), same files.

  cc_task             tasks_w_io.py:22-109    prediction chunk -> tasks.cc_task -> label chunk, six HDF5
                                              continuation files, seg-info CSV
  merge_ccs_task      tasks_w_io.py:112-143   all continuations + seg infos -> tasks.merge_ccs_task -> merged seg
                                              info, chunk id maps
  remap_ids_task      tasks_w_io.py:558-594   chunk id map (+ duplicate map) -> tasks.remap_ids_task -> output volume
  overlap_task        tasks_w_io.py:503-534   remapped chunk + base-segmentation chunk -> tasks.overlap_task -> CSV
  merge_overlaps_task tasks_w_io.py:537-555   all overlap CSVs -> tasks.merge_overlaps_task -> max_overlaps.df

The database branches of the reference (`io.is_db_url(storagestr)`) raise NotImplementedError here: database and
cloud back ends are out of scope, and so are the wrappers of tasks that are not on this path
(edges, anchors, duplicates, the database-backed merge sharded by hash).
"""
import time
from contextlib import contextmanager

from .. import io
from .. import types
from . import io as taskio
from . import tasks
from .tasks import timed


@contextmanager
def _task(name, timing_tag, storagestr):
    """Refuses non-local storage, and writes the task's duration under `timing_tag` when one is given (the
    reference's closing `if timing_tag is not None: ... write_task_timing` block)."""
    if io.is_db_url(storagestr):
        raise NotImplementedError("database storage is outside this package's scope; use a local processing directory")
    t0 = time.time()
    state = {"record": True}
    yield state
    if timing_tag is not None and state["record"]:
        timed("Writing total task time", taskio.write_task_timing, time.time() - t0, name, timing_tag, storagestr)


def _read_chunk(what, cvname, bounds, mip, parallel):
    return timed(what, io.read_cloud_volume_chunk, cvname, bounds, mip=mip, parallel=parallel)


def _write_chunk(what, data, cvname, bounds, mip, parallel):
    timed(what, io.write_cloud_volume_chunk, data, cvname, bounds, mip=mip, parallel=parallel)


def cc_task(desc_cvname, seg_cvname, storagestr,
            cc_thresh, sz_thresh, chunk_begin, chunk_end,
            mip=0, parallel=1, storagedir=None, hashmax=100,
            timing_tag=None):
    with _task("ccs", timing_tag, storagestr) as state:
        if storagedir is None:
            storagedir = storagestr
        bounds = types.BBox3d(chunk_begin, chunk_end)
        # a chunk that already has a unique-id file was completed before: running it again would hand out its ids twice
        try:
            done = timed(f"Testing task completion for chunk: {bounds}", taskio.read_chunk_unique_ids, storagestr, bounds)
        except Exception:
            done = []
        if len(done) > 0:
            print("Task already completed.")
            timed("(Re)writing unique ids to cache just in case", taskio.write_chunk_unique_ids, done, storagedir, bounds)
            state["record"] = False               # the reference returns here without a duration
            return
        desc_vol = _read_chunk(f"Reading network output chunk: {bounds}", desc_cvname, bounds, mip, parallel)
        ccs, continuations, seg_info = tasks.cc_task(desc_vol, cc_thresh, sz_thresh, offset=chunk_begin)
        _write_chunk(f"Writing seg chunk: {bounds}", ccs, seg_cvname, bounds, mip, parallel)
        timed("Writing chunk_continuations", taskio.write_chunk_continuations, continuations, storagedir, bounds)
        timed("Writing seg info to storage", taskio.write_chunk_seg_info, seg_info, storagestr, bounds)


def merge_ccs_task(storagestr, size_thr, max_face_shape, timing_tag=None):
    with _task("merge_ccs", timing_tag, storagestr):
        cont_grid, _ = timed("Reading continuations", taskio.read_all_continuations, storagestr)
        info_grid, info_dir = timed("Reading cleft infos", taskio.read_all_chunk_seg_infos, storagestr)
        merged, id_maps = tasks.merge_ccs_task(cont_grid, info_grid, size_thr, max_face_shape)
        timed("Writing merged cleft info", taskio.write_merged_seg_info, merged, storagestr)
        timed("Writing chunk id maps", taskio.write_chunk_id_maps, id_maps, io.extract_sorted_bboxes(info_dir), storagestr)


def overlap_task(seg_cvname, base_seg_cvname,
                 chunk_begin, chunk_end,
                 storagedir, mip=0, seg_mip=None,
                 parallel=1, timing_tag=None):
    with _task("chunk_overlap", timing_tag, storagedir):
        bounds = types.BBox3d(chunk_begin, chunk_end)
        seg_chunk = _read_chunk("Reading seg chunk", seg_cvname, bounds, mip, parallel)
        base_chunk = _read_chunk("Reading base seg chunk", base_seg_cvname, bounds, mip if seg_mip is None else seg_mip,
                                 parallel)
        matrix = tasks.overlap_task(seg_chunk, base_chunk)
        timed("Writing overlap matrix", taskio.write_chunk_overlap_mat, matrix, bounds, storagedir)


def merge_overlaps_task(storagestr, timing_tag=None):
    with _task("merge_overlap", timing_tag, storagestr):
        mats = timed("Reading overlap matrices", taskio.read_all_overlap_mats, storagestr)
        if isinstance(mats, tuple):                 # (grid of matrices, directory)
            mats = mats[0]
        timed("Writing max overlaps", taskio.write_max_overlaps, tasks.merge_overlaps_task(mats), storagestr)


def remap_ids_task(seg_in_cvname, seg_out_cvname,
                   chunk_begin, chunk_end, storagestr,
                   dup_map_storagestr=None,
                   mip=0, parallel=1, timing_tag=None):
    with _task("remap", timing_tag, storagestr):
        bounds = types.BBox3d(chunk_begin, chunk_end)
        id_map = timed("Reading chunk id map", taskio.read_chunk_id_map, storagestr, bounds)
        dup_map = timed("Reading duplicate id map", taskio.read_filtered_dup_id_map,
                        storagestr if dup_map_storagestr is None else dup_map_storagestr, id_map.values())
        seg = _read_chunk("Reading cleft chunk", seg_in_cvname, bounds, mip, parallel)
        seg = tasks.remap_ids_task(seg, id_map, dup_map, copy=False)
        _write_chunk("Writing results", seg, seg_out_cvname, bounds, mip, parallel)
