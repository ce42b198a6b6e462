"""
On-disk formats of the chunk tasks on a LOCAL processing directory.

Same file names, directory layout and CSV conventions as the reference's
synaptor/proc/io (seginfo.py, overlap.py, idmap.py) over its local backend
(synaptor/io/backends/local.py:47-56: `DataFrame.to_csv(index=True, header=True)`,
`pd.read_csv(index_col=0)`), so the files written here are read by the untouched
reference stages and vice versa.  The HDF5 continuation files
(synaptor/proc/io/continuation.py:53-69, 94-118) are written and read by `h5min`, a
minimal HDF5 emitter of our own (h5py / libhdf5 are not in this image).  Cloud /
database back ends are out of scope.
"""
import os
import re

import numpy as np
import pandas as pd
import scipy.sparse as sp

from .. import colnames as cn
from ...types.bbox import BBox3d
from . import filenames as fn

BBOX_REGEXP = re.compile("-?[0-9]+_-?[0-9]+_-?[0-9]+--?[0-9]+_-?[0-9]+_-?[0-9]+")
_SPLIT = re.compile("(-?[0-9]+_-?[0-9]+_-?[0-9]+)-(-?[0-9]+_-?[0-9]+_-?[0-9]+)")


# ---- chunk tags (synaptor/io/utils.py:40-47, 65-98) ---------------------------------------------
def fname_chunk_tag(chunk_bounds):
    """`x0_y0_z0-x1_y1_z1` of a BBox3d."""
    lo, hi = chunk_bounds.min(), chunk_bounds.max()
    return "{0}_{1}_{2}-{3}_{4}_{5}".format(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2])


def bbox_from_tag(tag):
    m = _SPLIT.search(tag)
    assert m is not None, "split delimiter not found in tag"
    beg = tuple(map(int, m.group(1).split("_")))
    end = tuple(map(int, m.group(2).split("_")))
    return BBox3d(beg, end)


def bbox_from_fname(path):
    m = BBOX_REGEXP.search(path)
    assert m is not None, "bbox not found in path"
    return bbox_from_tag(m.group(0))


def read_dframe(path):
    assert os.path.isfile(path)
    return pd.read_csv(path, index_col=0)


def write_dframe(dframe, path, header=True, index=True):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    dframe.to_csv(path, index=index, header=header)


# ---- segment info (synaptor/proc/io/seginfo.py:20-50, 110-130) -----------------------------------
def chunk_info_fname(proc_url, chunk_bounds):
    return os.path.join(proc_url, fn.seginfo_dirname, fn.seginfo_fmtstr.format(tag=fname_chunk_tag(chunk_bounds)))


def write_chunk_seg_info(dframe, proc_url, chunk_bounds):
    write_dframe(dframe, chunk_info_fname(proc_url, chunk_bounds))


def read_chunk_seg_info(proc_url, chunk_bounds):
    return read_dframe(chunk_info_fname(proc_url, chunk_bounds))


def read_all_chunk_seg_infos(proc_url):
    """Object array of DataFrames shaped like the chunk grid (lexicographic chunk order), and the directory."""
    d = os.path.join(proc_url, fn.seginfo_dirname)
    fnames = sorted(os.listdir(d), key=lambda f: bbox_from_fname(f).min())
    return _grid_array({bbox_from_fname(f).min(): read_dframe(os.path.join(d, f)) for f in fnames}), d


def write_merged_seg_info(dframe, proc_url, hash_tag=None):
    """One file for the whole merge, or one per dst-id hash bucket of the sharded merge (seginfo.py:183-194)."""
    name = fn.merged_seginfo_fname if hash_tag is None else fn.merged_seginfo_fmtstr.format(tag=hash_tag)
    write_dframe(dframe, os.path.join(proc_url, name))


def read_merged_seg_info(proc_url):
    return read_dframe(os.path.join(proc_url, fn.merged_seginfo_fname))


def _grid_array(lookup):
    """{chunk begin: object} -> object ndarray indexed by grid position (synaptor/io/utils.py:105-142)."""
    keys = sorted(lookup)
    xs, ys, zs = (sorted(set(k[a] for k in keys)) for a in range(3))
    arr = np.empty((len(xs), len(ys), len(zs)), dtype=object)
    for k in keys:
        arr[xs.index(k[0]), ys.index(k[1]), zs.index(k[2])] = lookup[k]
    return arr


# ---- overlap matrices (synaptor/proc/io/overlap.py:19-60) ----------------------------------------
def empty_matrix():
    return sp.coo_matrix((0, 0), dtype=np.int64)


def overlap_mat_from_dframe(df):
    rs, cs, vs = list(df[cn.rows]), list(df[cn.cols]), list(df[cn.vals])
    if len(rs) == 0 or len(cs) == 0 or len(vs) == 0:
        return empty_matrix()
    return sp.coo_matrix((vs, (rs, cs)))


def chunk_overlap_fname(proc_url, chunk_bounds):
    return os.path.join(proc_url, fn.overlaps_dirname, fn.overlaps_fmtstr.format(tag=fname_chunk_tag(chunk_bounds)))


def write_chunk_overlap_mat(overlap_mat, chunk_bounds, proc_url):
    rs, cs, vs = sp.find(overlap_mat)
    write_dframe(pd.DataFrame({cn.rows: rs, cn.cols: cs, cn.vals: vs}), chunk_overlap_fname(proc_url, chunk_bounds))


def read_chunk_overlap_mat(fname):
    return overlap_mat_from_dframe(read_dframe(fname))


def read_all_overlap_mats(proc_url):
    d = os.path.join(proc_url, fn.overlaps_dirname)
    fnames = sorted(os.listdir(d), key=lambda f: bbox_from_fname(f).min())
    return _grid_array({bbox_from_fname(f).min(): read_chunk_overlap_mat(os.path.join(d, f)) for f in fnames}), d


def write_max_overlaps(max_overlaps, proc_url):
    """Sparse matrix of maximal overlaps -> max_overlaps.df, same three columns (proc/io/overlap.py:140-156; the
    reference passes the file name where `header` belongs there -- the intended path is used here)."""
    rs, cs, vs = sp.find(max_overlaps)
    write_dframe(pd.DataFrame({cn.rows: rs, cn.cols: cs, cn.vals: vs}), os.path.join(proc_url, fn.max_overlaps_fname))


# ---- id maps (synaptor/proc/io/idmap.py:19-37, 160-205) ------------------------------------------
def cleft_map_fname(proc_url, chunk_bounds):
    return os.path.join(proc_url, fn.idmap_dirname, fn.idmap_fmtstr.format(tag=fname_chunk_tag(chunk_bounds)))


def make_dframe_from_dict(id_map):
    df = pd.DataFrame(pd.Series(id_map), columns=[cn.dst_id])
    df.index.name = cn.src_id
    return df


def write_chunk_id_map(id_map, proc_url, chunk_bounds):
    write_dframe(make_dframe_from_dict(id_map), cleft_map_fname(proc_url, chunk_bounds))


def write_chunk_id_maps(chunk_id_maps, chunk_bounds, proc_url):
    for (id_map, bounds) in zip(chunk_id_maps.flat, chunk_bounds):
        write_chunk_id_map(id_map, proc_url, bounds)


def read_chunk_id_map(proc_url, chunk_bounds):
    df = read_dframe(cleft_map_fname(proc_url, chunk_bounds))
    return dict(zip(df.index, df[cn.dst_id]))


# ---- chunk id -> unique (global) id maps and the duplicate / size-threshold map of the merge sharded by hash
#      (synaptor/proc/io/idmap.py:26-92, 208-262) ---------------------------------------------------
def unique_ids_fname(storagestr, chunk_bounds):
    return os.path.join(storagestr, fn.uniquemap_dirname, fn.uniquemap_fmtstr.format(tag=fname_chunk_tag(chunk_bounds)))


def write_chunk_unique_ids(mapping, storagestr, chunk_bounds):
    write_dframe(make_dframe_from_dict(mapping), unique_ids_fname(storagestr, chunk_bounds))


def read_unique_ids(filename):
    df = read_dframe(filename)
    return dict(zip(df.index, df[cn.dst_id]))


def read_chunk_unique_ids(proc_url, chunk_bounds):
    return read_unique_ids(unique_ids_fname(proc_url, chunk_bounds))


def write_dup_id_map(id_map, storagestr, hash_index=None):
    name = fn.dup_map_fname if hash_index is None else fn.tagged_dup_fname.format(i=hash_index)
    write_dframe(make_dframe_from_dict(id_map), os.path.join(storagestr, name))


def read_dup_id_map(proc_url):
    """{src id: dst id}; like the reference, a missing file gives an empty mapping (with a warning)."""
    try:
        df = read_dframe(os.path.join(proc_url, fn.dup_map_fname))
    except Exception as e:
        print(e)
        print("WARNING: no dup id map found, passing empty dup mapping")
        df = pd.DataFrame({cn.dst_id: []})
    return dict(zip(df.index, df[cn.dst_id]))


def read_filtered_dup_id_map(storagestr, src_ids, chunksize=100000):
    """The rows of the duplicate map whose source id is in `src_ids` (idmap.py:228-249; a missing file gives an
    empty mapping with the reference's warning)."""
    src_ids = set(src_ids)
    mapping = dict()
    try:
        for subdf in pd.read_csv(os.path.join(storagestr, fn.dup_map_fname), index_col=0, chunksize=chunksize):
            rows = subdf.loc[sorted(src_ids.intersection(subdf.index))]
            mapping.update(dict(zip(rows.index, rows[cn.dst_id])))
    except Exception as e:
        print(e)
        print("WARNING: no dup id map found, passing empty dup mapping")
        mapping = dict()
    return mapping


# ---- task durations (synaptor/proc/io/timing.py:22-69).  The reference's FILE branch cannot run (timing_fname
#      formats "{tag}" positionally, `_write_timing_file` is called without its time argument), so only the intent
#      is mirrored: one text file `task_durations/{task_name}_{tag}` holding the seconds. ------------------------
def timing_fname(proc_dir, task_name, timing_tag):
    return os.path.join(proc_dir, fn.timing_dirname, f"{task_name}_{fn.timing_fmtstr.format(tag=timing_tag)}")


def write_task_timing(time, task_name, tag, proc_url):
    path = timing_fname(proc_url, task_name, tag)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w+") as f:
        f.write(f"{time}")


def read_task_timing(proc_url, task_name, tag):
    with open(timing_fname(proc_url, task_name, tag)) as f:
        return float(f.read().strip())


# ---- continuation file names and the face-hash table (synaptor/proc/io/continuation.py:22-50, 207-216) ---------
FACE_REGEXP = re.compile("f[0-2]_(hi|lo)")
CONTIN_TABLENAME = "continuations"


def fname_face_tag(face):
    return f"f{face.axis}_{'hi' if face.hi_index else 'lo'}"


def face_from_tag(tag):
    from ..seg.continuation import Face
    return Face(int(tag[1]), tag[-2:] == "hi")


def face_from_filename(fname):
    m = FACE_REGEXP.search(fname)
    assert m is not None, f"face not found in fname: {fname}"
    return face_from_tag(m.group(0))


def face_filename(proc_url, chunk_bounds, face, local=False):
    basename = fn.contin_fmtstr.format(chunk_tag=fname_chunk_tag(chunk_bounds), face_tag=fname_face_tag(face))
    return os.path.join(proc_url, basename) if local else os.path.join(proc_url, fn.contin_dirname, basename)


def prep_face_hashes(face_hashes, chunk_bounds, proc_dir):
    """(DataFrame of continuation file name and face hash per face, table name) -- what the merge sharded by hash
    looks a face's files up by."""
    faces, hashes = zip(*list(face_hashes.items()))
    filenames = [face_filename(proc_dir, chunk_bounds, face) for face in faces]
    return pd.DataFrame({cn.contin_filename: filenames, cn.facehash: hashes}), CONTIN_TABLENAME


# ---- continuation payload (synaptor/proc/io/continuation.py:53-69, 94-118): one HDF5 file per chunk face with the
#      scalar datasets `face_axis` (int64) and `hi_face` (h5py's boolean: int8 enum FALSE / TRUE) and the group
#      `face_coords` holding one (n, 2) int64 dataset per segment id, named by the id in decimal.  Written by h5min
#      (same objects, same types, the file-format variant h5py's default settings produce), so the untouched
#      reference's `_read_face_file` / `merge_ccs` stage reads them; see h5min's docstring for how that is pinned. ----
def _write_face_file(face_continuations, fname, face=None):
    """continuation.py:94-118 -- same arguments, same error when neither a face nor a continuation is given."""
    from . import h5min
    if face is None:
        if len(face_continuations) == 0:
            raise Exception("Need a face argument or continuations to determine face info for file")
        face = face_continuations[0].face
    os.makedirs(os.path.dirname(os.path.abspath(fname)), exist_ok=True)
    coords = {}
    for c in face_continuations:
        name = f"{c.segid}"
        if name in coords:          # h5py: "Unable to create dataset (name already exists)"
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        coords[name] = np.asarray(c.face_coords)
    h5min.write_file(fname, {"face_axis": np.asarray(face.axis), "hi_face": np.asarray(face.hi_index),
                             "face_coords": coords})


def _read_face_file(fname):
    """continuation.py:53-69 -- continuations in the order h5py's `group.keys()` yields them (link-name order)."""
    from . import h5min
    from ..seg.continuation import Continuation, Face
    f = h5min.read_file(fname)
    face = Face(f["face_axis"][()], f["hi_face"][()])
    return [Continuation(int(segid), face, coords) for (segid, coords) in f["face_coords"].items()]


def _read_legacy_file(fname):
    """The older one-file-per-chunk layout, groups `{axis}/{high|low}/{segid}` (continuation.py:72-91):
    {Face: [Continuation]} with an empty list for a face the file does not hold."""
    from . import h5min
    from ..seg.continuation import Continuation, Face
    f = h5min.read_file(fname)
    continuations = dict()
    for face in Face.all_faces():
        grp = f.get(f"{face.axis}", {}).get("high" if face.hi_index else "low", {})
        continuations[face] = [Continuation(int(i), face, coords) for (i, coords) in grp.items()]
    return continuations


def read_face_continuations(proc_url, chunk_bounds, face):
    return _read_face_file(face_filename(proc_url, chunk_bounds, face))


def read_face_filenames(filenames):
    return list(map(_read_face_file, filenames))


def write_face_continuations(continuations, proc_url, chunk_bounds, face):
    _write_face_file(continuations, face_filename(proc_url, chunk_bounds, face), face)


def write_chunk_continuations(continuations, proc_url, chunk_bounds):
    """The six face files of a chunk ({Face: [Continuation]} as returned by cc_task); continuation.py:184-192."""
    from ..seg.continuation import Face
    for face in Face.all_faces():
        _write_face_file(continuations[face], face_filename(proc_url, chunk_bounds, face), face)


def read_chunk_continuations(proc_url, chunk_bounds):
    """continuation.py:147-160"""
    from ..seg.continuation import Face
    return {face: _read_face_file(face_filename(proc_url, chunk_bounds, face)) for face in Face.all_faces()}


def read_all_continuations(storagestr):
    """All continuation files of a processing directory as an object array shaped like the chunk grid, each cell
    a {Face: [Continuation]} dict, and the directory (continuation.py:219-261)."""
    contin_dir = os.path.join(storagestr, fn.contin_dirname)
    filenames = [os.path.join(contin_dir, f) for f in sorted(os.listdir(contin_dir)) if f.endswith(".h5")]
    assert len(filenames) > 0, "No filenames returned"
    chunk_to_conts = {}
    for f in filenames:
        chunk_to_conts.setdefault(bbox_from_fname(f).min(), {})[face_from_filename(f)] = _read_face_file(f)
    return _grid_array(chunk_to_conts), contin_dir


# ---- the same payload as a NumPy .npz archive (round 2's stop-gap; kept for files written by earlier versions):
#      arrays `face_axis`, `hi_face`, `segids` (m,), `counts` (m,), `face_coords` (sum counts, 2) int64 ----------
def npz_face_filename(proc_url, chunk_bounds, face, local=False):
    return face_filename(proc_url, chunk_bounds, face, local=local)[:-len(".h5")] + ".npz"


def write_face_file_npz(face_continuations, fname, face=None):
    if face is None:
        if len(face_continuations) == 0:
            raise Exception("Need a face argument or continuations to determine face info for file")
        face = face_continuations[0].face
    os.makedirs(os.path.dirname(os.path.abspath(fname)), exist_ok=True)
    coords = [np.asarray(c.face_coords, dtype=np.int64).reshape(-1, 2) for c in face_continuations]
    np.savez(fname, face_axis=np.int64(face.axis), hi_face=np.bool_(face.hi_index),
             segids=np.array([int(c.segid) for c in face_continuations], dtype=np.int64),
             counts=np.array([len(c) for c in coords], dtype=np.int64),
             face_coords=np.concatenate(coords) if coords else np.zeros((0, 2), dtype=np.int64))


def read_face_file_npz(fname):
    from ..seg.continuation import Continuation, Face
    with np.load(fname) as f:
        face = Face(int(f["face_axis"]), bool(f["hi_face"]))
        ends = np.cumsum(f["counts"])
        coords = f["face_coords"]
        return [Continuation(int(s), face, coords[e - n:e]) for (s, n, e) in zip(f["segids"], f["counts"], ends)]
