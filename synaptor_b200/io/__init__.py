"""
Volume and path helpers the task wrappers (proc/tasks_w_io.py) call -- LOCAL storage only.

The reference reads and writes volumes through CloudVolume (synaptor/io/backends/cloudvolume.py:6-36) and files
through cloud / database back ends (synaptor/io/base.py); those back ends are out of scope.
What the hot path's callers need from that package is mirrored here for a volume that lives in ONE `.npy` file on
the local disk (memory-mapped; any memory order -- CloudVolume hands out F-ordered chunks, so an F-ordered file
exercises the same seam) and a processing directory on the local disk:

  read_cloud_volume_chunk / write_cloud_volume_chunk   same names and arguments as the reference's, `cv_name` = path of
                                                       the .npy file (an optional file:// prefix is dropped)
  init_seg_volume                                      creates the zero-filled uint32 label volume (cloudvolume.py:39-60)
  is_db_url, fname_chunk_tag, bbox_from_fname, extract_sorted_bboxes   synaptor/io/base.py:307, io/utils.py:40-102
"""
import os
import re

import numpy as np


def fname_chunk_tag(chunk_bounds):
    from ..proc.io import fname_chunk_tag as f
    return f(chunk_bounds)


def bbox_from_fname(path):
    from ..proc.io import bbox_from_fname as f
    return f(path)

_REMOTE = re.compile(r"^(gs|s3|https?|precomputed|graphene|matrix|boss)://")
_DB = re.compile(r"^(postgres(ql)?|mysql|sqlite|mssql)(\+\w+)?://")


def is_db_url(uri):
    return bool(_DB.match(str(uri)))


def _local_path(cv_name):
    name = str(cv_name)
    if _REMOTE.match(name) or is_db_url(name):
        raise NotImplementedError(f"{name}: cloud / database storage is outside this package's scope; "
                                  "use a local .npy volume")
    return name[len("file://"):] if name.startswith("file://") else name


def _volume(cv_name, mode):
    path = _local_path(cv_name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"no local volume at {path} (create label volumes with init_seg_volume)")
    vol = np.load(path, mmap_mode=mode)
    if vol.ndim == 4 and vol.shape[3] == 1:       # CloudVolume's trailing channel axis
        vol = vol[:, :, :, 0]
    if vol.ndim != 3:
        raise ValueError(f"{path}: expected a 3-d volume, got shape {vol.shape}")
    return vol


def read_cloud_volume_chunk(cv_name, bbox, mip=0, parallel=1, progress=False, request_payer=None):
    """The chunk `bbox` of a local volume.  Like CloudVolume with `bounded = False, fill_missing = True`
    (cloudvolume.py:15-20) the part of the box outside the volume reads as zeros."""
    if mip != 0:
        raise NotImplementedError("local volumes have one resolution (mip 0)")
    vol = _volume(cv_name, "r")
    lo, hi = tuple(int(v) for v in bbox.min()), tuple(int(v) for v in bbox.max())
    clo = tuple(min(max(a, 0), s) for (a, s) in zip(lo, vol.shape))
    chi = tuple(min(max(b, 0), s) for (b, s) in zip(hi, vol.shape))
    inside = vol[clo[0]:chi[0], clo[1]:chi[1], clo[2]:chi[2]]
    if clo == lo and chi == hi:
        return np.array(inside, order="K")        # keeps the file's memory order (F-ordered file -> F-ordered chunk)
    out = np.zeros(tuple(b - a for (a, b) in zip(lo, hi)), dtype=vol.dtype, order="F" if np.isfortran(vol) else "C")
    out[clo[0] - lo[0]:chi[0] - lo[0], clo[1] - lo[1]:chi[1] - lo[1], clo[2] - lo[2]:chi[2] - lo[2]] = inside
    return out


def write_cloud_volume_chunk(data, cv_name, bbox, mip=0, parallel=1, non_aligned=False, progress=False):
    """`cv[bbox] = data.astype(cv.dtype)` (cloudvolume.py:23-36); the part of the box outside the volume is dropped."""
    if mip != 0:
        raise NotImplementedError("local volumes have one resolution (mip 0)")
    vol = _volume(cv_name, "r+")
    lo, hi = tuple(int(v) for v in bbox.min()), tuple(int(v) for v in bbox.max())
    if tuple(data.shape) != tuple(b - a for (a, b) in zip(lo, hi)):
        raise ValueError(f"data of shape {tuple(data.shape)} does not fill the box {lo}-{hi}")
    clo = tuple(min(max(a, 0), s) for (a, s) in zip(lo, vol.shape))
    chi = tuple(min(max(b, 0), s) for (b, s) in zip(hi, vol.shape))
    vol[clo[0]:chi[0], clo[1]:chi[1], clo[2]:chi[2]] = \
        data[clo[0] - lo[0]:chi[0] - lo[0], clo[1] - lo[1]:chi[1] - lo[1], clo[2] - lo[2]:chi[2] - lo[2]].astype(vol.dtype)
    vol.flush()


def init_seg_volume(cv_name, resolution=None, vol_size=None, description=None, owners=None, offset=(0, 0, 0),
                    sources=None, chunk_size=(64, 64, 64), dtype=np.uint32, fortran_order=True):
    """Creates the zero-filled label volume the chunk tasks write into (cloudvolume.py:39-60; resolution,
    provenance and chunking have no meaning for a .npy file and are ignored)."""
    if tuple(offset) != (0, 0, 0):
        raise NotImplementedError("local volumes start at the origin")
    path = _local_path(cv_name)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    vol = np.lib.format.open_memmap(path, mode="w+", dtype=np.dtype(dtype), shape=tuple(int(v) for v in vol_size),
                                    fortran_order=fortran_order)
    vol.flush()
    return path


def extract_sorted_bboxes(local_dir):
    """Bounding boxes of every file of a directory, sorted lexicographically (io/utils.py:94-102)."""
    return sorted((bbox_from_fname(f) for f in os.listdir(local_dir)), key=lambda bb: bb.min())
