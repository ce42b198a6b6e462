"""
Segment continuations: the contacts of segments with the six chunk faces.
Types and function signatures of synaptor/proc/seg/continuation.py:15-133.  The
face slices come from the device (syn_extract_faces); the per-id grouping is a
vectorised numpy pass over the O(surface) face voxels instead of the reference's
Python loop.
"""
from collections import namedtuple

import numpy as np

from ... import _device as D
from .. import hashing


class Continuation(object):
    """A segment touching a face of the chunk: id, face, (n, 2) in-face coordinates."""
    __slots__ = ("segid", "face", "face_coords")

    def __init__(self, segid, face, face_coords):
        self.segid = segid
        self.face = face
        self.face_coords = face_coords

    def opposite_face(self):
        return self.face.opposite()

    def __repr__(self):
        return f"<Continuation {self.segid}-{self.face}>"

    __str__ = __repr__


class Face(object):
    """One of the six faces of a chunk: (axis, hi_index)."""
    __slots__ = ("axis", "hi_index")

    def __init__(self, axis, hi_index):
        self.axis = axis
        self.hi_index = hi_index

    def opposite(self):
        return Face(self.axis, not self.hi_index)

    def extract(self, data):
        return np.take(data, -1 if self.hi_index else 0, axis=self.axis)

    @staticmethod
    def all_faces():
        return (Face(axis, hi) for axis in range(3) for hi in (True, False))

    def __eq__(self, other):
        return self.axis == other.axis and self.hi_index == other.hi_index

    def __ne__(self, other):
        return not (self == other)

    def __hash__(self):
        return hash((self.axis, self.hi_index))

    def __repr__(self):
        return f"<Face {self.axis},{'high' if self.hi_index else 'low'}>"

    __str__ = __repr__


# a continuation file with the chunk bounds and the face it holds (continuation.py:79-80)
ContinFile = namedtuple("ContinFile", ["filename", "bbox", "face"])


def continuations_of_face(face_arr, face):
    """
    [Continuation] for one 2-D face array: ids in first-occurrence raster order, each
    with its (n, 2) int64 coordinates in raster order (continuation.py:117-133).
    """
    i, j = np.nonzero(face_arr)
    if len(i) == 0:
        return []
    vals = face_arr[i, j]
    uniq, first, inv = np.unique(vals, return_index=True, return_inverse=True)
    perm = np.argsort(inv, kind="stable")
    counts = np.bincount(inv, minlength=len(uniq))
    starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
    coords = np.stack((i, j), axis=1).astype(np.int64)[perm]
    out = []
    for u in np.argsort(first, kind="stable"):
        out.append(Continuation(uniq[u], face, coords[starts[u]:starts[u] + counts[u]]))
    return out


def faces_from_device(labels_dev):
    """cuda labels -> list of six numpy uint32 2-D faces in Face.all_faces() order (one D2H copy)."""
    faces, buf = D.extract_faces_device(labels_dev)
    host = buf.cpu().numpy().view(np.uint32)
    out, o = [], 0
    for f in faces:
        n = f.numel()
        out.append(host[o:o + n].reshape(tuple(f.shape)))
        o += n
    return out


def continuations_from_faces(face_arrays):
    continuations, ids = dict(), set()
    for (face, arr) in zip(Face.all_faces(), face_arrays):
        conts = continuations_of_face(arr, face)
        continuations[face] = conts
        ids.update(c.segid for c in conts)
    return continuations, ids


def extract_all_continuations(seg):
    """
    ({Face: [Continuation]}, set of ids on any face) for a chunk segmentation
    (continuation.py:87-114).  Accepts a numpy array or a cuda tensor.
    """
    torch = D._torch()
    if isinstance(seg, torch.Tensor):
        dev = seg.view(torch.int32)
    else:
        seg = np.asarray(seg)
        if seg.size == 0:
            return {f: [] for f in Face.all_faces()}, set()
        dev = D.to_device_u32(torch, seg, torch.device("cuda", torch.cuda.current_device()))
    return continuations_from_faces(faces_from_device(dev))


def hash_chunk_faces(chunk_begin, chunk_end, maxval):
    """
    {Face: bucket below maxval} such that the two faces on either side of a chunk boundary
    share a bucket (continuation.py:136-153): the key is the chunk's end corner with, for a
    low face, the face axis replaced by the chunk's begin, plus the axis.
    """
    face_hashes = dict()
    for face in Face.all_faces():
        corner = list(chunk_end)
        if not face.hi_index:
            corner[face.axis] = chunk_begin[face.axis]
        face_hashes[face] = hashing.hashval((corner[0], corner[1], corner[2], face.axis), maxval=maxval)
    return face_hashes
