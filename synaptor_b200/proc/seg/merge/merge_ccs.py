"""
Matching continuations across chunk faces (reference:
synaptor/proc/seg/merge/merge_ccs.py:35-139).  The facing 2-D faces are rebuilt
from the continuations as in the reference; the pair extraction and the graph
components run on the device (syn_face_pair_edges, syn_union_min_map).
"""
import numpy as np

from .. import continuation
from ... import utils
from ... import colnames as cn
from .... import _device as D


def pair_continuation_files(contin_files):
    """
    Groups continuation files that describe the same chunk boundary (merge_ccs.py:11-32): the key of a hi face
    is the chunk's end corner + axis, that of a lo face the end corner with the face axis replaced by the chunk's
    begin -- the key seg.hash_chunk_faces hashes.  Returns a list of lists (duplicates removed); a complete
    boundary has two entries, a face on the edge of the volume one.
    """
    pairs = dict()
    for cf in contin_files:
        corner = list(cf.bbox.max())
        if not cf.face.hi_index:
            corner[cf.face.axis] = cf.bbox.min()[cf.face.axis]
        pairs.setdefault((*corner, cf.face.axis), []).append(cf)
    return [list(set(v)) for v in pairs.values()]


def reconstruct_face(continuations, shape=(1152, 1152)):
    face = np.zeros(shape, dtype=np.uint32)
    for c in continuations:
        face[tuple(np.asarray(c.face_coords).T)] = c.segid
    return face


def _edges_device(face_pairs):
    """[(face_a, face_b)] numpy uint32 2-D pairs -> (m, 2) uint32 array of distinct (a, b) with both nonzero."""
    torch = D._torch()
    device = torch.device("cuda", torch.cuda.current_device())
    if not face_pairs:
        return np.zeros((0, 2), dtype=np.uint32)
    fa = np.concatenate([a.ravel() for (a, _) in face_pairs])
    fb = np.concatenate([b.ravel() for (_, b) in face_pairs])
    da = torch.from_numpy(fa.view(np.int32)).to(device)
    db = torch.from_numpy(fb.view(np.int32)).to(device)
    cap = max(1024, fa.size)
    edges = torch.empty((cap, 2), dtype=torch.int32, device=device)
    n_edges = torch.zeros(1, dtype=torch.int64, device=device)
    D.face_pair_edges_device(da, db, edges, n_edges)
    n = int(n_edges.cpu()[0])
    e = edges[:n].cpu().numpy().view(np.uint32)
    return np.unique(e, axis=0) if n else e.reshape(0, 2)


def match_continuations(conts1, conts2, face_shape=(1152, 1152)):
    """Distinct (id here, id there) pairs where the two reconstructed faces are both nonzero."""
    e = _edges_device([(reconstruct_face(conts1, face_shape), reconstruct_face(conts2, face_shape))])
    return [(a, b) for (a, b) in e]


def find_connected_continuations(continuation_arr, max_face_shape=(1152, 1152)):
    """Edges between global ids of segments that meet across every interior chunk face."""
    sizes = continuation_arr.shape
    pairs = []
    for index in np.ndindex(sizes):
        for axis in range(3):
            if index[axis] == sizes[axis] - 1:
                continue
            other = list(index)
            other[axis] += 1
            here = continuation_arr[index][continuation.Face(axis, True)]
            there = continuation_arr[tuple(other)][continuation.Face(axis, False)]
            if len(here) == 0 or len(there) == 0:
                continue
            pairs.append((reconstruct_face(here, max_face_shape), reconstruct_face(there, max_face_shape)))
    e = _edges_device(pairs)
    return [(int(a), int(b)) for (a, b) in e]


def filter_matches_by_overlap(matches, overlap_df, segid_col=cn.seg_id, overlap_col=cn.ovl_segid):
    if overlap_df.index.name == segid_col:
        lookup = dict(zip(overlap_df.index, overlap_df[overlap_col]))
    else:
        assert segid_col in overlap_df.columns, f"column {segid_col} not found"
        lookup = dict(zip(overlap_df[segid_col], overlap_df[overlap_col]))
    return [m for m in matches if lookup[m[0]] == lookup[m[1]]]


def merge_continuations(continuation_arr, overlap_df=None, max_face_shape=(1152, 1152), overlap_col=cn.ovl_segid):
    """{global id: smallest global id of its merged component} for every id that takes part in a match."""
    matches = find_connected_continuations(continuation_arr, max_face_shape=max_face_shape)
    if overlap_df is not None:
        matches = filter_matches_by_overlap(matches, overlap_df, overlap_col=overlap_col)
    return utils.make_id_map_from_edges(matches)
