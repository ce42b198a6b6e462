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
    from oracle import oracle as orc
    return orc.volume_job(vol, base_vol, cshape, thresh, dust, size_thr, vol_offset=vol_offset)
