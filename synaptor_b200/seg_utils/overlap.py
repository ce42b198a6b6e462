"""
count_overlaps -- signature of synaptor/seg_utils/overlap.py:10-61.  The two
np.unique sorts, the boolean gathers and the Python Counter over every
overlapping voxel are replaced by one CUDA pass (syn_count_overlaps): a
(label, base id) pair histogram in an HBM hash table, entries ordered by first
occurrence (the Counter's insertion order).
"""
import numpy as np
from scipy import sparse

from .. import _device as D


def unique_u64(seg):
    torch = D._torch()
    dev = D.to_device_u64(torch, seg, torch.device("cuda", torch.cuda.current_device()))
    return D.unique_u64_device(dev)


def make_coo(rows, cols, vals, n_rows, n_cols):
    """COO matrix with the entries kept in the given order (overlap.py:60)."""
    if n_rows == 0 or n_cols == 0:
        return sparse.coo_matrix(([], ([], [])), shape=(0, 0))
    if n_cols - 1 > np.iinfo(np.int64).max or (len(cols) and int(cols.max()) > np.iinfo(np.int64).max):
        raise ValueError("base segment ids >= 2**63 do not fit scipy's int64 indices")
    return sparse.coo_matrix((np.asarray(vals, dtype=np.int64),
                              (np.asarray(rows, dtype=np.int64), np.asarray(cols).astype(np.int64))),
                             shape=(int(n_rows), int(n_cols)))


def count_overlaps(seg1, seg2, orig_ids=False):
    """
    Overlap matrix between two segmentations.

    Returns (scipy.sparse.coo_matrix, seg1 ids, seg2 ids): entry (r, c) = number of
    voxels where seg1 == r-th id and seg2 == c-th id (ids themselves when
    orig_ids), entries in first-occurrence raster order, shape (max1+1, max2+1)
    when orig_ids else (n1, n2); an empty side gives a (0, 0) matrix.
    """
    from .describe import nonzero_unique_ids
    torch = D._torch()
    s1 = np.asarray(seg1)
    s2 = np.asarray(seg2)
    if s1.shape != s2.shape or s1.ndim != 3:
        raise ValueError(f"segmentations must be 3-D and equally shaped, got {s1.shape} and {s2.shape}")
    if s1.dtype.kind in "iu" and s1.dtype.itemsize > 4 and s1.size and int(s1.max()) >= 2 ** 32:
        raise ValueError("seg1 ids must fit in 32 bits on the device path")
    seg1_ids = nonzero_unique_ids(s1.astype(np.uint32) if s1.dtype != np.uint32 else s1).astype(s1.dtype)
    seg2_ids = nonzero_unique_ids(s2.astype(np.uint64) if s2.dtype != np.uint64 else s2).astype(s2.dtype)
    if len(seg1_ids) == 0 or len(seg2_ids) == 0:
        return sparse.coo_matrix(([], ([], [])), shape=(0, 0)), seg1_ids, seg2_ids
    device = torch.device("cuda", torch.cuda.current_device())
    res = D.count_overlaps_device(D.to_device_u32(torch, s1, device), D.to_device_u64(torch, s2, device))
    rows, cols, vals = res.pairs_numpy()
    if orig_ids:
        mat = make_coo(rows, cols, vals, int(seg1_ids.max()) + 1, int(seg2_ids.max()) + 1)
    else:
        r = np.searchsorted(seg1_ids.astype(np.int64), rows)
        c = np.searchsorted(seg2_ids.astype(np.uint64), cols)
        mat = make_coo(r, c, vals, seg1_ids.size, seg2_ids.size)
    return mat, seg1_ids, seg2_ids
