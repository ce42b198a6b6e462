"""Merging per-chunk overlap matrices (reference: synaptor/proc/overlap/merge/merge_overlaps.py:10-60)."""
import numpy as np
import scipy.sparse as sp


def apply_chunk_id_maps(overlap_arr, chunk_id_maps):
    """Rows of every chunk matrix mapped through its chunk id map, all entries concatenated."""
    rs, cs, vs = [], [], []
    for (ov, id_map) in zip(overlap_arr.flat, chunk_id_maps.flat):
        if ov.nnz == 0:
            continue
        rs.append(np.array([id_map[r] for r in ov.row], dtype=np.int64))
        cs.append(np.asarray(ov.col, dtype=np.int64))
        vs.append(np.asarray(ov.data))
    if not rs:
        return sp.coo_matrix((0, 0), dtype=np.int64)
    return sp.coo_matrix((np.concatenate(vs), (np.concatenate(rs), np.concatenate(cs))))


def consolidate_overlaps(overlap_mat_arr, dtype=np.uint64):
    """One COO matrix holding the entries of all chunk matrices with duplicates summed."""
    rs, cs, vs = [], [], []
    for m in overlap_mat_arr.flat:
        r, c, v = sp.find(m)
        rs.append(r.astype(dtype)); cs.append(c.astype(dtype)); vs.append(v.astype(dtype))
    if not rs:
        return sp.coo_matrix((0, 0), dtype=np.int64)
    full = sp.coo_matrix((np.concatenate(vs), (np.concatenate(rs), np.concatenate(cs))))
    full.sum_duplicates()
    return full
