"""Overlap helpers -- signatures of synaptor/proc/overlap/overlap.py:12-52."""
import numpy as np
import pandas as pd
from scipy import sparse

from ... import seg_utils
from .. import colnames as cn

count_overlaps = seg_utils.count_overlaps


def find_max_overlaps(overlap_mat):
    """Per row the column with the largest count; the FIRST entry wins ties (entry order matters)."""
    r = np.asarray(overlap_mat.row)
    c = np.asarray(overlap_mat.col)
    v = np.asarray(overlap_mat.data)
    if len(r) == 0:
        # (the reference's `coo_matrix(([], ([], [])))` raises on current scipy; an empty matrix is returned instead)
        return sparse.coo_matrix((0, 0), dtype=np.int64)
    # stable sort by (row, -value) keeps entry order among equal values; first of each row wins
    order = np.lexsort((np.arange(len(r)), -v.astype(np.int64), r))
    rs = r[order]
    firsts = np.concatenate(([True], rs[1:] != rs[:-1]))
    pick = order[firsts]
    # the reference emits rows in order of their first appearance in the entry list
    _, first_pos = np.unique(r, return_index=True)
    pick = pick[np.argsort(first_pos, kind="stable")]
    return sparse.coo_matrix((v[pick], (r[pick], c[pick])))


def convert_to_dict(overlap_mat):
    return dict(zip(overlap_mat.row, overlap_mat.col))


def add_overlapping_seg(dframe, seg, base_seg, colname=cn.ovl_segid):
    overlaps = count_overlaps(seg, base_seg, orig_ids=True)[0]
    max_overlaps = convert_to_dict(find_max_overlaps(overlaps))
    overlap_df = pd.DataFrame.from_dict(max_overlaps, orient="index", columns=[colname])
    new_dframe = pd.concat((dframe, overlap_df), axis=1)
    new_dframe.index.name = dframe.index.name
    return new_dframe
