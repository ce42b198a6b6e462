"""Seg-info DataFrame builder -- signatures of synaptor/proc/seg/misc.py:14-47."""
import numpy as np
import pandas as pd

from .. import colnames as cn


def empty_cleft_df(index_name=cn.seg_id):
    df = pd.DataFrame({k: [] for k in cn.seg_info_cols})
    df.index.name = index_name
    return df


def make_seg_info_dframe(centers, sizes, bboxes, index_name=cn.seg_id):
    """
    DataFrame indexed by segment id with columns size, centroid_{x,y,z},
    bbox_{b,e}{x,y,z} (all int64) from the three per-id dicts.
    """
    assert len(centers) == len(sizes) == len(bboxes), "mismatched inputs"
    if len(centers) == 0:
        return empty_cleft_df(index_name)
    ids = list(sizes.keys())
    table = np.empty((len(ids), 10), dtype=np.int64)
    for (r, i) in enumerate(ids):
        table[r, 0] = sizes[i]
        table[r, 1:4] = centers[i]
        table[r, 4:10] = bboxes[i].astuple()
    return seg_info_from_table(np.asarray(ids), table, index_name)


def seg_info_from_table(ids, table, index_name=cn.seg_id):
    """ids (n,), table (n, 10) int64 in column order -> the same DataFrame, without the dict detour."""
    if len(ids) == 0:
        return empty_cleft_df(index_name)
    df = pd.DataFrame(np.asarray(table, dtype=np.int64), index=np.asarray(ids), columns=cn.seg_info_cols)
    df.index.name = index_name
    return df
