"""
Merging the per-chunk seg-info tables (reference: synaptor/proc/seg/merge/merge_df.py:9-68).
Per merged id: size = sum, bbox = min / max, centroid = int(sum_i c_i * (s_i / S))
evaluated in float64 and truncated.  Deliberate difference: the result is indexed
by integer id in ascending order (the reference's frame has a float64 index in
fragment-size order, an artefact of its NaN fill); contents per id are identical.
"""
import numpy as np
import pandas as pd

from ... import colnames as cn


def add_new_ids(seginfo_df, mapping, new_id_colname="new_ids"):
    seginfo_df[new_id_colname] = seginfo_df.index.map(mapping)


def merge_seginfo_df(seginfo_df, new_id_colname="new_ids"):
    # ids: the cleft_segid column when the frame carries one (tables read back from a store), else the index
    ids = np.asarray(seginfo_df[cn.seg_id] if cn.seg_id in seginfo_df.columns else seginfo_df.index, dtype=np.int64)
    new = seginfo_df[new_id_colname].to_numpy(dtype=np.float64)
    dst = np.where(np.isnan(new), ids, np.nan_to_num(new)).astype(np.int64)
    tab = seginfo_df[cn.seg_info_cols].to_numpy(dtype=np.int64)
    uniq, inv = np.unique(dst, return_inverse=True)
    n = len(uniq)
    out = np.zeros((n, 10), dtype=np.int64)
    np.add.at(out[:, 0], inv, tab[:, 0])
    w = tab[:, 0].astype(np.float64) / out[inv, 0].astype(np.float64)
    for a in range(3):
        acc = np.zeros(n, dtype=np.float64)
        np.add.at(acc, inv, tab[:, 1 + a].astype(np.float64) * w)
        out[:, 1 + a] = acc.astype(np.int64)
        lo = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(lo, inv, tab[:, 4 + a])
        out[:, 4 + a] = lo
        hi = np.full(n, np.iinfo(np.int64).min, dtype=np.int64)
        np.maximum.at(hi, inv, tab[:, 7 + a])
        out[:, 7 + a] = hi
    df = pd.DataFrame(out, index=uniq, columns=cn.seg_info_cols)
    df.index.name = cn.seg_id
    return df


def enforce_size_threshold(seginfo_df, size_thr):
    """Drops rows with size < size_thr in place; returns {dropped id: 0}."""
    violations = seginfo_df.index[seginfo_df[cn.size] < size_thr]
    seginfo_df.drop(violations.tolist(), inplace=True)
    return {v: 0 for v in violations}
