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
    """
    Rows sharing a value of `new_id_colname` become one row (merge_df.py:13-43): size = sum, bbox = min / max,
    centroid = int(sum_i c_i * (s_i / S)) -- float64 products summed by pandas' groupby().sum() in frame order,
    exactly the reference's arithmetic, then truncated.  Columns other than the ten seg-info ones (e.g. the
    `max_overlap` column of cc_task(overlap_seg=...)) are carried over from the SMALLEST fragment of each merged
    id, which is what the reference's `sort_values([size]).drop_duplicates(new_id)` keeps.
    """
    ids = np.asarray(seginfo_df[cn.seg_id] if cn.seg_id in seginfo_df.columns else seginfo_df.index, dtype=np.int64)
    new = seginfo_df[new_id_colname].to_numpy(dtype=np.float64)
    dst = np.where(np.isnan(new), ids, np.nan_to_num(new)).astype(np.int64)
    tab = seginfo_df[cn.seg_info_cols].to_numpy(dtype=np.int64)
    work = pd.DataFrame({"dst": dst, cn.size: tab[:, 0]})
    grouped = work.groupby("dst", sort=True)
    sizes = grouped[cn.size].sum()
    uniq = sizes.index.to_numpy(dtype=np.int64)
    n = len(uniq)
    out = np.zeros((n, 10), dtype=np.int64)
    out[:, 0] = sizes.to_numpy(dtype=np.int64)
    w = tab[:, 0].astype(np.float64) / sizes.reindex(dst).to_numpy(dtype=np.float64)
    for a in range(3):
        work["c"] = tab[:, 1 + a].astype(np.float64) * w
        out[:, 1 + a] = work.groupby("dst", sort=True)["c"].sum().to_numpy().astype(np.int64)
        work["lo"], work["hi"] = tab[:, 4 + a], tab[:, 7 + a]
        out[:, 4 + a] = work.groupby("dst", sort=True)["lo"].min().to_numpy(dtype=np.int64)
        out[:, 7 + a] = work.groupby("dst", sort=True)["hi"].max().to_numpy(dtype=np.int64)
    df = pd.DataFrame(out, index=uniq, columns=cn.seg_info_cols)
    df.index.name = cn.seg_id
    extra = [c for c in seginfo_df.columns if c not in cn.seg_info_cols and c not in (new_id_colname, cn.seg_id)]
    if extra:
        src = seginfo_df[extra].copy()
        src["__dst"] = dst
        src["__size"] = tab[:, 0]
        first = src.sort_values(["__size"]).drop_duplicates("__dst").set_index("__dst")
        for c in extra:
            df[c] = first[c].reindex(uniq).to_numpy()
    return df


def enforce_size_threshold(seginfo_df, size_thr):
    """Drops rows with size < size_thr in place; returns {dropped id: 0}."""
    violations = seginfo_df.index[seginfo_df[cn.size] < size_thr]
    seginfo_df.drop(violations.tolist(), inplace=True)
    return {v: 0 for v in violations}
