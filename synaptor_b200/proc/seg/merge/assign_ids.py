"""Global id assignment for the chunk tables (reference: synaptor/proc/seg/merge/assign_ids.py:8-84)."""
import numpy as np
import pandas as pd


def empty_obj_array(shape):
    arr = np.empty(int(np.prod(shape)), dtype=object)
    return arr.reshape(shape)


def new_id_map(df, next_id):
    """{chunk-local id: global id} for the rows of df in table order, starting at next_id."""
    segids = df.index.tolist()
    return dict(zip(segids, range(next_id, next_id + len(segids)))), next_id + len(segids)


def assign_unique_ids_serial(cleft_info_arr):
    """
    Walks the chunk grid in np.ndindex order and numbers every table row 1, 2, ...
    Returns (one concatenated DataFrame indexed by global id, object array of
    {local id: global id}).  Like the reference, the per-chunk frames are re-indexed
    in place.
    """
    chunk_id_maps = empty_obj_array(cleft_info_arr.shape)
    parts = []
    next_id = 1
    for idx in np.ndindex(cleft_info_arr.shape):
        df = cleft_info_arr[idx]
        chunk_id_maps[idx], next_id = new_id_map(df, next_id)
        df.rename(chunk_id_maps[idx], inplace=True)
        parts.append(df)
    return pd.concat(parts), chunk_id_maps


def apply_id_map(continuations, id_map):
    """Rewrites Continuation.segid through id_map, in place; accepts a list or a {face: list} dict."""
    lists = continuations.values() if isinstance(continuations, dict) else [continuations]
    for conts in lists:
        for c in conts:
            c.segid = id_map[c.segid]


def apply_chunk_id_maps(continuation_arr, chunk_id_maps):
    for (c_dict, id_map) in zip(continuation_arr.flat, chunk_id_maps.flat):
        apply_id_map(c_dict, id_map)
    return continuation_arr


def update_chunk_id_maps(chunk_id_maps, cont_id_map):
    """Composes every chunk map with cont_id_map (ids without an entry stay)."""
    for mapping in chunk_id_maps.flat:
        for (k, v) in mapping.items():
            mapping[k] = cont_id_map.get(v, v)
    return chunk_id_maps
