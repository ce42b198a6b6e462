"""Id-map helpers (reference: synaptor/proc/seg/merge/misc.py:8-34)."""
import pandas as pd

from ... import colnames as cn


def make_map_dframe(id_map):
    if len(id_map) == 0:
        return pd.DataFrame({cn.src_id: [], cn.dst_id: []})
    return pd.DataFrame({cn.src_id: list(id_map.keys()), cn.dst_id: list(id_map.values())})


def expand_id_map(id_map, all_ids):
    for i in set(all_ids).difference(id_map.keys()):
        id_map[i] = i
    return id_map


def update_id_map(id_map, next_map, reused_ids=False):
    """Composes id_map with next_map (reference: synaptor/proc/edge/merge/misc.py:4-19)."""
    new_keys = set(next_map.keys())
    for (k, v) in id_map.items():
        id_map[k] = next_map.get(v, v)
        new_keys.discard(v)
    if not reused_ids:
        for k in new_keys:
            id_map[k] = next_map[k]
    return id_map
