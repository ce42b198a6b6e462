"""
Graph components + min-id map (reference: synaptor/proc/utils.py:39-86, which uses
python-igraph).  Here the union-find runs on the device (syn_union_min_map).
"""
import numpy as np

from .. import _device as D


def find_connected_components(matches):
    """List of id lists, one per connected component of the edge list `matches`."""
    id_map = make_id_map_from_edges(matches)
    comps = {}
    for (k, v) in id_map.items():
        comps.setdefault(v, []).append(k)
    return list(comps.values())


def make_id_map(ccs):
    """{id: min(id in its component)} (utils.py:77-86)."""
    mapping = {}
    for cc in ccs:
        target = min(cc)
        for i in cc:
            mapping[i] = target
    return mapping


def make_id_map_from_edges(matches):
    """Edges [(a, b)] -> {id: smallest id of its component} for every id in an edge (device union-find)."""
    if len(matches) == 0:
        return {}
    torch = D._torch()
    e = np.asarray(list(matches), dtype=np.int64).reshape(-1, 2)
    ids, inv = np.unique(e, return_inverse=True)   # compact ids keep their order, so min is preserved
    ce = inv.reshape(-1, 2).astype(np.uint32)
    device = torch.device("cuda", torch.cuda.current_device())
    edges = torch.from_numpy(ce.view(np.int32)).to(device)
    n_edges = torch.tensor([len(ce)], dtype=torch.int64, device=device)
    roots = D.union_min_map_device(edges, n_edges, len(ids)).cpu().numpy().view(np.uint32)
    return {int(k): int(ids[r]) for (k, r) in zip(ids, roots)}
