"""
relabel_data -- signature of synaptor/seg_utils/relabel.py:7-27; the per-voxel
hash lookup of _relabel.cpp:19-38 becomes a dense-LUT gather kernel
(syn_relabel_lut).
"""
import numpy as np

from .. import _device as D


def build_lut(mapping, max_value):
    """Dense uint32 table for `mapping` restricted to keys <= max_value (identity elsewhere)."""
    n = int(max_value) + 1
    lut = np.arange(n, dtype=np.uint32)
    if len(mapping):
        keys = np.fromiter(mapping.keys(), dtype=np.int64, count=len(mapping))
        vals = np.fromiter(mapping.values(), dtype=np.int64, count=len(mapping))
        ok = (keys >= 0) & (keys < n)
        lut[keys[ok]] = vals[ok].astype(np.uint32)
    return lut


def relabel_data(d, mapping, copy=True):
    """
    Values of `d` that are keys of `mapping` are replaced, others stay
    (`if v in mapping: v = mapping[v]`, every voxel including zeros).
    With copy=False and a uint32 C-contiguous numpy array the result is written
    back into `d` (the reference's in-place contract for exact-dtype input).
    """
    torch = D._torch()
    if isinstance(d, torch.Tensor):
        dev = d if not copy else d.clone()
        mx = D.max_u32_device(dev.view(torch.int32))
        lut = torch.from_numpy(build_lut(mapping, mx).view(np.int32)).to(dev.device)
        D.relabel_lut_device(dev.view(torch.int32), lut)
        return dev
    arr = np.asarray(d)
    if arr.size == 0 or len(mapping) == 0:
        return np.copy(arr) if copy else arr
    dev = D.to_device_u32(torch, arr, torch.device("cuda", torch.cuda.current_device()))
    mx = D.max_u32_device(dev)
    lut = torch.from_numpy(build_lut(mapping, mx).view(np.int32)).to(dev.device)
    D.relabel_lut_device(dev, lut)
    out = dev.cpu().numpy().view(np.uint32).reshape(arr.shape)
    if not copy and arr.dtype == np.uint32 and isinstance(d, np.ndarray) and d.flags.writeable:
        d[...] = out
        return d
    return out


def relabel_data_1N(d, copy=True):
    """Relabel the nonzero ids to 1..N in ascending id order (relabel.py:30-43)."""
    from .describe import nonzero_unique_ids
    mapping = {int(v): i + 1 for (i, v) in enumerate(nonzero_unique_ids(d))}
    return relabel_data(d, mapping, copy=copy)
