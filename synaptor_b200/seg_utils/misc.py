"""
dilate_by_k -- signature of synaptor/seg_utils/misc.py:28-46.  The chunk_ccs path
only ever dilates a boolean mask (proc/seg/conncomps.py:45), for which the
reference's chamfer distance transform (_seg_utils.cpp:25-82,180-200) equals a
binary dilation by the 2-D L1 ball over axes 1 and 2; that is what the device
kernel computes on the bitmask.  Label-propagating dilation of multi-id
segmentations is outside this path and raises.
"""
import numpy as np


def dilate_by_k(seg, k, copy=True):
    """0/1 mask dilated by k in 2-D Manhattan distance -> uint32 0/1 array (pybind forcecast result)."""
    from ..proc.seg import conncomps
    arr = np.asarray(seg)
    if arr.dtype != bool:
        vals = np.unique(arr)
        if not np.all(np.isin(vals, (0, 1))):
            raise NotImplementedError("dilate_by_k on multi-id segmentations is not on the chunk_ccs path")
    pred = arr.astype(np.float32)
    dil = conncomps._dilated_mask(pred, 0.5, int(k))
    return dil.astype(np.uint32)
