"""
Segment descriptors: sizes, centroids, bounding boxes.
Signatures of synaptor/seg_utils/describe.py:8-80; one fused CUDA pass
(syn_segment_stats) replaces np.unique, _describe.centers_of_mass
(_describe.cpp:23-64) and scipy.ndimage.find_objects.
"""
import numpy as np

from .. import _device as D
from ..types import BBox3d

# dense per-id table; refuse absurd id ranges instead of silently eating memory
MAX_DENSE_ID = 1 << 28


def _stats(seg):
    """-> (ids ascending int64, table rows [n, 11] int64) for the nonzero ids present."""
    torch = D._torch()
    seg = np.asarray(seg) if not isinstance(seg, torch.Tensor) else seg
    if seg.ndim != 3:
        raise ValueError(f"expected a 3-D segmentation, got shape {tuple(seg.shape)}")
    if seg.size == 0 if not isinstance(seg, torch.Tensor) else seg.numel() == 0:
        return np.zeros(0, dtype=np.int64), np.zeros((0, 11), dtype=np.int64)
    dev = D.to_device_u32(torch, seg, torch.device("cuda", torch.cuda.current_device()))
    max_id = D.max_u32_device(dev)
    if max_id > MAX_DENSE_ID:
        raise ValueError(f"segment ids up to {max_id} exceed the dense-table limit {MAX_DENSE_ID}")
    table = D.segment_stats_device(dev, max_id).cpu().numpy()
    present = np.nonzero(table[:, 1])[0]
    present = present[present != 0]
    return present.astype(np.int64), table[present]


def nonzero_unique_ids(seg):
    """Sorted nonzero ids of a segmentation (describe.py:8-11)."""
    seg = np.asarray(seg)
    if seg.dtype.kind in "iu" and seg.dtype.itemsize == 8:
        from .overlap import unique_u64
        ids = unique_u64(seg)
        return ids[ids != 0].astype(seg.dtype)
    ids, _ = _stats(seg)
    return ids.astype(seg.dtype if seg.dtype.kind in "iu" else np.uint32)


def centers_of_mass(seg, offset=(0, 0, 0)):
    """{id: (x, y, z)}; float32 `round(sum / size)` per axis + offset (describe.py:14-32)."""
    ids, rows = _stats(seg)
    cents = {int(i): (int(r[2]), int(r[3]), int(r[4])) for (i, r) in zip(ids, rows)}
    if offset != (0, 0, 0):
        cents = {i: (c[0] + offset[0], c[1] + offset[1], c[2] + offset[2]) for (i, c) in cents.items()}
    return cents


def bounding_boxes(ccs, offset=(0, 0, 0)):
    """{id: BBox3d} with exclusive upper corner, shifted by offset (describe.py:35-59)."""
    ids, rows = _stats(ccs)
    boxes = {int(i): BBox3d((int(r[5]), int(r[6]), int(r[7])), (int(r[8]), int(r[9]), int(r[10])))
             for (i, r) in zip(ids, rows)}
    if offset == (0, 0, 0):
        return boxes
    return {i: b.translate(offset) for (i, b) in boxes.items()}


def segment_sizes(seg):
    """{id: voxel count} for nonzero ids, ascending id order (describe.py:62-80)."""
    ids, rows = _stats(seg)
    key = np.asarray(seg).dtype.type if np.asarray(seg).dtype.kind in "iu" else np.uint32
    return {key(i): np.int64(r[1]) for (i, r) in zip(ids, rows)}
