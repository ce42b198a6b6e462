"""
Connected components -- signatures of synaptor/proc/seg/conncomps.py:11-52.

`connected_components` thresholds (strict `>`, float32) and labels into 3-D
components numbered 1..N in raster order of each component's first voxel (what
cc3d returns at conncomps.py:22), on the GPU: threshold->bitmask, z-run
union-find, popcount-scan canonical numbering, one dense label write.
"""
import numpy as np

from ... import _device as D


def _run(d, thresh, dil_k, connectivity, dust=0, overlap_seg=None):
    torch = D._torch()
    if not isinstance(d, torch.Tensor):
        d = np.asarray(d)
        if d.ndim != 3:
            raise ValueError(f"expected a 3-D volume, got shape {d.shape}")
        if d.size == 0:
            return None
        t32 = D.float32_threshold(d.dtype, thresh)
    else:
        t32 = np.float32(thresh)
    device = torch.device("cuda", torch.cuda.current_device())
    pred = D.to_device_f32(torch, d, device)
    ovl = None
    if overlap_seg is not None:
        if tuple(int(x) for x in overlap_seg.shape) != tuple(int(x) for x in pred.shape):
            raise ValueError("overlap_seg and the prediction volume must have the same shape")
        ovl = D.to_device_u64(torch, overlap_seg, device)
    return D.chunk_ccs_device(pred, t32, dust, connectivity=connectivity, dil_k=dil_k, overlap_seg=ovl)


def connected_components(d, thresh=0, overlap_seg=None, dtype=np.uint32, connectivity=6):
    """
    Threshold `d > thresh` and label the connected components (default 6-connectivity,
    the reference's hard-coded choice; 18 and 26 are extensions).  Returns a
    C-contiguous array of `dtype` (default uint32).
    With `overlap_seg` (conncomps.py:24-28) the labelling is multi-label: voxels above threshold join only
    when they carry the same non-zero `overlap_seg` id; voxels whose id is 0 are background.
    """
    res = _run(d, thresh, 0, connectivity, overlap_seg=overlap_seg)
    if res is None:
        return np.zeros(np.asarray(d).shape, dtype=dtype)
    return res.labels_numpy().astype(dtype, copy=False)


def _dilated_mask(output, cc_thresh, dil_param):
    torch = D._torch()
    d = np.asarray(output)
    if d.size == 0:
        return np.zeros(d.shape, dtype=np.uint32)
    t32 = D.float32_threshold(d.dtype, cc_thresh)
    pred = D.to_device_f32(torch, d, torch.device("cuda", torch.cuda.current_device()))
    return D.threshold_dilate_device(pred, t32, dil_param).cpu().numpy().view(np.uint32)


def dilated_components(output, dil_param, cc_thresh, connectivity=6):
    """
    Connected components with dilation (conncomps.py:32-52): dilate the thresholded
    mask by dil_param in 2-D Manhattan distance (axes 1, 2), label the dilated mask,
    zero the voxels that were not above threshold.
    """
    if dil_param == 0:
        return connected_components(output, cc_thresh, connectivity=connectivity)
    res = _run(output, cc_thresh, int(dil_param), connectivity)
    if res is None:
        return np.zeros(np.asarray(output).shape, dtype=np.uint32)
    return res.labels_numpy()
