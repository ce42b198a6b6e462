"""
Device-level driver of the CUDA pipeline.

torch is used ONLY as the allocator / buffer interchange (device memory, pinned
staging, streams); every computation happens in libsynaptor_b200.so, reached
through ctypes (`_cabi`).  There is no CPU fallback anywhere in this module.
"""
import ctypes
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _cabi as C

_ws_cache = {}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("synaptor_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def _stream_ptr(torch, stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _workspace(torch, device, nbytes):
    # one workspace per (device, host thread)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
           threading.get_ident())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        _ws_cache.pop(key, None)
        ws = None
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def release_workspaces():
    _ws_cache.clear()


_side_streams = {}


def side_stream(torch, device, name, shared=False):
    """Per-device auxiliary torch streams ("in": H2D copies, "out": D2H copies) for the host-array entry points.
    Per host thread unless `shared`."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), name,
           0 if shared else threading.get_ident())
    s = _side_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _side_streams[key] = s
    return s


def pinned_like(torch, t):
    return torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)


# ---- pinned staging blocks for the label volume coming back ------------------------------------
# A 512^3 label volume is a 537 MB pinned block.  cudaHostAlloc of that size takes ~200 ms, and torch's caching
# host allocator does not reliably hand a just-freed block back under the multi-stream pattern of the pipelined
# path (measured: stalls of 200 ms every few chunks), so the blocks are pooled here: a block goes back to the
# pool when the numpy array that was returned to the caller is garbage collected.
_pinned_pool = {}
_pinned_lock = threading.Lock()
_PINNED_POOL_MAX = 8
# page-locked bytes the pool may HOLD (free blocks only; blocks behind arrays the caller still references are the
# caller's): beyond it a released block is simply freed.  SYNAPTOR_B200_PINNED_POOL_GB overrides (0 = no pooling).
_PINNED_POOL_MAX_BYTES = int(float(os.environ.get("SYNAPTOR_B200_PINNED_POOL_GB", "16")) * (1 << 30))
_pinned_pool_bytes = 0


def pinned_acquire(torch, shape, dtype):
    global _pinned_pool_bytes
    key = (tuple(int(x) for x in shape), dtype)
    with _pinned_lock:
        free = _pinned_pool.get(key)
        if free:
            h = free.pop()
            _pinned_pool_bytes -= h.numel() * h.element_size()
            return h
    return torch.empty(key[0], dtype=dtype, device="cpu", pin_memory=True)


def pinned_reserve(torch, shape, dtype, n):
    """Makes sure `n` blocks of this shape are in the pool (cudaHostAlloc of a 512^3 label block takes ~0.3 s:
    better before a pipelined run than in the middle of it)."""
    key = (tuple(int(x) for x in shape), dtype)
    with _pinned_lock:
        have = len(_pinned_pool.get(key, []))
    fresh = [torch.empty(key[0], dtype=dtype, device="cpu", pin_memory=True) for _ in range(max(0, n - have))]
    for h in fresh:
        pinned_release(h)


def pinned_release(host):
    global _pinned_pool_bytes
    key = (tuple(host.shape), host.dtype)
    nbytes = host.numel() * host.element_size()
    with _pinned_lock:
        free = _pinned_pool.setdefault(key, [])
        if len(free) < _PINNED_POOL_MAX and _pinned_pool_bytes + nbytes <= _PINNED_POOL_MAX_BYTES:
            free.append(host)
            _pinned_pool_bytes += nbytes


def pinned_numpy(host, view_dtype):
    """numpy view of a pooled pinned block; the block returns to the pool when the array (and every view of it)
    is gone."""
    import weakref
    arr = host.numpy().view(view_dtype)
    owner = arr.base if arr.base is not None else arr
    weakref.finalize(owner, pinned_release, host)
    return arr


def float32_threshold(dtype, thresh):
    """
    The float32 value t such that `float32(d) > t` equals numpy's `d > thresh` for
    every element of an array of `dtype` (synaptor/proc/seg/conncomps.py:17).
    float32 arrays: numpy casts the python scalar to float32 (round to nearest).
    float16 / small-int arrays: every element is exact in float32, so the threshold
    is rounded DOWN to float32 (d > thresh  <=>  d > largest float32 <= thresh).
    """
    dtype = np.dtype(dtype)
    if isinstance(thresh, np.generic) and not isinstance(thresh, (np.float32, np.float16)) and dtype == np.float32 \
            and np.dtype(type(thresh)).itemsize > 4 and np.dtype(type(thresh)).kind == "f":
        raise TypeError("a float64 numpy scalar threshold promotes the comparison to float64; pass a python float")
    if dtype == np.float32:
        return np.float32(thresh)
    if dtype == np.float16:
        return np.float32(np.float16(thresh))
    if dtype.kind in "biu" and dtype.itemsize <= 2:
        t = np.float32(thresh)
        if float(t) > float(thresh):
            t = np.nextafter(t, np.float32(-np.inf))
        return t
    raise TypeError(f"unsupported prediction dtype {dtype}: the device path compares in float32 "
                    "(float32, float16, bool and 8/16-bit integer inputs are exact)")


def to_device_f32(torch, d, device):
    """numpy / torch prediction volume -> contiguous float32 CUDA tensor [X,Y,Z]."""
    if isinstance(d, torch.Tensor):
        if d.dtype != torch.float32:
            raise TypeError("CUDA/torch predictions must be float32")
        return d.to(device, non_blocking=True).contiguous()
    d = np.asarray(d)
    if d.ndim != 3:
        raise ValueError(f"expected a 3-D volume, got shape {d.shape}")
    if d.dtype != np.float32:
        float32_threshold(d.dtype, 0)  # raises for inexact dtypes
    # conncomps.py:21 makes the (F-ordered, CloudVolume) chunk C-contiguous on the host; here the bytes go to the
    # device as they are and the transpose runs there
    out = torch.empty(tuple(int(s) for s in d.shape), dtype=torch.float32, device=device)
    return upload_f32(torch, out, d)


def to_device_u64(torch, a, device):
    if isinstance(a, torch.Tensor):
        if a.dtype not in (torch.int64, torch.uint64):
            raise TypeError("base segmentation tensors must be (u)int64")
        return a.view(torch.int64).to(device, non_blocking=True).contiguous()
    a = np.asarray(a)
    if a.dtype.kind not in "iu":
        raise TypeError(f"base segmentation must be an integer array, got {a.dtype}")
    if a.ndim != 3:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(device, non_blocking=True)
    out = torch.empty(tuple(int(s) for s in a.shape), dtype=torch.int64, device=device)
    return upload_u64(torch, out, a)


def to_device_u32(torch, a, device):
    if isinstance(a, torch.Tensor):
        if a.dtype not in (torch.int32, torch.uint32):
            raise TypeError("label tensors must be (u)int32")
        return a.view(torch.int32).to(device, non_blocking=True).contiguous()
    a = np.asarray(a)
    if a.dtype != np.uint32:
        if a.dtype.kind not in "iub":
            raise TypeError(f"label volume must be an integer array, got {a.dtype}")
        a = a.astype(np.uint32)  # pybind forcecast semantics of the reference's uint32 overloads
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32)).to(device, non_blocking=True)


# ---- pageable host arrays ---------------------------------------------------------------------------
# A copy from pageable memory goes through the driver's own bounce buffer on the calling thread (measured: ~10 GB/s,
# 3.7x slower than from pinned memory).  What the reference's caller hands over IS pageable (numpy arrays from
# CloudVolume), so large pageable arrays are staged here instead: a few host threads fill pinned staging blocks
# (numpy copies release the GIL) while the DMA of the previous block runs.
_STAGE_BYTES = 64 << 20
_STAGE_BLOCKS = 4
_STAGE_THREADS = 8
_stagers = {}


class _Stager:
    def __init__(self, torch):
        from concurrent.futures import ThreadPoolExecutor
        self.blocks = [torch.empty(_STAGE_BYTES, dtype=torch.uint8, device="cpu", pin_memory=True)
                       for _ in range(_STAGE_BLOCKS)]
        self.events = [None] * _STAGE_BLOCKS
        self.next = 0
        self.pool = ThreadPoolExecutor(max_workers=_STAGE_THREADS)

    def copy(self, torch, dst_bytes, src_bytes):
        """dst_bytes: flat uint8 cuda tensor; src_bytes: flat uint8 numpy view of a pageable array."""
        n = src_bytes.size
        for off in range(0, n, _STAGE_BYTES):
            m = min(_STAGE_BYTES, n - off)
            k = self.next
            self.next = (k + 1) % _STAGE_BLOCKS
            if self.events[k] is not None:
                self.events[k].synchronize()             # the DMA that last read this block is done
            blk = self.blocks[k].numpy()
            step = -(-m // _STAGE_THREADS)
            step = (step + 4095) & ~4095
            list(self.pool.map(lambda a: np.copyto(blk[a:min(a + step, m)], src_bytes[off + a:off + min(a + step, m)]),
                               range(0, m, step)))
            dst_bytes[off:off + m].copy_(self.blocks[k][:m], non_blocking=True)
            e = torch.cuda.Event()
            e.record(torch.cuda.current_stream())
            self.events[k] = e


def _copy_host_to_device(torch, dst, t):
    """dst (cuda) <- t (CPU tensor sharing a numpy array's memory), on the current stream."""
    nbytes = t.numel() * t.element_size()
    if nbytes >= (8 << 20) and not t.is_pinned():
        key = (dst.device.index, threading.get_ident())
        st = _stagers.get(key)
        if st is None:
            st = _stagers[key] = _Stager(torch)
        st.copy(torch, dst.view(-1).view(torch.uint8), t.numpy().reshape(-1).view(np.uint8))
    else:
        dst.copy_(t, non_blocking=True)


def _host_tensor(torch, a, np_dtype, torch_view=None):
    """numpy array -> (CPU tensor sharing its memory, f_order flag).  C-contiguous arrays are used as they are;
    F-contiguous ones (what CloudVolume hands the reference's tasks, io/backends/cloudvolume.py:6-20) are taken
    as their transposed C-contiguous view [Z, Y, X] -- the bytes are copied unchanged and the transpose happens
    on the device; anything else is made C-contiguous on the host (conncomps.py:21)."""
    a = np.asarray(a)
    if a.dtype != np_dtype:
        a = a.astype(np_dtype)
    f_order = False
    if not a.flags.c_contiguous:
        if a.ndim == 3 and a.flags.f_contiguous:
            a, f_order = a.T, True
        else:
            a = np.ascontiguousarray(a)
    if torch_view is not None:
        a = a.view(torch_view)
    return torch.from_numpy(a), f_order


def _transpose_into(torch, dst, src_zyx):
    """dst [X,Y,Z] <- src [Z,Y,X] on the device (syn_transpose_zyx: tiled through shared memory)."""
    nx, ny, nz = (int(s) for s in dst.shape)
    with torch.cuda.device(dst.device):
        C.check(C.lib().syn_transpose_zyx(src_zyx.data_ptr(), dst.data_ptr(), nx, ny, nz, int(dst.element_size()),
                                          _stream_ptr(torch)))
    return dst


def upload_f32(torch, dst, src):
    """Copies a host prediction chunk (numpy, any order / pinned or pageable; or a tensor) into the resident
    float32 cuda tensor `dst` on the current stream."""
    if isinstance(src, torch.Tensor):
        dst.copy_(src, non_blocking=True)
        return dst
    if np.asarray(src).dtype != np.float32:
        float32_threshold(np.asarray(src).dtype, 0)          # raises for inexact dtypes
    t, f_order = _host_tensor(torch, src, np.float32)
    if tuple(t.shape) != (tuple(dst.shape)[::-1] if f_order else tuple(dst.shape)):
        raise ValueError(f"chunk shape {tuple(np.asarray(src).shape)} != {tuple(dst.shape)}")
    if not f_order:
        _copy_host_to_device(torch, dst, t)
        return dst
    stage = torch.empty(t.shape, dtype=dst.dtype, device=dst.device)
    _copy_host_to_device(torch, stage, t)
    return _transpose_into(torch, dst, stage)


def upload_u64(torch, dst, src):
    """As upload_f32 for uint64 ids into an int64-typed cuda tensor."""
    if isinstance(src, torch.Tensor):
        dst.copy_(src.view(torch.int64), non_blocking=True)
        return dst
    a = np.asarray(src)
    if a.dtype.kind not in "iu":
        raise TypeError(f"base segmentation must be an integer array, got {a.dtype}")
    t, f_order = _host_tensor(torch, a, np.uint64, np.int64)
    if tuple(t.shape) != (tuple(dst.shape)[::-1] if f_order else tuple(dst.shape)):
        raise ValueError(f"chunk shape {tuple(a.shape)} != {tuple(dst.shape)}")
    if not f_order:
        _copy_host_to_device(torch, dst, t)
        return dst
    stage = torch.empty(t.shape, dtype=dst.dtype, device=dst.device)
    _copy_host_to_device(torch, stage, t)
    return _transpose_into(torch, dst, stage)


@dataclass
class ChunkResult:
    """Device-resident outputs of one syn_chunk_ccs call."""
    labels: object                 # cuda int32 tensor [X,Y,Z] holding uint32 bit patterns
    seg_table: object              # cuda int64 [max_segs, 11]; first counts['kept'] rows valid
    pair_rows: object = None       # cuda int32 (uint32 bits)
    pair_cols: object = None       # cuda int64 (uint64 bits)
    pair_vals: object = None       # cuda int64
    counts: dict = field(default_factory=dict)
    workspace: object = None
    caps: tuple = None             # (max_runs, max_segs, max_pairs) this result was produced with

    def table_numpy(self):
        n = self.counts["kept"]
        return self.seg_table[:n].cpu().numpy()

    def labels_numpy(self):
        """D2H through pinned memory (torch's caching host allocator recycles the block)."""
        import torch
        host = pinned_acquire(torch, self.labels.shape, self.labels.dtype)
        host.copy_(self.labels)
        return pinned_numpy(host, np.uint32)

    def labels_numpy_async(self, stream):
        """Starts the D2H of the label volume on `stream` (after the work queued on the current stream);
        returns (pinned host tensor, event).  Synchronise the event before touching the array."""
        import torch
        cur = torch.cuda.current_stream()
        host = pinned_acquire(torch, self.labels.shape, self.labels.dtype)
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            host.copy_(self.labels, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        self.labels.record_stream(stream)
        return host, ev

    def pairs_numpy(self):
        n = self.counts["pairs"]
        return (self.pair_rows[:n].cpu().numpy().view(np.uint32).astype(np.int64),
                self.pair_cols[:n].cpu().numpy().view(np.uint64),
                self.pair_vals[:n].cpu().numpy())


def _counts_dict(arr):
    return {
        "runs": int(arr[C.CNT_RUNS]), "components": int(arr[C.CNT_COMPONENTS]), "kept": int(arr[C.CNT_KEPT]),
        "pairs": int(arr[C.CNT_PAIRS]), "max_base": int(np.uint64(arr[C.CNT_MAX_BASE].view(np.uint64))),
        "max_label": int(arr[C.CNT_MAX_LABEL]), "errflags": int(arr[C.CNT_ERRFLAGS]),
        "fg_voxels": int(arr[C.CNT_FG_VOXELS]), "table_slots": int(arr[C.CNT_TABLE_SLOTS]),
        "active_words": int(arr[C.CNT_ACTIVE_WORDS]), "tile_records": int(arr[C.CNT_TILE_RECORDS]),
        "flagged_tiles": int(arr[C.CNT_FLAGGED_TILES]), "bnd_words": int(arr[C.CNT_BND_WORDS]),
        "dbg": [int(v) for v in arr[C.CNT_DBG0:]],
    }


def default_capacities(shape):
    nvox = int(shape[0]) * int(shape[1]) * int(shape[2])
    worst_runs = int(shape[0]) * int(shape[1]) * ((int(shape[2]) + 1) // 2)
    max_runs = min(max(1 << 16, nvox // 8), max(worst_runs, 1))
    max_segs = min(max(1 << 14, nvox // 256), max(worst_runs, 1))
    max_pairs = max(1 << 14, nvox // 1024)   # grown on overflow (SYN_ERR_CAPACITY retry)
    return max_runs, max_segs, max_pairs


def chunk_ccs_device(pred, thresh, dust_thresh, connectivity=6, dil_k=0, offset=(0, 0, 0), base=None,
                     stream=None, sync=True, max_runs=None, max_segs=None, max_pairs=None, out=None,
                     overlap_seg=None):
    """
    Runs the fused chunk_ccs (+ chunk_overlaps when `base` is given) pipeline on
    CUDA tensors that are already resident.  `pred`: float32 [X,Y,Z] cuda tensor;
    `base`: int64-typed cuda tensor holding uint64 ids; `thresh`: float32 value.
    With sync=True (default) capacities are grown and the call retried on overflow,
    and `counts` is filled; with sync=False nothing is read back (benchmark mode:
    use read_counts() afterwards).
    `out`: optional ChunkResult from a previous call with the same shape to reuse
    its output buffers.
    `overlap_seg`: int64-typed cuda tensor of uint64 ids -> the multi-label variant
    (conncomps.py:24-28): voxels join only when they share a non-zero base id.
    """
    torch = _torch()
    L = C.lib()
    if not (isinstance(pred, torch.Tensor) and pred.is_cuda and pred.dtype == torch.float32 and pred.dim() == 3
            and pred.is_contiguous()):
        raise TypeError("pred must be a contiguous 3-D float32 CUDA tensor")
    nx, ny, nz = (int(s) for s in pred.shape)
    device = pred.device
    d_runs, d_segs, d_pairs = default_capacities((nx, ny, nz))
    max_runs = int(max_runs or d_runs)
    max_segs = int(max_segs or d_segs)
    max_pairs = int(max_pairs or d_pairs) if base is not None else 0
    if base is not None and not (isinstance(base, torch.Tensor) and base.is_cuda and base.dtype == torch.int64
                                 and base.is_contiguous() and tuple(base.shape) == (nx, ny, nz)
                                 and base.device == pred.device):
        raise TypeError("base must be a contiguous int64-typed CUDA tensor of the prediction's shape and device")
    if out is not None and (tuple(out.labels.shape) != (nx, ny, nz) or out.labels.device != pred.device
                            or out.seg_table.device != pred.device):
        out = None          # a result of another shape / device cannot lend its buffers
    off = (ctypes.c_int64 * 3)(int(offset[0]), int(offset[1]), int(offset[2]))
    counts = np.zeros(C.NCOUNTS, dtype=np.int64)

    for _attempt in range(8):
        with torch.cuda.device(device):
            nbytes = L.syn_ccs_workspace_bytes(nx, ny, nz, max_runs, max_pairs)
            if nbytes < 0:
                C.check(int(nbytes))
            ws = _workspace(torch, device, nbytes)
            if out is not None and out.seg_table.shape[0] >= max_segs and (
                    base is None or (out.pair_rows is not None and out.pair_rows.numel() >= max_pairs)):
                labels, table = out.labels, out.seg_table
                prow, pcol, pval = out.pair_rows, out.pair_cols, out.pair_vals
            else:
                labels = torch.empty((nx, ny, nz), dtype=torch.int32, device=device)
                table = torch.empty((max_segs, C.SEG_NCOLS), dtype=torch.int64, device=device)
                prow = pcol = pval = None
                if base is not None:
                    prow = torch.empty(max_pairs, dtype=torch.int32, device=device)
                    pcol = torch.empty(max_pairs, dtype=torch.int64, device=device)
                    pval = torch.empty(max_pairs, dtype=torch.int64, device=device)
            if overlap_seg is not None:
                if base is not None:
                    raise ValueError("overlap_seg and base are mutually exclusive")
                if not (isinstance(overlap_seg, torch.Tensor) and overlap_seg.is_cuda and overlap_seg.dtype == torch.int64
                        and overlap_seg.is_contiguous() and tuple(overlap_seg.shape) == (nx, ny, nz)):
                    raise TypeError("overlap_seg must be a contiguous int64-typed CUDA tensor of the prediction's shape")
                if int(connectivity) != 6 or int(dil_k) != 0:
                    raise ValueError("cc_task(overlap_seg=...) is 6-connected and has no dilation (conncomps.py:28)")
                rc = L.syn_chunk_ccs_multilabel(
                    pred.data_ptr(), overlap_seg.data_ptr(), nx, ny, nz, ctypes.c_float(float(thresh)),
                    int(dust_thresh), ctypes.cast(off, ctypes.c_void_p), labels.data_ptr(), table.data_ptr(),
                    table.shape[0], ws.data_ptr(), ws.numel(), max_runs,
                    counts.ctypes.data_as(ctypes.c_void_p) if sync else None, _stream_ptr(torch, stream))
            else:
                rc = L.syn_chunk_ccs(
                    pred.data_ptr(), nx, ny, nz, ctypes.c_float(float(thresh)), int(connectivity), int(dil_k),
                    int(dust_thresh), ctypes.cast(off, ctypes.c_void_p), labels.data_ptr(), table.data_ptr(),
                    table.shape[0], base.data_ptr() if base is not None else None,
                    prow.data_ptr() if prow is not None else None, pcol.data_ptr() if pcol is not None else None,
                    pval.data_ptr() if pval is not None else None, prow.numel() if prow is not None else 0,
                    ws.data_ptr(), ws.numel(), max_runs,
                    counts.ctypes.data_as(ctypes.c_void_p) if sync else None, _stream_ptr(torch, stream))
        res = ChunkResult(labels, table, prow, pcol, pval, _counts_dict(counts) if sync else {}, ws,
                          (max_runs, max_segs, max_pairs))
        if rc == C.SYN_OK:
            return res
        if rc != C.SYN_ERR_CAPACITY:
            C.check(rc)
        cd = res.counts
        ef = cd["errflags"]
        out = None
        if ef & C.EF_RUNS:
            max_runs = max(cd["runs"], max_runs + 1)
        if ef & C.EF_SEGS:
            max_segs = max(cd["kept"], max_segs + 1)
        if ef & C.EF_TABLE:
            max_pairs = max_pairs * 4
        if ef & C.EF_PAIRS:
            max_pairs = max(cd["pairs"], max_pairs + 1)
    raise C.CapacityError(C.SYN_ERR_CAPACITY, "capacities still too small after 8 attempts", counts)


class ChunkPipeline:
    """
    Fixed-buffer chunk_ccs (+ chunk_overlaps) for a stream of equally shaped chunks.

    The 13 kernel launches + memset of one syn_chunk_ccs call are captured ONCE into a CUDA graph (stream
    capture of the same C-ABI call, on the pipeline's own stream) and replayed per chunk: the ~25 us of launch
    gaps of a 0.63 ms step disappear.  The caller fills `pred` (float32) and `base` (int64 view of the uint64 ids)
    -- e.g. `pipe.pred.copy_(host_tensor, non_blocking=True)` -- calls run(), and reads `result` (device tensors;
    `counts()` synchronises and returns the counters, growing the capacities and re-capturing on overflow).
    """

    def __init__(self, shape, dust_thresh, thresh=0.5, connectivity=6, dil_k=0, with_base=True, offset=(0, 0, 0),
                 device=None, pred=None, base=None):
        torch = _torch()
        self.torch = torch
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.shape = tuple(int(x) for x in shape)
        self.args = dict(thresh=np.float32(thresh), dust_thresh=int(dust_thresh), connectivity=int(connectivity),
                         dil_k=int(dil_k), offset=tuple(int(o) for o in offset))
        # input buffers: the pipeline's own, or existing resident tensors handed in by the caller
        self.pred = pred if pred is not None else torch.empty(self.shape, dtype=torch.float32, device=self.device)
        self.base = base if base is not None else (
            torch.empty(self.shape, dtype=torch.int64, device=self.device) if with_base else None)
        if not (tuple(self.pred.shape) == self.shape and self.pred.dtype == torch.float32 and self.pred.is_cuda):
            raise TypeError("pred must be a float32 CUDA tensor of the pipeline's shape")
        self.stream = torch.cuda.Stream(device=self.device)
        self.result = None
        self.graph = None
        self.launches_per_run = 0

    def _eager(self, sync):
        a = self.args
        caps = self.result.caps if self.result is not None else (None, None, None)
        return chunk_ccs_device(self.pred, a["thresh"], a["dust_thresh"], connectivity=a["connectivity"],
                                dil_k=a["dil_k"], offset=a["offset"], base=self.base, sync=sync, out=self.result,
                                max_runs=caps[0], max_segs=caps[1], max_pairs=caps[2] or None)

    def _capture(self):
        torch = self.torch
        L = C.lib()
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.result = self._eager(sync=True)          # sizes the buffers (capacity retries happen here)
            if not getattr(self, "_tight", False):
                # capacities from what this stream of chunks actually holds (+ headroom) instead of the worst-case
                # defaults: smaller per-run arrays and a pair table the ordering kernels walk in a fraction of the
                # time; counts() grows them again on an overflow
                c = self.result.counts
                caps = (max(1 << 16, int(c["runs"] * 1.25) + 1024), max(1 << 12, int(c["kept"] * 1.25) + 256),
                        max(1 << 12, 2 * c["pairs"] + 1024) if self.base is not None else 0)
                old_caps = self.result.caps
                if caps[0] < old_caps[0] or (caps[2] and caps[2] < old_caps[2]):
                    self.result = None
                    release_workspaces()
                    a = self.args
                    self.result = chunk_ccs_device(self.pred, a["thresh"], a["dust_thresh"], connectivity=a["connectivity"],
                                                   dil_k=a["dil_k"], offset=a["offset"], base=self.base, sync=True,
                                                   max_runs=min(caps[0], old_caps[0]), max_segs=old_caps[1],
                                                   max_pairs=(min(caps[2], old_caps[2]) or None))
                self._tight = True
            self._eager(sync=False)                       # warm: occupancy queries, function attributes
            self.stream.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.syn_launch_count()
            with torch.cuda.graph(g, stream=self.stream):
                self._eager(sync=False)
            self.launches_per_run = int(L.syn_launch_count() - n0)
        self.graph = g

    def run(self):
        """One chunk: replays the captured graph on the pipeline's stream, ordered after the caller's current
        stream and before whatever the caller queues next."""
        torch = self.torch
        if self.graph is None:
            self._capture()
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == self.stream.cuda_stream:
            self.graph.replay()
        else:
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                self.graph.replay()
            cur.wait_stream(self.stream)
        return self.result

    def counts(self):
        """Synchronises and returns the counters of the last run; on a capacity overflow the buffers are grown,
        the graph is re-captured and the chunk is run again."""
        c = read_counts(self.result, stream=self.stream)
        if c["errflags"]:
            self.graph = None
            self._capture()              # the eager sync call inside grows the capacities for the current inputs
            c = read_counts(self.result, stream=self.stream)
        return c


def read_counts(result, stream=None):
    torch = _torch()
    counts = np.zeros(C.NCOUNTS, dtype=np.int64)
    C.check(C.lib().syn_read_counts(result.workspace.data_ptr(), counts.ctypes.data_as(ctypes.c_void_p),
                                    _stream_ptr(torch, stream)))
    result.counts = _counts_dict(counts)
    return result.counts


def count_overlaps_device(seg1, seg2, max_pairs=None, stream=None):
    """seg1: int32-typed (uint32 bits) cuda [X,Y,Z]; seg2: int64-typed (uint64 bits).  Returns ChunkResult-like."""
    torch = _torch()
    L = C.lib()
    nx, ny, nz = (int(s) for s in seg1.shape)
    device = seg1.device
    max_pairs = int(max_pairs or default_capacities((nx, ny, nz))[2])
    counts = np.zeros(C.NCOUNTS, dtype=np.int64)
    for _attempt in range(8):
        with torch.cuda.device(device):
            nbytes = L.syn_overlap_workspace_bytes(nx, ny, nz, max_pairs)
            if nbytes < 0:
                C.check(int(nbytes))
            ws = _workspace(torch, device, nbytes)
            prow = torch.empty(max_pairs, dtype=torch.int32, device=device)
            pcol = torch.empty(max_pairs, dtype=torch.int64, device=device)
            pval = torch.empty(max_pairs, dtype=torch.int64, device=device)
            rc = L.syn_count_overlaps(seg1.data_ptr(), seg2.data_ptr(), nx, ny, nz, prow.data_ptr(), pcol.data_ptr(),
                                      pval.data_ptr(), max_pairs, ws.data_ptr(), ws.numel(),
                                      counts.ctypes.data_as(ctypes.c_void_p), _stream_ptr(torch, stream))
        res = ChunkResult(seg1, None, prow, pcol, pval, _counts_dict(counts), ws)
        if rc == C.SYN_OK:
            return res
        if rc != C.SYN_ERR_CAPACITY:
            C.check(rc)
        ef = res.counts["errflags"]
        if ef & C.EF_TABLE:
            max_pairs *= 4
        if ef & C.EF_PAIRS:
            max_pairs = max(res.counts["pairs"], max_pairs + 1)
    raise C.CapacityError(C.SYN_ERR_CAPACITY, "pair capacity still too small after 8 attempts", counts)


def extract_faces_device(labels, stream=None):
    """labels: cuda int32 [X,Y,Z] -> list of six cuda int32 2-D tensors in Face.all_faces() order."""
    torch = _torch()
    nx, ny, nz = (int(s) for s in labels.shape)
    sizes = [ny * nz, ny * nz, nx * nz, nx * nz, nx * ny, nx * ny]
    shapes = [(ny, nz), (ny, nz), (nx, nz), (nx, nz), (nx, ny), (nx, ny)]
    buf = torch.empty(sum(sizes), dtype=torch.int32, device=labels.device)
    with torch.cuda.device(labels.device):
        C.check(C.lib().syn_extract_faces(labels.data_ptr(), nx, ny, nz, buf.data_ptr(), _stream_ptr(torch, stream)))
    out, o = [], 0
    for (n, sh) in zip(sizes, shapes):
        out.append(buf[o:o + n].view(sh))
        o += n
    return out, buf


def relabel_lut_device(labels, lut, stream=None):
    """In-place LUT relabel of a cuda int32 (uint32 bits) tensor; lut: cuda int32 tensor."""
    torch = _torch()
    with torch.cuda.device(labels.device):
        C.check(C.lib().syn_relabel_lut(labels.data_ptr(), labels.numel(), lut.data_ptr(), lut.numel(),
                                        _stream_ptr(torch, stream)))
    return labels


def max_u32_device(labels, stream=None):
    torch = _torch()
    out = torch.zeros(1, dtype=torch.int32, device=labels.device)
    with torch.cuda.device(labels.device):
        C.check(C.lib().syn_max_u32(labels.data_ptr(), labels.numel(), out.data_ptr(), _stream_ptr(torch, stream)))
    return int(out.cpu().numpy().view(np.uint32)[0])


def segment_stats_device(labels, max_id=None, stream=None):
    """cuda int32 (uint32 bits) [X,Y,Z] -> cuda int64 [max_id+1, 11] table (row i = id i)."""
    torch = _torch()
    nx, ny, nz = (int(s) for s in labels.shape)
    if max_id is None:
        max_id = max_u32_device(labels, stream)
    table = torch.empty((max_id + 1, C.SEG_NCOLS), dtype=torch.int64, device=labels.device)
    with torch.cuda.device(labels.device):
        C.check(C.lib().syn_segment_stats(labels.data_ptr(), nx, ny, nz, int(max_id), table.data_ptr(),
                                          _stream_ptr(torch, stream)))
    return table


def face_pair_edges_device(face_a, face_b, edges, n_edges, stream=None):
    torch = _torch()
    with torch.cuda.device(face_a.device):
        C.check(C.lib().syn_face_pair_edges(face_a.data_ptr(), face_b.data_ptr(), face_a.numel(), edges.data_ptr(),
                                            edges.shape[0], n_edges.data_ptr(), _stream_ptr(torch, stream)))


def union_min_map_device(edges, n_edges, n_ids, stream=None):
    torch = _torch()
    out = torch.empty(int(n_ids), dtype=torch.int32, device=edges.device)
    with torch.cuda.device(edges.device):
        C.check(C.lib().syn_union_min_map(edges.data_ptr(), n_edges.data_ptr(), edges.shape[0], int(n_ids),
                                          out.data_ptr(), _stream_ptr(torch, stream)))
    return out


def threshold_dilate_device(pred, thresh, dil_k, stream=None):
    """float32 cuda [X,Y,Z] -> cuda int32 0/1 volume: (pred > thresh) dilated by dil_k (2-D L1, axes 1,2)."""
    torch = _torch()
    nx, ny, nz = (int(s) for s in pred.shape)
    out = torch.empty((nx, ny, nz), dtype=torch.int32, device=pred.device)
    with torch.cuda.device(pred.device):
        C.check(C.lib().syn_threshold_dilate(pred.data_ptr(), nx, ny, nz, ctypes.c_float(float(thresh)), int(dil_k),
                                             out.data_ptr(), _stream_ptr(torch, stream)))
    return out


def unique_u64_device(vals, cap=1 << 16, stream=None):
    """Distinct values of an int64-typed (uint64 bits) cuda tensor, as a sorted numpy uint64 array."""
    torch = _torch()
    flat = vals.reshape(-1)
    while True:
        out = torch.empty(int(cap), dtype=torch.int64, device=flat.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=flat.device)
        with torch.cuda.device(flat.device):
            C.check(C.lib().syn_unique_u64(flat.data_ptr(), flat.numel(), out.data_ptr(), int(cap), cnt.data_ptr(),
                                           _stream_ptr(torch, stream)))
        n = int(cnt.cpu()[0])
        if n <= cap:
            return np.sort(out[:n].cpu().numpy().view(np.uint64))
        # overflow.  The count is only a lower bound of the need once the hash table filled up (every failed
        # insert adds cap + 1 to it), so it is trusted only when it is plausible; the need never exceeds the
        # number of elements.
        total = int(flat.numel())
        cap = min(total, max(cap * 4, n if n <= total else 0))
