"""
Cross-chunk merge with the chunks resident on one or several GPUs
(reference: synaptor/proc/tasks.py:72-122 merge_ccs_task + :314-333 remap_ids_task,
distributed by the reference as one task per chunk plus one global merge task).

One process per GPU; every rank owns the chunks `owner[i] == rank` of the chunk grid
(np.ndindex order, round robin).  The per-chunk pass needs no communication.  The
merge has ONE real exchange step, done with torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests):

  all_gather  kept-row counts            -> global id base of every chunk (serial ids, assign_ids.py:8-42)
  all_gather  the six faces of every chunk, already in global ids
  all_gather  the per-chunk segment tables

after which every rank runs the same tiny deterministic device computation (face-pair
edges, union-find to the minimum id, merged table) and relabels its own chunks through a
LUT.  Nothing here computes on the host except shape bookkeeping.
"""
import itertools
from dataclasses import dataclass

import numpy as np

FACE_ORDER = [(axis, hi) for axis in range(3) for hi in (True, False)]   # continuation.py:57-59


def chunk_owners(grid, world):
    """owner rank of every chunk of `grid`, np.ndindex order, round robin."""
    n = int(np.prod(grid))
    return [i % world for i in range(n)]


def face_layout(shape):
    """(offsets, sizes, 2-D shapes) of the six faces inside one chunk's face buffer."""
    nx, ny, nz = (int(s) for s in shape)
    shapes = [(ny, nz), (ny, nz), (nx, nz), (nx, nz), (nx, ny), (nx, ny)]
    sizes = [a * b for (a, b) in shapes]
    offs = [0] + list(itertools.accumulate(sizes))[:-1]
    return offs, sizes, shapes


def interior_face_pairs(grid):
    """
    [(chunk index here, face slot here, chunk index there, face slot there)] for every
    interior face of the chunk grid: the hi face of a chunk against the lo face of its
    +1 neighbour along each axis (merge_ccs.py:69-114 visits exactly these pairs).
    """
    grid = tuple(int(g) for g in grid)
    lin = {idx: i for (i, idx) in enumerate(np.ndindex(grid))}
    pairs = []
    for idx in np.ndindex(grid):
        for axis in range(3):
            if idx[axis] == grid[axis] - 1:
                continue
            other = list(idx)
            other[axis] += 1
            pairs.append((lin[idx], 2 * axis, lin[tuple(other)], 2 * axis + 1))
    return pairs


def id_bases(kept_counts):
    """Exclusive prefix of the per-chunk row counts: chunk i's rows get ids base[i]+1 ... (assign_ids.py:8-25)."""
    c = np.asarray(kept_counts, dtype=np.int64)
    return np.concatenate(([0], np.cumsum(c)[:-1])).astype(np.int64), int(c.sum())


class DeviceOps:
    """The CUDA kernels behind the merge (libsynaptor_b200.so).  Tests swap in a CPU stand-in."""

    def __init__(self):
        from . import _device as D
        self.D = D
        self.torch = D._torch()

    def extract_faces(self, labels):
        return self.D.extract_faces_device(labels)[1]

    def read_counts(self, result):
        return self.D.read_counts(result)

    def ids_to_lut(self, table, n, base, lut_len, device):
        import ctypes
        from . import _cabi as C
        torch = self.torch
        lut = torch.zeros(int(lut_len), dtype=torch.int32, device=device)
        if n:
            with torch.cuda.device(device):
                C.check(C.lib().syn_ids_to_lut(table.data_ptr(), int(n), C.SEG_NCOLS, ctypes.c_uint32(int(base)),
                                               lut.data_ptr(), int(lut_len), self.D._stream_ptr(torch)))
        return lut

    def relabel(self, arr, lut):
        return self.D.relabel_lut_device(arr, lut)

    def face_pair_edges(self, fa, fb, edges, n_edges):
        self.D.face_pair_edges_device(fa, fb, edges, n_edges)

    def union_min_map(self, edges, n_edges, n_ids):
        return self.D.union_min_map_device(edges, n_edges, n_ids)

    def merge_grid(self, faces_all, tables_all, kept, slots, grid, shape, maxk, size_thr, total, max_edges):
        """Everything between the exchanges and the LUT composition in one C call (syn_merge_ccs_grid)."""
        import ctypes
        from . import _cabi as C
        torch = self.torch
        device = faces_all.device
        edges = torch.empty((max_edges, 2), dtype=torch.int32, device=device)
        n_edges = torch.empty(1, dtype=torch.int64, device=device)
        gmap = torch.empty(total + 1, dtype=torch.int32, device=device)
        rows = torch.empty((max(total, 1), C.SEG_NCOLS), dtype=torch.int64, device=device)
        merged = torch.empty((max(total, 1), C.SEG_NCOLS), dtype=torch.int64, device=device)
        fmap = torch.empty(total + 1, dtype=torch.int32, device=device)
        n_merged = torch.zeros(1, dtype=torch.int64, device=device)
        kept_h = np.ascontiguousarray(kept, dtype=np.int64)
        slots_h = np.ascontiguousarray(slots, dtype=np.int64)
        grid_h = np.asarray(grid, dtype=np.int64)
        shape_h = np.asarray(shape, dtype=np.int64)
        vp = ctypes.c_void_p
        with torch.cuda.device(device):
            C.check(C.lib().syn_merge_ccs_grid(
                faces_all.data_ptr(), tables_all.data_ptr(), kept_h.ctypes.data_as(vp), slots_h.ctypes.data_as(vp),
                grid_h.ctypes.data_as(vp), shape_h.ctypes.data_as(vp), int(maxk), int(size_thr), edges.data_ptr(),
                int(max_edges), n_edges.data_ptr(), gmap.data_ptr(), rows.data_ptr(), merged.data_ptr(),
                fmap.data_ptr(), n_merged.data_ptr(), self.D._stream_ptr(torch)))
        return merged, fmap, n_merged, n_edges

    def merge_overlaps(self, rows, cols, counts, n_ids):
        import ctypes
        from . import _cabi as C
        torch = self.torch
        argcol = torch.empty(max(int(n_ids), 1), dtype=torch.int64, device=rows.device)
        maxval = torch.empty(max(int(n_ids), 1), dtype=torch.int64, device=rows.device)
        with torch.cuda.device(rows.device):
            C.check(C.lib().syn_merge_overlaps(rows.data_ptr(), cols.data_ptr(), counts.data_ptr(), int(rows.numel()),
                                               int(n_ids), argcol.data_ptr(), maxval.data_ptr(),
                                               self.D._stream_ptr(torch)))
        return argcol, maxval

    def merge_tables(self, rows, n_rows, gmap, size_thr):
        import ctypes
        from . import _cabi as C
        torch = self.torch
        merged = torch.empty((max(int(n_rows), 1), C.SEG_NCOLS), dtype=torch.int64, device=gmap.device)
        fmap = torch.empty(int(n_rows) + 1, dtype=torch.int32, device=gmap.device)
        n_merged = torch.zeros(1, dtype=torch.int64, device=gmap.device)
        with torch.cuda.device(gmap.device):
            C.check(C.lib().syn_merge_seg_tables(rows.data_ptr() if n_rows else None, int(n_rows), gmap.data_ptr(),
                                                 int(size_thr), merged.data_ptr(), fmap.data_ptr(),
                                                 n_merged.data_ptr(), self.D._stream_ptr(torch)))
        return merged, fmap, n_merged


@dataclass
class MergeResult:
    merged_table: object      # [n_merged, 11] int64 (id, size, centroid, bbox), ascending id; device tensor
    n_merged: int
    id_base: np.ndarray       # per chunk of the grid
    fmap: object              # [total ids + 1] final id map (0 = dropped), device tensor
    chunk_luts: list          # per LOCAL chunk: chunk-local id -> final id (device tensors)
    n_edges: int


def _all_gather(torch, dist, t, world, group):
    if world == 1:
        return t.unsqueeze(0)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):   # backends without the flat variant
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def merge_ccs_device(results, grid, rank=None, world=None, size_thr=0, group=None, ops=None, remap=True):
    """
    results : ChunkResult (or list of them) of THIS rank's chunks, in increasing chunk index; each
              must carry `counts` (kept, max_label).  All chunks must have the same shape.
    grid    : chunk grid shape, e.g. (2, 2, 2); chunk i (np.ndindex order) lives on rank i % world.
    Relabels every local label volume in place to the merged global ids (remap=True) and returns a
    MergeResult.  Collective: every rank of `group` must call it.
    """
    import torch
    import torch.distributed as dist
    ops = ops or DeviceOps()
    if not isinstance(results, (list, tuple)):
        results = [results]
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if world > 1 else 0
    grid = tuple(int(g) for g in grid)
    nchunks = int(np.prod(grid))
    owners = chunk_owners(grid, world)
    mine = [i for i in range(nchunks) if owners[i] == rank]
    if len(mine) != len(results):
        raise ValueError(f"rank {rank} owns chunks {mine} but was given {len(results)} results")
    per_rank = -(-nchunks // world)
    if nchunks % world:
        raise ValueError("the chunk grid must divide evenly over the ranks")
    for r in results:
        if not r.counts:
            ops.read_counts(r)
    device = results[0].labels.device
    shape = tuple(int(s) for s in results[0].labels.shape)
    offs, sizes, shapes2d = face_layout(shape)
    facebuf = sum(sizes)

    # ---- exchange 1: row counts -> global id bases ------------------------------------------------
    kept_local = torch.tensor([[r.counts["kept"], r.counts["max_label"]] for r in results], dtype=torch.int64,
                              device=device)
    kept_all = _all_gather(torch, dist, kept_local, world, group).cpu().numpy()       # [world, per_rank, 2]
    kept = np.zeros(nchunks, dtype=np.int64)
    for i in range(nchunks):
        kept[i] = kept_all[owners[i], i // world, 0]
    base, total = id_bases(kept)
    maxk = int(kept.max()) if nchunks else 0

    # ---- local: chunk-local id -> global id LUT, faces in global ids -------------------------------
    luts, faces_local, tables_local = [], [], []
    for (ci, r) in zip(mine, results):
        n = int(kept[ci])
        lut = ops.ids_to_lut(r.seg_table, n, int(base[ci]), r.counts["max_label"] + 1, device)
        luts.append(lut)
        fb = ops.extract_faces(r.labels)
        ops.relabel(fb, lut)
        faces_local.append(fb)
        t = torch.zeros((max(maxk, 1), r.seg_table.shape[1]), dtype=torch.int64, device=device)
        if n:
            t[:n] = r.seg_table[:n]
        tables_local.append(t)

    # ---- exchange 2 + 3: faces and tables -----------------------------------------------------------
    faces_all = _all_gather(torch, dist, torch.stack(faces_local), world, group)     # [world, per_rank, facebuf]
    tables_all = _all_gather(torch, dist, torch.stack(tables_local), world, group)   # [world, per_rank, maxk, 11]

    pairs = interior_face_pairs(grid)
    cap = max(1024, sum(sizes[sa] for (_, sa, _, _) in pairs))
    if hasattr(ops, "merge_grid"):
        # ---- every rank: one C call -- edges over the interior faces, union-find to the smallest id, rows in
        # global id order, merged table + size threshold (identical inputs, deterministic => identical results)
        slots = [owners[i] * per_rank + i // world for i in range(nchunks)]
        merged, fmap, n_merged_dev, n_edges = ops.merge_grid(faces_all.contiguous(), tables_all.contiguous(), kept, slots,
                                                             grid, shape, max(maxk, 1), size_thr, total, cap)
    else:
        def chunk_faces(i):
            return faces_all[owners[i], i // world]

        # ---- every rank: edges over the interior faces, union-find to the smallest id ---------------------
        edges = torch.empty((cap, 2), dtype=torch.int32, device=device)
        n_edges = torch.zeros(1, dtype=torch.int64, device=device)
        for (ia, sa, ib, sb) in pairs:
            fa = chunk_faces(ia)[offs[sa]:offs[sa] + sizes[sa]]
            fb = chunk_faces(ib)[offs[sb]:offs[sb] + sizes[sb]]
            ops.face_pair_edges(fa, fb, edges, n_edges)
        gmap = ops.union_min_map(edges, n_edges, total + 1)

        # ---- merged table + size threshold ------------------------------------------------------------
        rows = torch.cat([tables_all[owners[i], i // world, :int(kept[i])] for i in range(nchunks)]) if total else \
            torch.zeros((0, results[0].seg_table.shape[1]), dtype=torch.int64, device=device)
        merged, fmap, n_merged_dev = ops.merge_tables(rows, total, gmap, size_thr)

    # ---- local: compose chunk LUT with the final map, relabel the volumes in place ---------------------
    final_luts = []
    for (r, lut) in zip(results, luts):
        ops.relabel(lut, fmap)        # lut[c] = fmap[global id of c]
        final_luts.append(lut)
        if remap:
            ops.relabel(r.labels, lut)
    n_merged = int(n_merged_dev.cpu()[0])
    return MergeResult(merged[:n_merged], n_merged, base, fmap, final_luts, int(n_edges.cpu()[0]))


def merge_overlaps_device(results, merge, rank=None, world=None, group=None, ops=None):
    """
    merge_overlaps_task (proc/tasks.py:300-311) for chunks resident on one or several GPUs, after
    merge_ccs_device: `results` are THIS rank's ChunkResults of the fused pass (pair_rows / pair_cols / pair_vals
    = the chunk's overlap triples in chunk-local labels, counts["pairs"]), `merge` the MergeResult (chunk_luts:
    chunk-local label -> merged global id, 0 = dropped).  One all_gather of the (padded) triples, then every rank
    sums duplicates and takes the arg-max base id per merged cleft on the device (syn_merge_overlaps).
    Returns {merged cleft id: (base id, voxel count)} as three numpy arrays (ids ascending, base ids, counts).
    Collective: every rank of `group` must call it.
    """
    import torch
    import torch.distributed as dist
    ops = ops or DeviceOps()
    if not isinstance(results, (list, tuple)):
        results = [results]
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if world > 1 else 0
    device = results[0].labels.device
    rows_l, cols_l, vals_l = [], [], []
    for (r, lut) in zip(results, merge.chunk_luts):
        n = int(r.counts["pairs"])
        rows = r.pair_rows[:n].clone()
        ops.relabel(rows, lut)                       # chunk-local label -> merged global id (0 = dropped)
        rows_l.append(rows)
        cols_l.append(r.pair_cols[:n])
        vals_l.append(r.pair_vals[:n])
    rows = torch.cat(rows_l) if rows_l else torch.zeros(0, dtype=torch.int32, device=device)
    cols = torch.cat(cols_l) if cols_l else torch.zeros(0, dtype=torch.int64, device=device)
    vals = torch.cat(vals_l) if vals_l else torch.zeros(0, dtype=torch.int64, device=device)
    if world > 1:
        n_local = torch.tensor([rows.numel()], dtype=torch.int64, device=device)
        n_all = _all_gather(torch, dist, n_local, world, group).cpu().numpy().reshape(-1)
        cap = int(n_all.max()) if len(n_all) else 0
        packed = torch.zeros((max(cap, 1), 3), dtype=torch.int64, device=device)    # zero rows are ignored
        if rows.numel():
            packed[:rows.numel(), 0] = rows.to(torch.int64) & 0xFFFFFFFF
            packed[:rows.numel(), 1] = cols
            packed[:rows.numel(), 2] = vals
        allp = _all_gather(torch, dist, packed, world, group).reshape(-1, 3)
        rows = allp[:, 0].to(torch.int32).contiguous()
        cols = allp[:, 1].contiguous()
        vals = allp[:, 2].contiguous()
    n_ids = int(merge.fmap.numel())
    argcol, maxval = ops.merge_overlaps(rows, cols, vals, n_ids)
    mv = maxval.cpu().numpy()[:n_ids]
    ids = np.nonzero(mv > 0)[0]
    return ids.astype(np.int64), argcol.cpu().numpy().view(np.uint64)[ids], mv[ids]
