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
import ctypes
import itertools
import os
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
    """[world, ...] gathered copy of `t`.  Which collective is used is decided by the backend alone, identically
    on every rank and before any communication (never by catching an error around a collective: ranks that
    disagree on the call would hang)."""
    if world == 1:
        return t.unsqueeze(0)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:
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


# ==================================================================================================
# The whole-volume job: chunk_ccs -> merge_ccs -> remap_ids -> chunk_overlaps with ONE exchange and
# ONE write of the label volume.
# ==================================================================================================
def grid_chunks(vol_shape, chunk_shape):
    """(grid, [(begin xyz, shape xyz)]) of the chunk grid in np.ndindex order (types/bbox/chunking.py:13-69:
    the last chunk along an axis is the remainder)."""
    from .types.bbox import chunk_bboxes
    boxes = chunk_bboxes(tuple(int(v) for v in vol_shape), tuple(int(c) for c in chunk_shape))
    grid = tuple(-(-int(v) // int(c)) for (v, c) in zip(vol_shape, chunk_shape))
    return grid, [(tuple(int(x) for x in b.min()), tuple(int(x) for x in b.shape())) for b in boxes]


def merge_plan(grid, world):
    """int32 plan of syn_merge_ccs_packed: {nchunks, npairs, slot of chunk c ..., (chunk here, hi face, chunk there,
    lo face) ...}.  Chunk i lives on rank i % world, in that rank's local slot i // world."""
    nchunks = int(np.prod(grid))
    per_rank = -(-nchunks // world)
    slots = [(i % world) * per_rank + i // world for i in range(nchunks)]
    pairs = interior_face_pairs(grid)
    plan = [nchunks, len(pairs)] + slots + [v for p in pairs for v in p]
    return np.asarray(plan, dtype=np.int32), per_rank, len(pairs)


class VolumeJob:
    """
    Device-resident segmentation of a chunked volume, chunks spread round robin over the ranks of `group`
    (one process per GPU).  One step =

      per own chunk : syn_chunk_ccs_begin   threshold, labelling, dust filter, segment table, face entries
      exchange      : ONE all_gather of the packed slots (tables + face entries, ~1-2 MB per chunk)
      every rank    : syn_merge_ccs_packed  global ids, face matching, union-find, merged table, size threshold
      per own chunk : syn_chunk_ccs_finish  foreground labels written ONCE, in final global ids; overlap counts

    (reference: proc/tasks.py:26-69 cc_task, :72-122 merge_ccs_task, :314-333 remap_ids_task, :290-297
    overlap_task on the remapped volume -- four task waves with files in between).  No host synchronisation
    inside a step; with graph=True the kernels of a step are replayed from two CUDA graphs around the all_gather.
    """

    def __init__(self, vol_shape, chunk_shape, thresh, dust_thresh, size_thr, connectivity=6, with_base=True,
                 rank=None, world=None, group=None, device=None, vol_offset=(0, 0, 0), graph=True, lanes=1,
                 share_sm=None):
        from . import _device as D
        from . import _cabi as C
        import torch.distributed as dist
        self.D, self.C = D, C
        # (a CPU device is accepted for the tests of the host-side plumbing: they replace the CUDA entry points)
        self.cuda = device is None or getattr(device, "type", "cuda") == "cuda"
        if self.cuda:
            self.torch = torch = D._torch()
        else:
            import torch
            self.torch = torch
        self.dist = dist
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if world > 1 else 0
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.grid, chunks = grid_chunks(vol_shape, chunk_shape)
        self.vol_shape = tuple(int(v) for v in vol_shape)
        self.chunk_begin = [tuple(b + o for (b, o) in zip(beg, vol_offset)) for (beg, _) in chunks]
        self.chunk_rel_begin = [beg for (beg, _) in chunks]
        self.chunk_shape = [shp for (_, shp) in chunks]
        self.nchunks = len(chunks)
        self.owners = chunk_owners(self.grid, self.world)
        self.mine = [i for i in range(self.nchunks) if self.owners[i] == self.rank]
        self.thresh, self.dust, self.size_thr = np.float32(thresh), int(dust_thresh), int(size_thr)
        self.conn, self.with_base, self.use_graph = int(connectivity), bool(with_base), bool(graph)
        # the exchange: "p2p" = own kernel storing the slots into the peers' symmetric memory over NVLink (multicast
        # through the NVSwitch when available) + a signal-pad barrier; "nccl" = one all_gather_into_tensor.
        # Decided identically on every rank before any communication (an environment variable, else the backend).
        self.exchange = "none"
        if self.world > 1:
            want = os.environ.get("SYN_EXCHANGE", "p2p" if self.cuda else "nccl")
            if want == "p2p" and self.cuda:
                # local, rank-independent facts only: every GPU of the box reaches every other one
                n = torch.cuda.device_count()
                ok = n >= self.world and all(torch.cuda.can_device_access_peer(a, b)
                                             for a in range(n) for b in range(n) if a != b)
                want = "p2p" if ok else "nccl"
            self.exchange = want if (want == "nccl" or self.cuda) else "nccl"
        plan, self.per_rank, self.npairs = merge_plan(self.grid, self.world)
        self.plan = torch.from_numpy(plan).to(self.device)
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None   # merged table, next to the labels
        self.lanes = int(lanes)
        # SYN_BEGIN_SHARE_SM (opt-in): k_stream with two resident CTAs per SM, for callers that run other chunks'
        # kernels next to a chunk pass (measured: job 255.9 -> 261.7 Gvoxel/s at N = 1, k_stream 338 -> 358 us)
        self.share_sm = bool(share_sm)
        self.lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)] \
            if self.cuda and self.lanes > 1 else []
        self.pred = [None] * len(self.mine)
        self.base = [None] * len(self.mine)
        self.labels = [torch.empty(self.chunk_shape[i], dtype=torch.int32, device=self.device) for i in self.mine]
        self.caps_runs = [D.default_capacities(self.chunk_shape[i])[0] for i in self.mine]
        self.max_pairs = [D.default_capacities(self.chunk_shape[i])[2] if with_base else 0 for i in self.mine]
        self.ws = [None] * len(self.mine)
        self.pairs = [None] * len(self.mine)
        self.cap_rows = self.cap_entries = None
        self.graphs = None
        self.launches_per_step = 0
        self.offsets = [(ctypes_i64x3(self.chunk_begin[i])) for i in self.mine]
        # faces with a neighbour chunk: only their voxels take part in the merge (merge_ccs.py:69-114)
        masks = [0] * self.nchunks
        for (ca, fa, cb, fb) in interior_face_pairs(self.grid):
            masks[ca] |= 1 << fa
            masks[cb] |= 1 << fb
        self.face_masks = [masks[i] for i in self.mine]

    # ---- buffers -----------------------------------------------------------------------------------
    @property
    def all_slots(self):
        """The gathered slots of all chunks that the current / last step works on."""
        return self._halves[self._half]

    def set_inputs(self, local_index, pred, base=None):
        """Resident inputs of own chunk #local_index (chunk self.mine[local_index]): float32 cuda tensor and,
        with_base, the int64-typed uint64 base ids.  The tensors must stay alive (and in place) while the job runs."""
        torch = self.torch
        shp = self.chunk_shape[self.mine[local_index]]
        if not (pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous() and tuple(pred.shape) == shp):
            raise ValueError(f"pred of chunk {self.mine[local_index]} must be a contiguous float32 cuda tensor {shp}")
        if self.with_base:
            if base is None or not (base.is_cuda and base.dtype == torch.int64 and base.is_contiguous()
                                    and tuple(base.shape) == shp):
                raise ValueError("base must be a contiguous int64-typed cuda tensor of the chunk's shape")
        self.pred[local_index], self.base[local_index] = pred, (base if self.with_base else None)
        self.graphs = None

    def _alloc(self, cap_rows, cap_entries):
        torch, L = self.torch, self.C.lib()
        # even capacities: every slot of the gathered buffer stays 16-byte aligned
        self.cap_rows, self.cap_entries = (int(cap_rows) + 1) & ~1, (int(cap_entries) + 1) & ~1
        self.slot_bytes = int(L.syn_packed_slot_bytes(self.cap_rows, self.cap_entries))
        nslots = self.per_rank * self.world
        self.my_slots = torch.zeros((self.per_rank, self.slot_bytes), dtype=torch.uint8, device=self.device)
        self._halves, self._half, self._symm = None, 0, None
        if self.world == 1:
            self._halves = [self.my_slots]
        elif self.exchange == "p2p" and not self._symm_alloc(nslots):
            self.exchange = "nccl"           # (agreed by all ranks inside _symm_alloc)
            self._halves = [torch.zeros((nslots, self.slot_bytes), dtype=torch.uint8, device=self.device)]
        elif self.exchange == "p2p":
            buf, hdl, half = self._symm
            self._halves = [buf[h * half:(h + 1) * half].view(nslots, self.slot_bytes) for h in (0, 1)]
            mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", False) and hdl.multicast_ptr else 0
            if os.environ.get("SYN_NO_MULTICAST"):
                mc = 0
            self._mc = mc
            self._peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
            self.torch.cuda.synchronize(self.device)
            hdl.barrier(channel=0, timeout_ms=30000)
        else:
            self._halves = [torch.zeros((nslots, self.slot_bytes), dtype=torch.uint8, device=self.device)]
        cap_total = self.nchunks * self.cap_rows
        self.max_edges = max(1024, self.nchunks * self.cap_entries // 2)
        nb = int(L.syn_merge_packed_workspace_bytes(self.nchunks, self.cap_rows, self.max_edges))
        if nb < 0:
            raise ValueError("chunk grid / capacities out of range for the packed merge")
        self.merge_ws = torch.empty(nb, dtype=torch.uint8, device=self.device)
        self.id_base = torch.zeros(self.nchunks + 1, dtype=torch.int64, device=self.device)
        self.merged = torch.empty((cap_total, self.C.SEG_NCOLS), dtype=torch.int64, device=self.device)
        self.fmap = torch.zeros(cap_total + 1, dtype=torch.int32, device=self.device)
        self.info = torch.zeros(8, dtype=torch.int64, device=self.device)
        self.graphs = None

    def _symm_alloc(self, nslots):
        """Symmetric (peer-mapped) exchange buffer, two halves used alternately: a rank one step ahead never
        overwrites what a peer is still merging (the per-step barrier keeps ranks within one step).  Returns False
        -- on EVERY rank, by an all_reduce of the outcome -- when any rank could not set it up."""
        torch = self.torch
        ok, res = 1, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            half = nslots * self.slot_bytes
            buf = symm_mem.empty(2 * half, dtype=torch.uint8, device=self.device)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
            buf.zero_()
            res = (buf, hdl, half)
        except Exception as e:      # noqa: BLE001 -- any failure means "use NCCL", decided collectively below
            print(f"synaptor_b200: symmetric-memory exchange unavailable on rank {self.rank} ({type(e).__name__}: {e}); "
                  "falling back to NCCL all_gather")
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.cpu()[0]) == 0:
            self._symm = None
            return False
        self._symm = res
        return True

    def _alloc_chunk(self, li):
        torch, L = self.torch, self.C.lib()
        nx, ny, nz = self.chunk_shape[self.mine[li]]
        nb = L.syn_ccs_workspace_bytes(nx, ny, nz, self.caps_runs[li], self.max_pairs[li])
        if nb < 0:
            self.C.check(int(nb))
        self.ws[li] = torch.empty(int(nb), dtype=torch.uint8, device=self.device)
        if self.with_base:
            mp = self.max_pairs[li]
            self.pairs[li] = (torch.empty(mp, dtype=torch.int32, device=self.device),
                              torch.empty(mp, dtype=torch.int64, device=self.device),
                              torch.empty(mp, dtype=torch.int64, device=self.device))
        self.graphs = None

    # ---- the three stages of a step (all asynchronous on the current stream) --------------------------
    def _begin(self, li, slot=None, cap_rows=None, cap_entries=None):
        """slot: uint8 tensor receiving the packed slot (default: this chunk's slot of the exchange buffer)."""
        L, C = self.C.lib(), self.C
        nx, ny, nz = self.chunk_shape[self.mine[li]]
        base = self.base[li]
        C.check(L.syn_chunk_ccs_begin(
            self.pred[li].data_ptr(), nx, ny, nz, ctypes_f32(self.thresh), self.conn, 0, self.dust,
            ctypes.cast(self.offsets[li], ctypes.c_void_p), self.labels[li].data_ptr(),
            base.data_ptr() if base is not None else None,
            self.max_pairs[li], (slot if slot is not None else self.my_slots[li]).data_ptr(),
            cap_rows or self.cap_rows, cap_entries or self.cap_entries,
            self.face_masks[li] | (0x100 if self.share_sm else 0), self.ws[li].data_ptr(),
            self.ws[li].numel(),
            self.caps_runs[li], self.D._stream_ptr(self.torch)))

    def _merge_ids(self):
        """The part of the merge the label writes wait for: id bases, edges, union-find, final id map."""
        L, C = self.C.lib(), self.C
        C.check(L.syn_merge_ccs_packed(
            self.all_slots.data_ptr(), self.slot_bytes, self.cap_rows, self.cap_entries, self.plan.data_ptr(),
            self.nchunks, self.npairs, self.size_thr, self.max_edges, self.merge_ws.data_ptr(), self.merge_ws.numel(),
            self.id_base.data_ptr(), None, self.fmap.data_ptr(), self.info.data_ptr(),
            self.D._stream_ptr(self.torch)))

    def _merge_table(self):
        """The merged segment table (on whatever stream is current: a side stream next to the label kernels)."""
        L, C = self.C.lib(), self.C
        C.check(L.syn_merge_ccs_table(
            self.all_slots.data_ptr(), self.slot_bytes, self.cap_rows, self.cap_entries, self.plan.data_ptr(),
            self.nchunks, self.size_thr, self.max_edges, self.merge_ws.data_ptr(), self.merge_ws.numel(),
            self.id_base.data_ptr(), self.merged.data_ptr(), self.info.data_ptr(), self.rank, self.world,
            self.D._stream_ptr(self.torch)))

    def _merge(self):
        self._merge_ids()
        self._merge_table()

    def _finish(self, li):
        L, C = self.C.lib(), self.C
        ci = self.mine[li]
        nx, ny, nz = self.chunk_shape[ci]
        base = self.base[li]
        pr = self.pairs[li] or (None, None, None)
        lut_base = self.id_base[ci:ci + 1]
        C.check(L.syn_chunk_ccs_finish(
            self.pred[li].data_ptr(), nx, ny, nz, self.conn, 0, self.labels[li].data_ptr(),
            base.data_ptr() if base is not None else None,
            pr[0].data_ptr() if pr[0] is not None else None, pr[1].data_ptr() if pr[1] is not None else None,
            pr[2].data_ptr() if pr[2] is not None else None, self.max_pairs[li], self.cap_rows,
            self.fmap.data_ptr(), lut_base.data_ptr(), self.fmap.numel(), self.ws[li].data_ptr(), self.ws[li].numel(),
            self.caps_runs[li], None, self.D._stream_ptr(self.torch)))

    def _gather(self):
        """The exchange of one step: afterwards self.all_slots holds the packed slots of every chunk."""
        if self.world == 1:
            return
        if self.exchange != "p2p":
            self.dist.all_gather_into_tensor(self._halves[0].view(-1), self.my_slots.view(-1), group=self.group)
            return
        buf, hdl, half = self._symm
        h = self._half = 1 - self._half
        off = h * half + self.rank * self.per_rank * self.slot_bytes
        peers = (ctypes.c_void_p * self.world)(*[p + off for p in self._peer_ptrs])
        self.C.check(self.C.lib().syn_push_slots(
            self.my_slots.data_ptr(), self.per_rank, self.slot_bytes, self.cap_rows, self.cap_entries,
            ctypes.cast(peers, ctypes.c_void_p), self.world, (self._mc + off) if self._mc else None,
            self.D._stream_ptr(self.torch)))
        hdl.barrier(channel=0, timeout_ms=30000)

    # ---- sizing --------------------------------------------------------------------------------------
    def prepare(self):
        """Sizes every buffer from one un-timed pass over the own chunks (collective: all ranks agree on the slot
        capacities), then captures the graphs.  Called by the first run()."""
        torch, D = self.torch, self.D
        for li in range(len(self.mine)):
            if self.pred[li] is None:
                raise RuntimeError(f"inputs of own chunk #{li} not set")
        need_rows, need_ent = 1, 1
        with self._on_stream():
            for li, ci in enumerate(self.mine):
                nx, ny, nz = self.chunk_shape[ci]
                for _attempt in range(6):
                    if self.ws[li] is None:
                        self._alloc_chunk(li)
                    big_rows = D.default_capacities((nx, ny, nz))[1]
                    big_ent = 2 * (ny * nz + nx * nz + nx * ny)
                    nb = int(self.C.lib().syn_packed_slot_bytes(big_rows, big_ent))
                    tmp = torch.zeros(nb, dtype=torch.uint8, device=self.device)
                    self._begin(li, tmp, big_rows, big_ent)
                    cnt = self._read_counts(li)
                    hdr = tmp[:128].view(torch.int64).cpu().numpy()
                    del tmp
                    if cnt["errflags"] & self.C.EF_RUNS:
                        self.caps_runs[li] = max(cnt["runs"], self.caps_runs[li] + 1)
                        self.ws[li] = None
                        continue
                    break
                else:
                    raise self.C.CapacityError(self.C.SYN_ERR_CAPACITY, "run capacity still too small", None)
                need_rows = max(need_rows, int(hdr[0]))
                need_ent = max(need_ent, int(hdr[1]))
                # keep the per-chunk workspaces tight: the run-indexed arrays dominate them
                tight = max(1 << 16, int(cnt["runs"] * 1.25) + 1024)
                if tight < self.caps_runs[li]:
                    self.caps_runs[li] = tight
                    self.ws[li] = None
                    self._alloc_chunk(li)
        need = torch.tensor([need_rows, need_ent], dtype=torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(need, op=self.dist.ReduceOp.MAX, group=self.group)
        need_rows, need_ent = (int(v) for v in need.cpu())
        if self.cap_rows is None or need_rows > self.cap_rows or need_ent > self.cap_entries:
            self._alloc(max(self.cap_rows or 0, int(need_rows * 1.25) + 64), max(self.cap_entries or 0,
                                                                                  int(need_ent * 1.25) + 256))
        if self.with_base and self.cuda:
            # one full step to learn the overlap-pair counts: the pair table (and the three kernels that walk it)
            # are sized for 2x the pairs seen instead of the worst-case default; an overflow later grows it again
            with self._on_stream():
                self._stage_a()
                self._gather()
                self._stage_b()
                self._sync()
                for li in range(len(self.mine)):
                    cnt = self._read_counts(li)
                    if cnt["errflags"] == 0:
                        tight = max(1 << 12, 2 * cnt["pairs"] + 1024)
                        if tight < self.max_pairs[li]:
                            self.max_pairs[li] = tight
                            self._alloc_chunk(li)
        self._capture()

    def _read_counts(self, li):
        counts = np.zeros(self.C.NCOUNTS, dtype=np.int64)
        import ctypes
        self.C.check(self.C.lib().syn_read_counts(self.ws[li].data_ptr(), counts.ctypes.data_as(ctypes.c_void_p),
                                                  self.D._stream_ptr(self.torch)))
        return self.D._counts_dict(counts)

    def _lanes(self):
        """Forks the lane streams off the current stream (chunks are independent until the merge: the chunk passes
        of neighbouring chunks go to alternating streams, so that one chunk's latency-bound labelling kernels run
        next to another chunk's HBM-bound streaming kernel)."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        for ls in self.lane_streams:
            ls.wait_stream(cur)
        return cur

    def _stage_a(self):
        n = len(self.mine)
        if not self.cuda or self.lanes <= 1 or n <= 1:
            for li in range(n):
                self._begin(li)
            return
        cur = self._lanes()
        for li in range(n):
            with self.torch.cuda.stream(self.lane_streams[li % self.lanes]):
                self._begin(li)
        for ls in self.lane_streams:
            cur.wait_stream(ls)

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()

    def _sync(self):
        if self.cuda:
            self.stream.synchronize()

    def _stage_b(self, after_ids=None):
        torch = self.torch
        self._merge_ids()
        if after_ids is not None:
            after_ids()
        if not self.cuda:
            self._merge_table()
            for li in range(len(self.mine)):
                self._finish(li)
            return
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)                 # fork: the table needs nothing but the id map
        with torch.cuda.stream(self.side):
            self._merge_table()
        n = len(self.mine)
        if self.lanes <= 1 or n <= 1:
            for li in range(n):
                self._finish(li)
        else:
            self._lanes()
            for li in range(n):
                with torch.cuda.stream(self.lane_streams[li % self.lanes]):
                    self._finish(li)
            for ls in self.lane_streams:
                cur.wait_stream(ls)
        cur.wait_stream(self.side)                 # join

    def _capture(self):
        torch, L = self.torch, self.C.lib()
        self.graphs = None
        with self._on_stream():
            n0 = L.syn_launch_count()
            self._stage_a()                      # warm: function attributes, occupancy queries
            self._gather()
            self._stage_b()
            self.launches_per_step = int(L.syn_launch_count() - n0)
            self._sync()
            if not self.use_graph or not self.cuda:
                return
            if self.world == 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._stage_a()
                    self._stage_b()
                self.graphs = (g,)
            else:
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=self.stream):
                    self._stage_a()
                gbs = []
                for h in range(len(self._halves)):       # stage B reads the half of the exchange buffer of its step
                    self._half = h
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb, stream=self.stream):
                        self._stage_b()
                    gbs.append(gb)
                self._half = 0
                self.graphs = (ga, gbs)

    # ---- running -------------------------------------------------------------------------------------
    def join(self):
        """Orders the caller's current stream after everything queued on the job's stream (see run(join=False))."""
        if self.cuda:
            self.torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def run(self, join=True):
        """Queues one step on the job's stream (ordered after the caller's current stream).  Asynchronous.
        join=False leaves the caller's stream un-ordered behind the step (call join() later): two jobs driven
        alternately then have two volumes in flight, and one volume's chunk passes overlap the other's label writes
        and merge."""
        torch = self.torch
        if self.cap_rows is None or (self.use_graph and self.cuda and self.graphs is None):
            self.prepare()
        if not self.cuda:
            self._stage_a()
            self._gather()
            self._stage_b()
            return
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self.graphs is None:
                self._stage_a()
                self._gather()
                self._stage_b()
            elif len(self.graphs) == 1:
                self.graphs[0].replay()
            else:
                self.graphs[0].replay()
                self._gather()
                self.graphs[1][self._half].replay()
        if join:
            cur.wait_stream(self.stream)

    def check(self):
        """Synchronises; True when the last step fitted every capacity.  Otherwise grows what overflowed
        (collective) and returns False: run() again."""
        torch = self.torch
        self._sync()
        info = self.info.cpu().numpy()
        ok = int(info[3]) == 0
        grow_pairs = False
        for li in range(len(self.mine)):
            cnt = self._read_counts(li)
            ef = cnt["errflags"]
            if ef & self.C.EF_RUNS:
                self.caps_runs[li] = max(cnt["runs"], self.caps_runs[li] + 1)
            if ef & self.C.EF_TABLE:
                self.max_pairs[li] *= 4
            if ef & self.C.EF_PAIRS:
                self.max_pairs[li] = max(cnt["pairs"], self.max_pairs[li] + 1)
            if ef & (self.C.EF_RUNS | self.C.EF_TABLE | self.C.EF_PAIRS):
                ok, grow_pairs = False, True
                self.ws[li] = None
        flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
        if int(flag.cpu()[0]) == 0:
            return True
        # rows / entries / edges: resize from the headers (prepare() re-measures them)
        if int(info[3]) & 2:
            self.cap_entries = None if self.cap_entries is None else self.cap_entries * 2
        self.cap_rows = None
        self.prepare()
        return False

    # ---- host arrays in, host results out ------------------------------------------------------------------
    def _host_setup(self, preds, bases):
        torch, D = self.torch, self.D
        n_own = len(self.mine)
        if len(preds) != n_own or (self.with_base and (bases is None or len(bases) != n_own)):
            raise ValueError(f"rank {self.rank} owns {n_own} chunks")
        if any(p is None for p in self.pred) or not getattr(self, "_own_inputs", False):
            for li, ci in enumerate(self.mine):
                shp = self.chunk_shape[ci]
                self.pred[li] = torch.empty(shp, dtype=torch.float32, device=self.device)
                self.base[li] = torch.empty(shp, dtype=torch.int64, device=self.device) if self.with_base else None
            self._own_inputs = True
            self.graphs = None
        if self.cap_rows is None:            # first call: blocking upload + sizing pass
            for li in range(n_own):
                D.upload_f32(torch, self.pred[li], preds[li])
                if self.with_base:
                    D.upload_u64(torch, self.base[li], bases[li])
            torch.cuda.synchronize(self.device)
            g = self.use_graph
            self.use_graph = False           # the host path launches per chunk as its inputs arrive
            self.prepare()
            self.use_graph = g
            self._tickets = [None, None]

    def _staging(self, slot):
        """Pinned host copies of the small outputs of one in-flight volume (two volumes can be in flight)."""
        torch = self.torch
        st = getattr(self, "_stage", None)
        if st is None or st.get("caps") != (self.cap_rows, self.cap_entries, tuple(self.max_pairs)):
            st = self._stage = {"caps": (self.cap_rows, self.cap_entries, tuple(self.max_pairs)), "slots": {}}
        if slot not in st["slots"]:
            def pin(t):
                return torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            d = {"info": pin(self.info), "id_base": pin(self.id_base), "merged": pin(self.merged), "fmap": pin(self.fmap),
                 "tables": pin(self.my_slots), "counts": [pin(w[:self.C.NCOUNTS * 8]) for w in self.ws],
                 "pairs": [tuple(pin(t) for t in pr) if pr is not None else None for pr in self.pairs]}
            st["slots"][slot] = d
        return st["slots"][slot]

    def _host_enqueue(self, preds, bases, prev=None):
        """Queues one volume: input copies (copy-in stream), chunk passes + exchange + merge + label writes (job
        stream), label volumes and the small outputs back to pinned host memory (copy-out stream).  Nothing here
        waits for the GPU.  `prev`: the ticket of the volume queued before this one, still in flight."""
        torch, D = self.torch, self.D
        n_own = len(self.mine)
        s_in = D.side_stream(torch, self.device, "in")
        s_out = D.side_stream(torch, self.device, "out")
        slot = 0 if prev is None else 1 - prev["slot"]
        stg = self._staging(slot)
        cur = torch.cuda.current_stream(self.device)
        s_in.wait_stream(cur)
        ev_in = []
        with torch.cuda.stream(s_in):
            for li in range(n_own):
                if prev is not None:
                    s_in.wait_event(prev["ev_fin"][li])      # the previous volume still reads this chunk's inputs
                D.upload_f32(torch, self.pred[li], preds[li])
                if self.with_base:
                    D.upload_u64(torch, self.base[li], bases[li])
                e = torch.cuda.Event()
                e.record(s_in)
                ev_in.append(e)
        hosts, ev_out, ev_fin = [], [], []
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if prev is not None:
                self.stream.wait_event(prev["ev_small"])      # its tables / counters / pairs are on the host now
            for li in range(n_own):
                self.stream.wait_event(ev_in[li])
                if prev is not None:
                    self.stream.wait_event(prev["ev_out"][li])   # its label chunk has left the device buffer
                self._begin(li)
            self._gather()
            self._merge_ids()
            self.side.wait_stream(self.stream)
            with torch.cuda.stream(self.side):
                self._merge_table()
            for li in range(n_own):
                self._finish(li)
                e = torch.cuda.Event()
                e.record(self.stream)
                ev_fin.append(e)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(e)
                    h = D.pinned_acquire(torch, self.labels[li].shape, self.labels[li].dtype)
                    h.copy_(self.labels[li], non_blocking=True)
                    e2 = torch.cuda.Event()
                    e2.record(s_out)
                hosts.append(h)
                ev_out.append(e2)
            self.stream.wait_stream(self.side)
            e_all = torch.cuda.Event()
            e_all.record(self.stream)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_all)
            stg["info"].copy_(self.info, non_blocking=True)
            stg["id_base"].copy_(self.id_base, non_blocking=True)
            stg["fmap"].copy_(self.fmap, non_blocking=True)
            stg["merged"].copy_(self.merged, non_blocking=True)
            stg["tables"].copy_(self.my_slots, non_blocking=True)
            for li in range(n_own):
                stg["counts"][li].copy_(self.ws[li][:self.C.NCOUNTS * 8], non_blocking=True)
                if self.pairs[li] is not None:
                    for (h, d) in zip(stg["pairs"][li], self.pairs[li]):
                        h.copy_(d, non_blocking=True)
            ev_small = torch.cuda.Event()
            ev_small.record(s_out)
        return {"slot": slot, "hosts": hosts, "ev_out": ev_out, "ev_fin": ev_fin, "ev_small": ev_small, "stg": stg}

    def _host_collect(self, t):
        """Waits for a queued volume and builds its host results; None when a capacity overflowed."""
        D, C = self.D, self.C
        t["ev_small"].synchronize()
        stg = t["stg"]
        info = stg["info"].numpy()
        bad = int(info[3]) != 0
        counts = []
        for li in range(len(self.mine)):
            cd = D._counts_dict(stg["counts"][li].numpy().view(np.int64).copy())
            counts.append(cd)
            bad = bad or (cd["errflags"] & (C.EF_RUNS | C.EF_TABLE | C.EF_PAIRS | C.EF_SEGS)) != 0
        if bad:
            for e in t["ev_out"]:
                e.synchronize()
            for h in t["hosts"]:
                D.pinned_release(h)
            return None
        n_merged = int(info[1])
        merged = stg["merged"].numpy()[:n_merged].copy()
        base_ids = stg["id_base"].numpy()
        fmap = stg["fmap"].numpy().view(np.uint32)
        slots = stg["tables"].numpy()
        tables, maps, ovs = [], [], []
        for li, ci in enumerate(self.mine):
            kept = int(base_ids[ci + 1] - base_ids[ci])
            tab = slots[li, 128:128 + self.cap_rows * 88].view(np.int64).reshape(self.cap_rows, 11)[:kept].copy()
            fm = fmap[int(base_ids[ci]) + 1: int(base_ids[ci]) + 1 + kept]
            tables.append(tab)
            maps.append(dict(zip(tab[:, 0].tolist(), fm.tolist())))
            if self.with_base:
                cd = counts[li]
                n = cd["pairs"]
                pr = stg["pairs"][li]
                shape = (0, 0) if cd["max_label"] == 0 or cd["max_base"] == 0 else (cd["max_label"] + 1, cd["max_base"] + 1)
                ovs.append((pr[0].numpy()[:n].view(np.uint32).astype(np.int64), pr[1].numpy()[:n].view(np.uint64).copy(),
                            pr[2].numpy()[:n].copy(), shape))
            else:
                ovs.append(None)
        labels = []
        for (h, e) in zip(t["hosts"], t["ev_out"]):
            e.synchronize()
            labels.append(D.pinned_numpy(h, np.uint32))
        return labels, merged, tables, maps, ovs

    def run_host(self, preds, bases=None):
        """
        One volume with HOST inputs (numpy arrays of the own chunks, C- or F-ordered, pinned or pageable): the input
        copies of the chunks run on a copy-in stream ahead of the chunk passes, the final label chunks come back on
        a copy-out stream while the remaining chunks are still being written.  Returns (list of uint32 label chunks
        in final ids (views of pooled pinned blocks), merged table numpy [n, 11], list of own chunk tables, list of
        {chunk-local id: final id}, list of (rows, cols, vals, shape) overlap triples).
        """
        self._host_setup(preds, bases)
        for _attempt in range(4):
            out = self._host_collect(self._host_enqueue(preds, bases))
            if out is not None:
                return out
            self.check()                     # grows what overflowed (collective)
        raise self.C.CapacityError(self.C.SYN_ERR_CAPACITY, "capacities never settled", None)

    def run_host_stream(self, volumes):
        """
        run_host over a sequence of volumes, software-pipelined: `volumes` yields (preds, bases) of the own chunks;
        the results come back in order.  The input copies of volume k + 1 are queued BEFORE volume k is collected,
        so the label volumes going back to the host, and the host-side construction of the result objects, run
        under the next volume's input copies (PCIe is full duplex): the link carries input copies back to back.
        Collective when world > 1: every rank must iterate the same number of volumes.
        """
        prev = prev_args = None
        for (preds, bases) in volumes:
            self._host_setup(preds, bases)
            cur = self._host_enqueue(preds, bases, prev=prev)
            if prev is not None:
                out = self._host_collect(prev)
                if out is None:                              # overflow: drain, grow, redo that volume on its own
                    self._host_collect(cur)
                    self.check()
                    yield self.run_host(*prev_args)
                    cur = self._host_enqueue(preds, bases)
                else:
                    yield out
            prev, prev_args = cur, (preds, bases)
        if prev is not None:
            out = self._host_collect(prev)
            if out is None:
                self.check()
                out = self.run_host(*prev_args)
            yield out

    # ---- results (each synchronises) -------------------------------------------------------------------
    def counts(self, local_index):
        self._sync()
        return self._read_counts(local_index)

    def merged_table(self, all_ranks=False):
        """[n, 11] int64 numpy: id, size, centroid xyz, bbox begin xyz, bbox end xyz (ascending id).  With several
        ranks every rank produces the rows of the merged ids that belong to ITS chunks; all_ranks=True gathers the
        shares (collective) into the whole table."""
        self._sync()
        n = int(self.info.cpu()[1])
        part = self.merged[:n].cpu().numpy()
        if not all_ranks or self.world == 1:
            return part
        parts = [None] * self.world
        self.dist.all_gather_object(parts, part, group=self.group)
        full = np.concatenate(parts)
        return full[np.argsort(full[:, 0], kind="stable")]

    def exchange_info(self):
        """How the packed slots travel between the ranks."""
        if self.world == 1:
            return "none (one rank)"
        if self.exchange != "p2p":
            return "NCCL all_gather_into_tensor of the fixed-size slots"
        return ("own kernel (syn_push_slots): valid part of the slots stored into the peers' symmetric memory over NVLink, "
                + ("one multimem.st per 16 B to the NVSwitch multicast address" if self._mc else "one store per peer")
                + ", then the symmetric-memory signal-pad barrier")

    def info_dict(self):
        self._sync()
        i = self.info.cpu().numpy()
        return {"total_ids": int(i[0]), "n_merged": int(i[1]), "n_edges": int(i[2]), "flags": int(i[3])}

    def chunk_table(self, chunk_index):
        """Segment table of ANY chunk of the grid (every rank holds all of them after the exchange)."""
        self._sync()
        slot = int(self.plan[2 + chunk_index].cpu())
        raw = self.all_slots[slot]
        kept = int(raw[:8].view(self.torch.int64).cpu()[0])
        t = raw[128:128 + self.cap_rows * 88].view(self.torch.int64).view(self.cap_rows, 11)
        return t[:kept].cpu().numpy()

    def chunk_id_map(self, chunk_index):
        """{chunk-local id: final global id} of a chunk (the reference's chunk_id_maps entry, tasks.py:72-122)."""
        t = self.chunk_table(chunk_index)
        base = int(self.id_base.cpu()[chunk_index])
        fm = self.fmap[base + 1: base + 1 + len(t)].cpu().numpy().view(np.uint32)
        return dict(zip(t[:, 0].tolist(), fm.tolist()))

    def overlaps(self, local_index):
        """(rows, cols, vals, shape) of own chunk #local_index: overlap_task on the remapped chunk."""
        cnt = self.counts(local_index)
        n = cnt["pairs"]
        pr = self.pairs[local_index]
        rows = pr[0][:n].cpu().numpy().view(np.uint32).astype(np.int64)
        cols = pr[1][:n].cpu().numpy().view(np.uint64)
        vals = pr[2][:n].cpu().numpy()
        shape = (0, 0) if cnt["max_label"] == 0 or cnt["max_base"] == 0 else (cnt["max_label"] + 1, cnt["max_base"] + 1)
        return rows, cols, vals, shape


def ctypes_i64x3(v):
    import ctypes
    return (ctypes.c_int64 * 3)(int(v[0]), int(v[1]), int(v[2]))


def ctypes_f32(v):
    import ctypes
    return ctypes.c_float(float(v))
