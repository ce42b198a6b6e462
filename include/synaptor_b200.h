/*
 * synaptor_b200.h -- C ABI of libsynaptor_b200.so (sm_100a).
 *
 * Drop-in boundary for Synaptor's chunk_ccs / chunk_overlaps / merge_ccs /
 * remap_ids hot path.  Every entry point replaces one native or Python-level
 * call of the reference (cited per function, paths relative to the reference
 * tree).  Conventions:
 *
 *   - all volume pointers are CALLER-OWNED DEVICE pointers (e.g.
 *     torch.Tensor.data_ptr()), C-contiguous [x][y][z], z fastest -- the layout
 *     the reference has after np.ascontiguousarray (proc/seg/conncomps.py:21);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success or a negative SYN_ERR_* code and never
 *     throws; syn_last_error() gives the message for the calling thread;
 *   - variable-length results use caller-provided capacity; the true sizes come
 *     back in the `counts` array, and SYN_ERR_CAPACITY says "retry larger";
 *   - no hidden host synchronisation unless `counts_host` is non-NULL (then the
 *     call ends with one D2H copy of the counts + a stream synchronise).
 */
#ifndef SYNAPTOR_B200_H
#define SYNAPTOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SYN_API __attribute__((visibility("default")))
#else
#define SYN_API
#endif

#define SYN_OK 0
#define SYN_ERR_INVALID (-1)   /* bad argument (shape, connectivity, null pointer, ...) */
#define SYN_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define SYN_ERR_CAPACITY (-3)  /* max_runs / max_segs / max_pairs too small: see counts */
#define SYN_ERR_WORKSPACE (-4) /* workspace_bytes smaller than syn_ccs_workspace_bytes() */

/* indices into the int64 counts[SYN_NCOUNTS] array (device copy lives at the
 * start of the workspace; host copy is filled when counts_host != NULL) */
enum {
    SYN_CNT_RUNS = 0,        /* z-runs of the labelled mask (needed max_runs) */
    SYN_CNT_COMPONENTS = 1,  /* raw connected components before the dust filter */
    SYN_CNT_KEPT = 2,        /* rows of the segment table (needed max_segs) */
    SYN_CNT_PAIRS = 3,       /* distinct (label, base id) overlap pairs (needed max_pairs) */
    SYN_CNT_MAX_BASE = 4,    /* max over the whole base volume (bit pattern of a uint64) */
    SYN_CNT_MAX_LABEL = 5,   /* largest label written to the label volume */
    SYN_CNT_ERRFLAGS = 6,    /* bit 0 runs overflow, bit 1 segs overflow, bit 2 pair table overflow,
                                bit 3 pair output overflow */
    SYN_CNT_FG_VOXELS = 7,   /* voxels above threshold */
    SYN_CNT_TABLE_SLOTS = 8, /* occupied slots in the pair hash table */
    SYN_CNT_ACTIVE_WORDS = 9, /* 32-voxel mask words holding at least one labelled voxel */
    SYN_CNT_TILE_RECORDS = 10, /* per-tile partial statistics records written by the tile pass */
    SYN_CNT_FLAGGED_TILES = 11, /* tiles left to the global kernels (too dense for shared memory) */
    SYN_CNT_BND_WORDS = 12,  /* non-empty mask words on a tile border (input of the global union pass) */
    SYN_CNT_TICKET = 13,     /* internal: CTAs of the streaming kernel that have finished */
    SYN_CNT_DBG0 = 16,       /* 16..31: phase timers of instrumented builds (-DSYN_PHASE_TIMING), else 0 */
    SYN_NCOUNTS = 32
};

/* columns of one segment-table row (int64 each), in the order of the reference's
 * DataFrame (proc/seg/misc.py:14-37, proc/colnames.py:6-24) preceded by the id */
enum {
    SYN_SEG_ID = 0, SYN_SEG_SIZE = 1, SYN_SEG_CX = 2, SYN_SEG_CY = 3, SYN_SEG_CZ = 4,
    SYN_SEG_BX = 5, SYN_SEG_BY = 6, SYN_SEG_BZ = 7, SYN_SEG_EX = 8, SYN_SEG_EY = 9, SYN_SEG_EZ = 10,
    SYN_SEG_NCOLS = 11
};

SYN_API int syn_version(void);
SYN_API const char *syn_last_error(void);

/* Bytes of device workspace syn_chunk_ccs needs for this shape.
 * max_runs  : capacity for z-runs (worst case nx*ny*ceil(nz/2));
 * max_pairs : capacity of the overlap pair table (0 when base_seg is NULL). */
SYN_API int64_t syn_ccs_workspace_bytes(int64_t nx, int64_t ny, int64_t nz, int64_t max_runs, int64_t max_pairs);

/*
 * syn_chunk_ccs -- the whole of proc.tasks.cc_task's array work
 * (synaptor/proc/tasks.py:26-69) and, when base_seg != NULL, of
 * proc.tasks.overlap_task (tasks.py:290-297) fused into the same pass:
 *
 *   mask = pred > thresh                         proc/seg/conncomps.py:17
 *   [dil_k > 0: 2-D L1 dilation over axes 1,2]   conncomps.py:32-52 -> seg_utils/_seg_utils.cpp:25-82,180-200
 *   3-D connected components, labels 1..N in     conncomps.py:22 -> cc3d.connected_components(connectivity=6)
 *     raster order of first voxel
 *   ids touching any chunk face are exempt       proc/seg/continuation.py:87-114, tasks.py:45-51
 *   ids with size < dust_thresh removed (-> 0)   seg_utils/filtering.py:7-40 -> _relabel.cpp:19-38
 *   size / centroid / bbox per surviving id      seg_utils/describe.py:14-80, _describe.cpp:23-64,
 *                                                scipy.ndimage.find_objects (half-open boxes)
 *   (label, base id) pair counts, first-         seg_utils/overlap.py:10-61 (collections.Counter order)
 *     occurrence raster order, + max(base)
 *
 * connectivity: 6 (the reference), 18 or 26 (extensions).
 * labels_out  : uint32 [nx][ny][nz].
 * seg_table   : int64 [max_segs][SYN_SEG_NCOLS], rows in ascending id; centroid
 *               and bbox already shifted by offset_xyz (describe.py:28,55-59).
 * pair_*      : device arrays of capacity max_pairs (ignored when base_seg == NULL).
 * counts_host : optional int64[SYN_NCOUNTS]; when given the call synchronises and
 *               returns SYN_ERR_CAPACITY if any capacity was exceeded.
 */
SYN_API int syn_chunk_ccs(const float *pred, int64_t nx, int64_t ny, int64_t nz,
                  float thresh, int connectivity, int dil_k, int64_t dust_thresh,
                  const int64_t *offset_xyz /* host int64[3] or NULL */,
                  uint32_t *labels_out,
                  int64_t *seg_table, int64_t max_segs,
                  const uint64_t *base_seg,
                  uint32_t *pair_rows, uint64_t *pair_cols, int64_t *pair_counts, int64_t max_pairs,
                  void *workspace, int64_t workspace_bytes, int64_t max_runs,
                  int64_t *counts_host, void *stream);

/* Copies the device counts at the head of `workspace` to host and synchronises. */
SYN_API int syn_read_counts(const void *workspace, int64_t *counts_host, void *stream);

/*
 * syn_count_overlaps -- seg_utils.count_overlaps(seg1, seg2, orig_ids=True)
 * (synaptor/seg_utils/overlap.py:10-61) on arbitrary uint32 / uint64 volumes.
 * Fills pair_* in first-occurrence raster order; counts[SYN_CNT_PAIRS],
 * counts[SYN_CNT_MAX_LABEL] = max(seg1), counts[SYN_CNT_MAX_BASE] = max(seg2).
 * Workspace: syn_overlap_workspace_bytes(n_voxels, max_pairs).
 */
SYN_API int64_t syn_overlap_workspace_bytes(int64_t nx, int64_t ny, int64_t nz, int64_t max_pairs);
SYN_API int syn_count_overlaps(const uint32_t *seg1, const uint64_t *seg2, int64_t nx, int64_t ny, int64_t nz,
                       uint32_t *pair_rows, uint64_t *pair_cols, int64_t *pair_counts, int64_t max_pairs,
                       void *workspace, int64_t workspace_bytes,
                       int64_t *counts_host, void *stream);

/*
 * syn_extract_faces -- the six np.take(seg, -1|0, axis) slices of
 * proc/seg/continuation.py:52-55, in Face.all_faces() order
 * (axis0 hi, axis0 lo, axis1 hi, axis1 lo, axis2 hi, axis2 lo), written back to
 * back into faces_out (uint32, 2*(ny*nz + nx*nz + nx*ny) elements).
 */
SYN_API int syn_extract_faces(const uint32_t *labels, int64_t nx, int64_t ny, int64_t nz,
                      uint32_t *faces_out, void *stream);

/*
 * syn_transpose_zyx -- the np.ascontiguousarray of proc/seg/conncomps.py:21 for the F-ordered [x,y,z] chunk that
 * CloudVolume hands the tasks (io/backends/cloudvolume.py:6-20), on the device: src_zyx holds the array's bytes as
 * they are (= C-contiguous [nz][ny][nx]), dst_xyz receives C-contiguous [nx][ny][nz].  elem_bytes 4 or 8.
 */
SYN_API int syn_transpose_zyx(const void *src_zyx, void *dst_xyz, int64_t nx, int64_t ny, int64_t nz, int elem_bytes,
                              void *stream);

/*
 * syn_relabel_lut -- seg_utils.relabel_data (seg_utils/relabel.py:7-27 ->
 * _relabel.cpp:19-38) with the dict given as a dense table: values v < lut_len
 * become lut[v] (identity entries for non-keys), others stay.  In place.
 */
SYN_API int syn_relabel_lut(uint32_t *labels, int64_t n, const uint32_t *lut, int64_t lut_len, void *stream);

/*
 * syn_segment_stats -- seg_utils.segment_sizes / centers_of_mass /
 * bounding_boxes (seg_utils/describe.py:14-80, _describe.cpp:23-64) of a label
 * volume whose ids are <= max_id.  stats: int64 [max_id+1][SYN_SEG_NCOLS]
 * (row i = id i; size 0 for absent ids; no offset applied).
 */
SYN_API int syn_segment_stats(const uint32_t *labels, int64_t nx, int64_t ny, int64_t nz,
                      int64_t max_id, int64_t *stats, void *stream);

/* max over a uint32 volume (device result in out_dev[0]); helper for the host wrappers. */
SYN_API int syn_max_u32(const uint32_t *labels, int64_t n, uint32_t *out_dev, void *stream);

/*
 * syn_face_pair_edges -- proc/seg/merge/merge_ccs.py:117-128 (match_continuations):
 * for two facing 2-D faces (already in GLOBAL ids, 0 = nothing) emits the pair
 * (a[i], b[i]) wherever both are nonzero.  Duplicates are allowed (the union is
 * idempotent).  edges_out: uint32 [max_edges][2]; *n_edges_dev is incremented
 * atomically (caller zeroes it); entries beyond max_edges are dropped.
 */
SYN_API int syn_face_pair_edges(const uint32_t *face_a, const uint32_t *face_b, int64_t n,
                        uint32_t *edges_out, int64_t max_edges, unsigned long long *n_edges_dev, void *stream);

/*
 * syn_union_min_map -- proc/utils.py:39-86 (find_connected_components +
 * make_id_map): union-find over `n_edges` edges on ids < n_ids; map_out[i] = the
 * smallest id of i's component (map_out[i] = i for untouched ids).  The edge
 * count is read from n_edges_dev (device) and clamped to max_edges.
 */
SYN_API int syn_union_min_map(const uint32_t *edges, const unsigned long long *n_edges_dev, int64_t max_edges,
                      int64_t n_ids, uint32_t *map_out, void *stream);


/*
 * syn_threshold_dilate -- `mask = pred > thresh` (proc/seg/conncomps.py:17,43) then
 * seg_utils.dilate_by_k(mask, k) (seg_utils/misc.py:28-46 -> _seg_utils.cpp:25-82,
 * 180-200) for a boolean mask: out01 (uint32, same shape) receives 0/1, dilated by
 * k in 2-D Manhattan distance over axes 1 and 2.  dil_k = 0 gives the plain mask.
 */
SYN_API int syn_threshold_dilate(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int dil_k,
                                 uint32_t *out01, void *stream);

/*
 * syn_unique_u64 -- the distinct values of a uint64 array (np.unique of the base
 * segmentation, seg_utils/overlap.py:28, seg_utils/describe.py:8-11), unordered.
 * *count_dev receives the number of distinct values; if it exceeds cap the output
 * is truncated (retry with a larger cap).
 */
SYN_API int syn_unique_u64(const uint64_t *vals, int64_t n, uint64_t *out, int64_t cap,
                           unsigned long long *count_dev, void *stream);


/*
 * syn_ids_to_lut -- proc/seg/merge/assign_ids.py:34-42 (new_id_map) on the device:
 * lut[ids[i * stride]] = base + 1 + i for i < n, i.e. the chunk-local id of table row
 * i becomes global id base + 1 + i.  The caller zeroes `lut` first.
 */
SYN_API int syn_ids_to_lut(const int64_t *ids, int64_t n, int64_t stride, uint32_t base, uint32_t *lut,
                           int64_t lut_len, void *stream);

/*
 * syn_merge_seg_tables -- proc/seg/merge/merge_df.py:13-68 (merge_seginfo_df +
 * enforce_size_threshold).  rows: int64 [n_rows][SYN_SEG_NCOLS], row i = the table row
 * of GLOBAL id i+1; gmap: uint32 [n_rows+1] from syn_union_min_map.  merged_out receives
 * one row per merged id with size >= size_thr, ascending id (size = sum, bbox = min/max,
 * centroid = int(sum c_i * (s_i / S)) in float64, members in ascending id order, Kahan
 * summation); *n_merged_dev the row count; fmap_out [n_rows+1] the final id map
 * (merged ids under the threshold -> 0, merge_df.py:64-68).
 */
SYN_API int syn_merge_seg_tables(const int64_t *rows, int64_t n_rows, const uint32_t *gmap, int64_t size_thr,
                                 int64_t *merged_out, uint32_t *fmap_out, int64_t *n_merged_dev, void *stream);

/*
 * syn_merge_ccs_grid -- the device half of merge_ccs_task (proc/tasks.py:72-122) for a whole chunk grid in ONE
 * call: edges over every interior face pair (merge_ccs.py:69-128), union-find to the smallest id
 * (proc/utils.py:39-86), rows of all chunks in global id order (assign_ids.py:8-25), merged table + size threshold
 * (merge_df.py:13-68).  faces_all: uint32 [n_slots][facebuf], the six faces of every chunk ALREADY in global ids,
 * face order (axis 0 hi, axis 0 lo, axis 1 hi, ...), facebuf = 2 (ny nz + nx nz + nx ny); tables_all: int64
 * [n_slots][maxk][SYN_SEG_NCOLS]; slot_of_chunk / kept: HOST arrays over the chunks in np.ndindex order (where a
 * chunk's data sits in the gathered buffers, and its row count); grid_xyz / shape_xyz: HOST int64[3].
 * Outputs as syn_union_min_map (gmap_out [total+1]) and syn_merge_seg_tables; edges_scratch uint32
 * [max_edges][2], rows_scratch int64 [total][SYN_SEG_NCOLS].
 */
SYN_API int syn_merge_ccs_grid(const uint32_t *faces_all, const int64_t *tables_all, const int64_t *kept,
                               const int64_t *slot_of_chunk, const int64_t *grid_xyz, const int64_t *shape_xyz,
                               int64_t maxk, int64_t size_thr, uint32_t *edges_scratch, int64_t max_edges,
                               unsigned long long *n_edges_dev, uint32_t *gmap_out, int64_t *rows_scratch,
                               int64_t *merged_out, uint32_t *fmap_out, int64_t *n_merged_dev, void *stream);

/*
 * The chunk pass in two halves with the cross-chunk merge in between -- the single-box form of the reference's
 * chunk_ccs -> merge_ccs -> remap_ids -> chunk_overlaps workflow (proc/tasks.py:26-69, 72-122, 314-333, 290-297;
 * kube/task_generation/chunk_overlaps.py:14 runs the overlap task on the REMAPPED volume).  The label volume is
 * written once, already in merged global ids, and the overlaps are counted against those ids.
 *
 * A chunk's contribution to the merge is one PACKED SLOT of syn_packed_slot_bytes(cap_rows, cap_entries) bytes:
 *   int64 hdr[16]  : [0] rows of the segment table, [1] face entries, [2..7] first entry of face 0..5
 *                    (Face.all_faces() order, proc/seg/continuation.py:52-59), [8] flags (1 rows overflow,
 *                    2 entries overflow, 4 a capacity error of the chunk pass), [9..14] end of the entries of
 *                    face 0..5
 *   int64 table[cap_rows][SYN_SEG_NCOLS]   as syn_chunk_ccs' seg_table
 *   uint32 entry[cap_entries][2]           (position inside the face, table row + 1) of every face voxel that
 *                                          belongs to a surviving segment, raster order inside a face -- the
 *                                          content of the reference's Continuation objects (continuation.py:15-35)
 *
 * syn_chunk_ccs_begin  : syn_chunk_ccs up to the segment table; zero-fills the label sectors without foreground,
 *                        fills the packed slot.  face_mask: bit f (0..5) set = face f has a neighbour chunk and its
 *                        entries are wanted (the others stay empty); | SYN_BEGIN_SHARE_SM: the caller runs other
 *                        chunks' kernels on other streams at the same time, so the streaming kernel leaves room
 *                        for them on every SM.  No host synchronisation.
 * syn_merge_ccs_packed : merge_ccs_task (proc/tasks.py:72-122) over the gathered slots of ALL chunks:
 *                        id_base_out[c] = ids before chunk c in np.ndindex order (assign_ids.py:8-42; chunk c's row r
 *                        is global id id_base[c] + r + 1), edges between facing faces (merge_ccs.py:69-128),
 *                        union-find to the smallest id (proc/utils.py:39-86), merged table + size threshold
 *                        (merge_df.py:13-68).  plan_dev: device int32 [2 + nchunks + 4 npairs] =
 *                        {nchunks, npairs, slot of chunk c ..., (chunk here, hi face, chunk there, lo face) ...}.
 *                        merged_out int64 [nchunks cap_rows][SYN_SEG_NCOLS] (ascending id); fmap_out uint32
 *                        [nchunks cap_rows + 1]: global id -> final id (0 = under the size threshold);
 *                        info_out int64[8]: [0] total ids, [1] merged rows, [2] edges, [3] flags (1: a slot
 *                        overflowed, 2: max_edges too small).  No host synchronisation.  With merged_out NULL only
 *                        the id map is produced (what the label write waits for) and the caller runs
 * syn_merge_ccs_table  : the merged table alone (reads the id map left in `workspace` by syn_merge_ccs_packed), e.g.
 *                        on a second stream next to syn_chunk_ccs_finish.  owner_world > 1: only the rows of merged
 *                        ids that belong to a chunk of rank owner_rank (chunk c lives on rank c % owner_world) --
 *                        each rank then holds its share of the table; info[1] counts the rows written.
 * syn_chunk_ccs_finish : the rest of syn_chunk_ccs for a chunk started with syn_chunk_ccs_begin (same shape,
 *                        workspace, capacities, pointers): label of table row r = final_lut[final_lut_base_dev[0]
 *                        + r + 1] (final_lut NULL: r + 1), foreground label sectors, (label, base id) counts.
 */
#define SYN_BEGIN_SHARE_SM 0x100
SYN_API int64_t syn_packed_slot_bytes(int64_t cap_rows, int64_t cap_entries);
SYN_API int syn_chunk_ccs_begin(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int connectivity,
                                int dil_k, int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out,
                                const uint64_t *base_seg, int64_t max_pairs, void *packed_slot, int64_t cap_rows,
                                int64_t cap_entries, int face_mask, void *workspace, int64_t workspace_bytes,
                                int64_t max_runs, void *stream);
SYN_API int syn_chunk_ccs_finish(const float *pred, int64_t nx, int64_t ny, int64_t nz, int connectivity, int dil_k,
                                 uint32_t *labels_out, const uint64_t *base_seg, uint32_t *pair_rows,
                                 uint64_t *pair_cols, int64_t *pair_counts, int64_t max_pairs, int64_t cap_rows,
                                 const uint32_t *final_lut, const int64_t *final_lut_base_dev, int64_t final_lut_len,
                                 void *workspace, int64_t workspace_bytes, int64_t max_runs, int64_t *counts_host,
                                 void *stream);
/*
 * syn_push_slots -- the exchange of the packed slots over NVLink peer memory instead of an NCCL all_gather:
 * stores the valid part (header, kept rows, face entries) of this rank's n_slots slots at the SAME offsets into
 * the exchange buffers of the peers.  peer_dst: host array of n_peers device pointers (each peer's buffer,
 * already offset to this rank's first slot; the own buffer included); multicast_dst: the NVSwitch multicast
 * address of the same location (one store reaches every GPU), or NULL to store peer by peer.  The caller follows
 * up with a cross-rank barrier on the stream (e.g. the symmetric-memory barrier of torch.distributed).
 */
SYN_API int syn_push_slots(const void *my_slots, int64_t n_slots, int64_t slot_bytes, int64_t cap_rows,
                           int64_t cap_entries, void *const *peer_dst, int n_peers, void *multicast_dst, void *stream);
SYN_API int64_t syn_merge_packed_workspace_bytes(int64_t nchunks, int64_t cap_rows, int64_t max_edges);
SYN_API int syn_merge_ccs_table(const void *slots_all, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                                const int32_t *plan_dev, int64_t nchunks, int64_t size_thr, int64_t max_edges,
                                void *workspace, int64_t workspace_bytes, const int64_t *id_base, int64_t *merged_out,
                                int64_t *info, int owner_rank, int owner_world, void *stream);
SYN_API int syn_merge_ccs_packed(const void *slots_all, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                                 const int32_t *plan_dev, int64_t nchunks, int64_t npairs, int64_t size_thr,
                                 int64_t max_edges, void *workspace, int64_t workspace_bytes, int64_t *id_base_out,
                                 int64_t *merged_out, uint32_t *fmap_out, int64_t *info_out, void *stream);

/*
 * syn_merge_overlaps -- proc/tasks.py:300-311 (merge_overlaps_task: consolidate_overlaps,
 * overlap/merge/merge_overlaps.py:24-40, then find_max_overlaps, overlap/overlap.py:15-28).
 * rows / cols / counts: device arrays of n (label, base id, voxel count) triples of all chunks,
 * labels already in merged global ids (< n_ids; 0 = dropped, ignored).  Duplicates are summed;
 * argcol_out[id] = base id with the largest total for that label (ties: the smallest base id, the
 * first strict maximum of the (row, col)-sorted consolidated matrix), maxval_out[id] = that total
 * (0 and all-ones argcol where a label has no overlap).  Dense arrays of n_ids entries.
 */
SYN_API int syn_merge_overlaps(const uint32_t *rows, const uint64_t *cols, const int64_t *counts, int64_t n,
                               int64_t n_ids, uint64_t *argcol_out, int64_t *maxval_out, void *stream);

/*
 * syn_chunk_ccs_multilabel -- the `overlap_seg` variant of connected_components
 * (synaptor/proc/seg/conncomps.py:24-28, reached through cc_task(..., overlap_seg=...),
 * proc/tasks.py:41,64-67): `temp = zeros(uint64); temp[mask] = overlap_seg[mask]`, then a
 * MULTI-LABEL 6-connected labelling of temp -- two face-adjacent voxels join only when they
 * hold the same non-zero base id; voxels above threshold whose base id is 0 are background.
 * Numbering, dust filter, table and counts exactly as syn_chunk_ccs (no dilation, no pair
 * counting: the caller follows up with syn_count_overlaps for the `max_overlap` column).
 * overlap_seg: device uint64 [nx][ny][nz].
 */
SYN_API int syn_chunk_ccs_multilabel(const float *pred, const uint64_t *overlap_seg, int64_t nx, int64_t ny, int64_t nz,
                                     float thresh, int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out,
                                     int64_t *seg_table, int64_t max_segs, void *workspace, int64_t workspace_bytes,
                                     int64_t max_runs, int64_t *counts_host, void *stream);

/*
 * Measurement hooks (no reference counterpart).
 * syn_launch_count   : kernels launched by this library in this process so far.
 * syn_profile_enable : non-zero = syn_chunk_ccs records a CUDA event on its stream after every kernel;
 *                      0 = off.  Not thread-safe; for bench.py.
 * syn_profile_read   : synchronises on the last recorded call and returns the number K of
 *                      intervals (<= SYN_PROF_MAX), stage_ms[i] = elapsed ms of interval i.
 * syn_profile_name   : name of interval i of the last recorded call = the kernel that ended it
 *                      ("stream", "pair_init", "tile_ccl", "union_bnd", "flatten_rank", "stats_merge",
 *                      "dust_table", "run_labels", "sector_labels", "pair_count", "pair_bucket_scan",
 *                      "pair_scatter", "pair_emit"; other names on the generic / dilation path).
 */
#define SYN_PROF_MAX 40
SYN_API long long syn_launch_count(void);
SYN_API int syn_profile_enable(int on);
SYN_API int syn_profile_read(float *stage_ms, int n);
SYN_API const char *syn_profile_name(int i);

#ifdef __cplusplus
}
#endif
#endif /* SYNAPTOR_B200_H */
