// merge.cuh -- the cross-chunk merge with the chunk results packed for ONE exchange.
//
// Replaces, on the device (reference: synaptor/proc/tasks.py:72-122 merge_ccs_task):
//   assign_unique_ids_serial      proc/seg/merge/assign_ids.py:8-42   (global ids = prefix of the row counts)
//   extract_all_continuations     proc/seg/continuation.py:87-133     (face voxels of the surviving segments)
//   merge_continuations           proc/seg/merge/merge_ccs.py:69-139  (edges between facing faces)
//   find_connected_components /   proc/utils.py:39-86                 (union-find to the smallest id)
//     make_id_map
//   merge_seginfo_df /            proc/seg/merge/merge_df.py:13-68    (size sum, bbox min/max, weighted centroid,
//     enforce_size_threshold                                           size threshold)
//
// A chunk's contribution to the exchange is one fixed-size PACKED SLOT:
//   int64 hdr[16]   : [0] kept rows, [1] face entries, [2..7] first entry of face 0..5, [8] flags,
//                     [9..14] end of the entries of face 0..5
//   int64 table[cap_rows][11]   the chunk's segment table (written in place by the chunk pipeline)
//   uint2 entry[cap_entries]    (position in the face, chunk-relative id = table row + 1) of every labelled
//                               face voxel, faces in Face.all_faces() order, raster order inside a face
// Every rank gathers the slots of all chunks (one all_gather) and runs the same deterministic kernels on them.
#pragma once
#include <cooperative_groups.h>

#include "ccl.cuh"

namespace syn {

constexpr int PK_HDR = 16;              // int64 words
constexpr int PK_KEPT = 0, PK_NENT = 1, PK_FOFF = 2, PK_FLAGS = 8, PK_FEND = 9;
constexpr int PK_F_ROWS = 1, PK_F_ENTRIES = 2, PK_F_CCS = 4;

__host__ __device__ __forceinline__ size_t pk_slot_bytes(i64 cap_rows, i64 cap_entries)
{
    return (size_t)PK_HDR * 8 + (size_t)cap_rows * SYN_SEG_NCOLS * 8 + (size_t)cap_entries * 8;
}

// ---------------------------------------------------------------------------
// Face entries of one chunk, straight from the bitmask and the union-find arrays (the label volume need not
// exist yet): an ordered compaction (single-pass scan, k_scan_1p) over the 2 (ny nz + nx nz + nx ny) face voxels.
// ---------------------------------------------------------------------------
struct FaceCtx {
    const u32 *lmask, *omask;
    RunIdx R;
    const u32 *par, *rank, *complab;
    const Counts *cnt;
    Geom g;
    u32 area[6];        // voxels of face f (0 when the face is not wanted: it has no neighbour chunk)
    u32 start[7];       // first item of face f in the scan; start[6] = total

    // item i of the scan -> face number, position inside the face, label (32-bit arithmetic throughout)
    template <bool WITH_LABEL>
    __device__ __forceinline__ u32 label(u32 i, int *face, u32 *pos) const
    {
        int f = 0;
#pragma unroll
        for (int q = 1; q < 6; ++q) f += (i >= start[q]);
        const u32 j = i - start[f];
        const u32 nx = (u32)g.nx, ny = (u32)g.ny, nz = (u32)g.nz;
        u32 x, y, z;
        if (f < 2) {                        // axis 0: hi then lo; in-face coordinates (y, z)
            x = f == 0 ? nx - 1 : 0;
            y = j / nz;
            z = j - y * nz;
        } else if (f < 4) {                 // axis 1: (x, z)
            y = f == 2 ? ny - 1 : 0;
            x = j / nz;
            z = j - x * nz;
        } else {                            // axis 2: (x, y)
            z = f == 4 ? nz - 1 : 0;
            if (g.nys < 32) { x = j >> g.nys; y = j & (ny - 1u); }
            else { x = j / ny; y = j - x * ny; }
        }
        *face = f;
        *pos = j;
        const u32 row = x * ny + y, k = z >> 5, b = z & 31u;
        const u32 w = row * g.wz + k;
        if (!((omask[w] >> b) & 1u)) return 0u;
        // a foreground voxel ON a chunk face belongs to a component that touches the face, and those are never
        // dust (proc/tasks.py:45-51): its label is non-zero -- the counting pass needs nothing beyond the mask bit
        if (!WITH_LABEL) return 1u;
        const u32 run = R.at(w, row, k) + __popc(R.starts(lmask, w, k) & mask_le(b)) - 1u;
        return complab[rank[par[run]]];
    }
};

struct FaceSrc {
    FaceCtx F;
    __device__ __forceinline__ u64 n() const { return (u64)F.start[6]; }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
        const bool bad = (F.cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) != 0;
#pragma unroll
        for (int t = 0; t < N; ++t) {
            item[t] = 0;
            if (!bad && i0 + t < n) {
                int f;
                u32 pos;
                item[t] = F.label<false>((u32)(i0 + t), &f, &pos);
            }
        }
    }
};

struct FaceDst {
    FaceCtx F;
    i64 *hdr;           // packed-slot header
    uint2 *entries;
    i64 cap_entries, cap_rows;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
        if (i0 == 0) {
            const i64 kept = F.cnt->v[SYN_CNT_KEPT];
            hdr[PK_KEPT] = kept;
            i64 fl = 0;
            if (F.cnt->v[SYN_CNT_ERRFLAGS] & (SYN_EF_RUNS | SYN_EF_SEGS)) fl |= PK_F_CCS;
            if (kept > cap_rows) fl |= PK_F_ROWS;
            if (fl) atomicOr((u64 *)&hdr[PK_FLAGS], (u64)fl);
        }
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const u32 i = (u32)(i0 + t);
            if (i0 + t < n) {
                // entry range of every wanted face that starts / ends here (an unwanted face keeps 0 .. 0)
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    if (F.area[f] && i == F.start[f]) hdr[PK_FOFF + f] = (i64)ex;
                    if (F.area[f] && i == F.start[f] + F.area[f] - 1u) hdr[PK_FEND + f] = (i64)(ex + item[t]);
                }
                if (item[t]) {
                    int f;
                    u32 pos;
                    const u32 lab = F.label<true>(i, &f, &pos);      // (0 only after a capacity error upstream)
                    if ((i64)ex < cap_entries) entries[ex] = make_uint2(pos, lab);
                    else atomicOr((u64 *)&hdr[PK_FLAGS], (u64)PK_F_ENTRIES);
                }
            }
            ex += item[t];
        }
    }
};

__global__ void k_pack_header_only(i64 *hdr, const Counts *cnt, i64 cap_rows)
{
    if (threadIdx.x == 0) {
        const i64 kept = cnt->v[SYN_CNT_KEPT];
        hdr[PK_KEPT] = kept;
        i64 fl = 0;
        if (cnt->v[SYN_CNT_ERRFLAGS] & (SYN_EF_RUNS | SYN_EF_SEGS)) fl |= PK_F_CCS;
        if (kept > cap_rows) fl |= PK_F_ROWS;
        hdr[PK_FLAGS] = fl;
    }
}

// ---------------------------------------------------------------------------
// The merge proper.  `plan` (device int32): [0] nchunks, [1] npairs, slot_of_chunk[nchunks],
// pairs[npairs][4] = (chunk here, face here (hi), chunk there, face there (lo)).
// ---------------------------------------------------------------------------
constexpr int MG_MAXCHUNKS = 1024;
constexpr u32 MG_NONE = 0xFFFFFFFFu;

struct Packed {
    const unsigned char *all;
    i64 slot_bytes, cap_rows, cap_entries;
    __device__ __forceinline__ const i64 *hdr(int slot) const
    {
        return reinterpret_cast<const i64 *>(all + (size_t)slot * slot_bytes);
    }
    __device__ __forceinline__ const i64 *table(int slot) const { return hdr(slot) + PK_HDR; }
    __device__ __forceinline__ const uint2 *entries(int slot) const
    {
        return reinterpret_cast<const uint2 *>(hdr(slot) + PK_HDR + cap_rows * SYN_SEG_NCOLS);
    }
};

struct MergeAcc2 {
    u64 size;
    i64 lo[3], hi[3];
    u32 count;   // fragments
    u32 off;     // start of the member list
    u32 cursor;  // fill position
    u32 outrow;  // row in the compacted output (kept destinations only)
};

// info[]: 0 total ids, 1 merged rows, 2 edges, 3 flags (bit 0: a slot reported an overflow, bit 1: edge capacity)
constexpr int MG_TOTAL = 0, MG_NMERGED = 1, MG_NEDGES = 2, MG_FLAGS = 3;

// chunk of global id gid (1-based): the last c with base[c] < gid
__device__ __forceinline__ int mg_chunk_of(const i64 *s_base, int nchunks, i64 gid)
{
    int lo = 0, hi = nchunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_base[mid] < gid) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ const i64 *mg_row(const Packed &P, const int *plan, const i64 *s_base, int nchunks, i64 gid)
{
    const int c = mg_chunk_of(s_base, nchunks, gid);
    return P.table(plan[2 + c]) + (gid - 1 - s_base[c]) * SYN_SEG_NCOLS;
}

// output rows: exclusive scan of the kept flags (count > 0 and size >= size_thr) in id order, through the
// single-pass look-back scan (k_scan_1p); a destination with more than one fragment also claims its stretch of
// the member array here (a bump allocator: the stretches need not be in id order)
struct MgKeepSrc {
    const MergeAcc2 *acc;
    const i64 *info;
    i64 size_thr;
    __device__ __forceinline__ u64 n() const { return (info[MG_FLAGS] & 1) ? 0 : (u64)info[MG_TOTAL] + 1; }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            item[t] = 0;
            if (i0 + t < n && i0 + t > 0) {
                const MergeAcc2 *a = acc + i0 + t;
                item[t] = (a->count > 0) && ((i64)a->size >= size_thr);
            }
        }
    }
};
struct MgKeepDst {
    MergeAcc2 *acc;
    u32 *cursor;        // bump allocator of the member array
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            if (i0 + t < n && i0 + t > 0) {
                MergeAcc2 *a = acc + i0 + t;
                if (item[t]) a->outrow = ex;
                const u32 c = a->count;
                if (c > 1) a->off = atomicAdd(cursor, c);
            }
            ex += item[t];
        }
    }
};

__device__ __forceinline__ void mg_sort_u32(u32 *m, u32 n)
{
    if (n <= 16) {
        for (u32 i = 1; i < n; ++i) {
            const u32 v = m[i];
            int j = (int)i - 1;
            while (j >= 0 && m[j] > v) { m[j + 1] = m[j]; --j; }
            m[j + 1] = v;
        }
        return;
    }
    // heap sort: a component with many fragments must not cost O(n^2) in one thread
    auto sift = [&](u32 start, u32 end) {
        u32 root = start;
        while (2 * root + 1 < end) {
            u32 ch = 2 * root + 1;
            if (ch + 1 < end && m[ch] < m[ch + 1]) ++ch;
            if (m[root] >= m[ch]) return;
            const u32 t = m[root]; m[root] = m[ch]; m[ch] = t;
            root = ch;
        }
    };
    for (int s = (int)(n / 2) - 1; s >= 0; --s) sift((u32)s, n);
    for (u32 e = n - 1; e > 0; --e) {
        const u32 t = m[0]; m[0] = m[e]; m[e] = t;
        sift(0, e);
    }
}

// merged rows (merge_df.py:13-61): centroid = int(sum_i c_i * (s_i / S)) in float64, members in ascending id
// order, Kahan-compensated like pandas' groupby().sum()
__global__ void __launch_bounds__(256) k_mg_emit(Packed P, const int *__restrict__ plan, const i64 *__restrict__ base,
                                                 const i64 *__restrict__ info, const MergeAcc2 *__restrict__ acc,
                                                 u32 *members, i64 *__restrict__ out)
{
    __shared__ i64 s_base[MG_MAXCHUNKS + 1];
    const int nchunks = plan[0];
    for (int c = threadIdx.x; c <= nchunks; c += blockDim.x) s_base[c] = base[c];
    __syncthreads();
    if (info[MG_FLAGS] & 1) return;
    const i64 total = info[MG_TOTAL];
    for (i64 d = blockIdx.x * (i64)blockDim.x + threadIdx.x + 1; d <= total; d += (i64)gridDim.x * blockDim.x) {
        const MergeAcc2 a = acc[d];
        if (a.count == 0 || a.outrow == MG_NONE) continue;
        u32 self = (u32)d;                 // a destination is its own smallest member
        u32 *m = &self;
        if (a.count > 1) {
            m = members + a.off;
            mg_sort_u32(m, a.count);
        }
        double sum[3] = {0.0, 0.0, 0.0}, comp[3] = {0.0, 0.0, 0.0};
        const double S = (double)a.size;
        for (u32 i = 0; i < a.count; ++i) {
            const i64 *r = mg_row(P, plan, s_base, nchunks, (i64)m[i]);
            const double w = __ddiv_rn((double)r[SYN_SEG_SIZE], S);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double val = __dmul_rn((double)r[SYN_SEG_CX + k], w);
                const double y = val - comp[k];           // Kahan, as pandas' group_sum
                const double t = sum[k] + y;
                comp[k] = (t - sum[k]) - y;
                sum[k] = t;
            }
        }
        i64 *o = out + (i64)a.outrow * SYN_SEG_NCOLS;
        o[SYN_SEG_ID] = d;
        o[SYN_SEG_SIZE] = (i64)a.size;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            o[SYN_SEG_CX + k] = (i64)sum[k];              // astype(int): truncation
            o[SYN_SEG_BX + k] = a.lo[k];
            o[SYN_SEG_EX + k] = a.hi[k];
        }
    }
}


// ---------------------------------------------------------------------------
// The exchange over NVLink peer memory: every rank stores the VALID part of its packed slots (header, kept rows,
// face entries) straight into the exchange buffer of every peer -- one multimem.st per 16 bytes to the NVSwitch
// multicast address of the symmetric buffer when there is one (the switch replicates the store to all GPUs, so a
// rank sends its slot once, not world times), else one store per peer through the peers' mapped pointers.
// A cross-rank barrier (signal pads of the symmetric memory) follows on the stream.
// grid = (CTAs, slots of this rank).
// ---------------------------------------------------------------------------
struct PeerPtrs {
    unsigned char *p[16];
};

__device__ __forceinline__ void multimem_st_16(void *mc, uint4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(256) k_push_slots(const unsigned char *__restrict__ src, i64 slot_bytes, i64 cap_rows,
                                                    i64 cap_entries, PeerPtrs peers, int n_peers, unsigned char *mc)
{
    const unsigned char *slot = src + (size_t)blockIdx.y * slot_bytes;
    const i64 *hdr = reinterpret_cast<const i64 *>(slot);
    i64 kept = hdr[PK_KEPT], nent = hdr[PK_NENT];
    kept = kept < 0 ? 0 : (kept > cap_rows ? cap_rows : kept);
    nent = nent < 0 ? 0 : (nent > cap_entries ? cap_entries : nent);
    // two ranges of 16-byte units: [header + rows) and [entries)
    const i64 r0 = (PK_HDR * 8 + kept * SYN_SEG_NCOLS * 8 + 15) / 16;
    const i64 e0 = (PK_HDR * 8 + cap_rows * SYN_SEG_NCOLS * 8) / 16, e1 = e0 + (nent * 8 + 15) / 16;
    const i64 n = r0 + (e1 - e0);
    const size_t off = (size_t)blockIdx.y * slot_bytes;
    // four independent 16-byte load -> store chains per thread and round: the stores cross NVLink, the more bytes
    // in flight the better
    const i64 nth = (i64)gridDim.x * blockDim.x;
    for (i64 i0 = blockIdx.x * (i64)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * nth) {
        uint4 v[4];
        i64 u[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const i64 i = i0 + t * nth;
            u[t] = i < r0 ? i : e0 + (i - r0);
            if (i < n) v[t] = *reinterpret_cast<const uint4 *>(slot + u[t] * 16);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (i0 + t * nth >= n) continue;
            if (mc) multimem_st_16(mc + off + u[t] * 16, v[t]);
            else {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q < n_peers) *reinterpret_cast<uint4 *>(peers.p[q] + off + u[t] * 16) = v[t];
            }
        }
    }
    __threadfence_system();
}

constexpr int MG_MAXPAIRS_SMEM = 3072;

// ---------------------------------------------------------------------------
// The part of the merge a chunk's label write waits for -- id bases, edges, union-find, merged sizes, final id
// map -- as one cooperative kernel (k_mg_ids); the merged TABLE (bbox, centroid, output rows) is produced by the
// split kernels (k_mg_tab_*, k_mg_emit) on a second stream, next to the label kernels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mg_ids(Packed P, const int *__restrict__ plan, i64 *__restrict__ base,
                                                i64 *__restrict__ info, u32 *par, u32 *__restrict__ gmap,
                                                u64 *msize, u32 *__restrict__ edges, i64 max_edges, i64 size_thr,
                                                u32 *__restrict__ fmap)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ i64 s_base[MG_MAXCHUNKS + 1];
    __shared__ u32 s_units[MG_MAXPAIRS_SMEM + 1];
    __shared__ u32 sm[33];
    __shared__ u32 s_bad;
    const int nchunks = plan[0], npairs = plan[1];
    const int lane = threadIdx.x & 31;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    {
        u32 k[4], v = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = (int)threadIdx.x * 4 + q;
            k[q] = 0;
            if (c < nchunks) {
                const i64 *h = P.hdr(plan[2 + c]);
                i64 kept = h[PK_KEPT];
                if (h[PK_FLAGS] || kept < 0 || kept > P.cap_rows || h[PK_NENT] > P.cap_entries) { s_bad = 1; kept = 0; }
                k[q] = (u32)kept;
            }
            v += k[q];
        }
        u32 tot;
        u32 ex = block_excl_scan(v, sm, &tot);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = (int)threadIdx.x * 4 + q;
            if (c < nchunks) s_base[c] = (i64)ex;
            ex += k[q];
        }
        if (threadIdx.x == 0) s_base[nchunks] = (i64)tot;
    }
    __syncthreads();
    const i64 total = s_base[nchunks];
    const bool bad = s_bad != 0;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c <= nchunks; c += blockDim.x) base[c] = s_base[c];
        if (threadIdx.x == 0) {
            info[MG_TOTAL] = total;
            info[MG_NMERGED] = 0;
            info[MG_NEDGES] = 0;
            info[MG_FLAGS] = bad ? 1 : 0;
        }
    }
    if (bad) return;                   // every CTA sees the same headers: a uniform exit, nobody waits at a barrier
    for (i64 i = tid; i <= total; i += nth) {
        par[i] = (u32)i;
        msize[i] = 0;
    }
    grid.sync();
    if (npairs > 0) {
        // ---- edges between facing faces; work unit = 256 hi-face entries of one pair ----
        const bool tab = npairs <= MG_MAXPAIRS_SMEM;
        u32 nunits = 0;
        if (tab) {
            if (threadIdx.x == 0) {
                u32 run = 0;
                for (int p = 0; p < npairs; ++p) {
                    const int *pr = plan + 2 + nchunks + 4 * p;
                    const i64 *hA = P.hdr(plan[2 + pr[0]]);
                    const i64 nA = hA[PK_FEND + pr[1]] - hA[PK_FOFF + pr[1]];
                    s_units[p] = run;
                    run += (u32)((nA + 255) / 256);
                }
                s_units[npairs] = run;
            }
            __syncthreads();
            nunits = s_units[npairs];
        }
        const u32 loops = tab ? nunits : (u32)npairs;
        for (u32 u = tab ? blockIdx.x : 0u; u < loops; u += tab ? gridDim.x : 1u) {
            int p;
            i64 i;
            if (tab) {
                int lo = 0, hi = npairs - 1;       // last pair with s_units[p] <= u
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_units[mid] <= u) lo = mid;
                    else hi = mid - 1;
                }
                p = lo;
                i = (i64)(u - s_units[p]) * 256 + threadIdx.x;
            } else {
                p = (int)u;
                i = (i64)blockIdx.x * 256 + threadIdx.x;
            }
            const int *pr = plan + 2 + nchunks + 4 * p;
            const int cA = pr[0], fA = pr[1], cB = pr[2], fB = pr[3];
            const int sA = plan[2 + cA], sB = plan[2 + cB];
            const i64 *hA = P.hdr(sA), *hB = P.hdr(sB);
            const i64 oA = hA[PK_FOFF + fA], nA = hA[PK_FEND + fA] - oA;
            const i64 oB = hB[PK_FOFF + fB], nB = hB[PK_FEND + fB] - oB;
            const uint2 *A = P.entries(sA) + oA, *B = P.entries(sB) + oB;
            const u32 bA = (u32)s_base[cA], bB = (u32)s_base[cB];
            const i64 step = tab ? nA + 256 : (i64)gridDim.x * 256;      // tab: one round per unit
            for (i64 i0 = i - lane; i0 < nA; i0 += step) {
                const i64 ii = i0 + lane;
                u32 a = 0, b = 0;
                if (ii < nA) {
                    const uint2 e = A[ii];
                    i64 lo = 0, hi = nB;
                    while (lo < hi) {
                        const i64 mid = (lo + hi) >> 1;
                        if (B[mid].x < e.x) lo = mid + 1;
                        else hi = mid;
                    }
                    if (lo < nB) {
                        const uint2 f = B[lo];
                        if (f.x == e.x && e.y && f.y) { a = bA + e.y; b = bB + f.y; }
                    }
                }
                bool act = a != 0;
                const u32 pa = __shfl_up_sync(SYN_FULL, a, 1), pb = __shfl_up_sync(SYN_FULL, b, 1);
                if (lane > 0 && pa == a && pb == b) act = false;
                const u32 m = __ballot_sync(SYN_FULL, act);
                if (!m) continue;
                u64 basei = 0;
                if (lane == 0) basei = atomicAdd((u64 *)&info[MG_NEDGES], (u64)__popc(m));
                basei = __shfl_sync(SYN_FULL, (unsigned long long)basei, 0);
                if (act) {
                    const u64 e = basei + __popc(m & mask_lt((u32)lane));
                    if ((i64)e < max_edges) { edges[2 * e] = a; edges[2 * e + 1] = b; }
                }
            }
        }
        grid.sync();
        i64 ne = *(volatile i64 *)&info[MG_NEDGES];
        if (ne > max_edges) {
            if (tid == 0) atomicOr((u64 *)&info[MG_FLAGS], 2ull);
            ne = max_edges;
        }
        for (i64 i = tid; i < ne; i += nth) {
            const u32 a = __ldcg(edges + 2 * i), b = __ldcg(edges + 2 * i + 1);
            if ((i64)a <= total && (i64)b <= total) uf_union(par, a, b);
        }
        grid.sync();
    }
    // ---- destination of every id, merged sizes ----
    for (i64 gid = tid; gid <= total; gid += nth) {
        if (gid == 0) { gmap[0] = 0; continue; }
        const u32 d = uf_find_ro(par, (u32)gid);
        gmap[gid] = d;
        atomicAdd(msize + d, (u64)mg_row(P, plan, s_base, nchunks, gid)[SYN_SEG_SIZE]);
    }
    grid.sync();
    // ---- final id map: merged segments under the size threshold map to 0 (merge_df.py:64-68) ----
    for (i64 gid = tid; gid <= total; gid += nth) {
        const u32 d = gid ? __ldcg(gmap + gid) : 0u;
        fmap[gid] = (gid && (i64)__ldcg(msize + d) >= size_thr) ? d : 0u;
    }
}

__global__ void __launch_bounds__(256) k_mg_tab_init(const i64 *__restrict__ info, MergeAcc2 *__restrict__ acc,
                                                     u64 *__restrict__ scan_zero, u32 n_scan_zero)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_scan_zero; i += gridDim.x * blockDim.x) scan_zero[i] = 0;
    if (info[MG_FLAGS] & 1) return;
    const i64 total = info[MG_TOTAL];
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= total; i += (i64)gridDim.x * blockDim.x) {
        MergeAcc2 a;
        a.size = 0;
        a.lo[0] = a.lo[1] = a.lo[2] = 0x7FFFFFFFFFFFFFFFll;
        a.hi[0] = a.hi[1] = a.hi[2] = -0x7FFFFFFFFFFFFFFFll - 1;
        a.count = a.off = a.cursor = 0;
        a.outrow = MG_NONE;
        acc[i] = a;
    }
}

// `world` > 1: only the destinations whose id belongs to a chunk of rank `rank` (chunk c lives on rank c % world)
// are accumulated -- every other destination keeps count 0 and is skipped by the scan, the member lists and the
// rows: each rank produces ITS share of the merged table instead of all ranks producing all of it.
__global__ void __launch_bounds__(256) k_mg_tab_acc(Packed P, const int *__restrict__ plan, const i64 *__restrict__ base,
                                                    const i64 *__restrict__ info, const u32 *__restrict__ gmap,
                                                    MergeAcc2 *acc, int rank, int world)
{
    __shared__ i64 s_base[MG_MAXCHUNKS + 1];
    const int nchunks = plan[0];
    for (int c = threadIdx.x; c <= nchunks; c += blockDim.x) s_base[c] = base[c];
    __syncthreads();
    if (info[MG_FLAGS] & 1) return;
    const i64 total = info[MG_TOTAL];
    for (i64 gid = blockIdx.x * (i64)blockDim.x + threadIdx.x + 1; gid <= total; gid += (i64)gridDim.x * blockDim.x) {
        const u32 d = gmap[gid];
        if (world > 1 && mg_chunk_of(s_base, nchunks, (i64)d) % world != rank) continue;
        const i64 *r = mg_row(P, plan, s_base, nchunks, gid);
        MergeAcc2 *a = acc + d;
        atomicAdd(&a->size, (u64)r[SYN_SEG_SIZE]);
        atomicAdd(&a->count, 1u);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&a->lo[k], r[SYN_SEG_BX + k]);
            atomicMax(&a->hi[k], r[SYN_SEG_EX + k]);
        }
    }
}

__global__ void __launch_bounds__(256) k_mg_tab_fill(const i64 *__restrict__ info, const u32 *__restrict__ gmap,
                                                     MergeAcc2 *acc, u32 *__restrict__ members)
{
    if (info[MG_FLAGS] & 1) return;
    const i64 total = info[MG_TOTAL];
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x + 1; i <= total; i += (i64)gridDim.x * blockDim.x) {
        MergeAcc2 *a = acc + gmap[i];
        if (a->count > 1) members[a->off + atomicAdd(&a->cursor, 1u)] = (u32)i;
    }
}

}  // namespace syn
