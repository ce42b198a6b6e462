// aux.cuh -- the smaller entry points: chunk faces, LUT relabel, statistics of an
// arbitrary label volume, and the cross-chunk merge primitives.
#pragma once
#include "common.cuh"

namespace syn {

// ---------------------------------------------------------------------------
// six faces in Face.all_faces() order (proc/seg/continuation.py:52-59)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_extract_faces(const u32 *__restrict__ labels, i64 nx, i64 ny, i64 nz,
                                                       u32 *__restrict__ out)
{
    const i64 a0 = ny * nz, a1 = nx * nz, a2 = nx * ny;
    const i64 total = 2 * (a0 + a1 + a2);
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        i64 j = i;
        i64 x, y, z;
        if (j < 2 * a0) {  // axis 0: hi then lo; in-face coords (y, z)
            bool hi = j < a0;
            if (!hi) j -= a0;
            x = hi ? nx - 1 : 0;
            y = j / nz;
            z = j - y * nz;
        } else if (j < 2 * a0 + 2 * a1) {  // axis 1: in-face coords (x, z)
            j -= 2 * a0;
            bool hi = j < a1;
            if (!hi) j -= a1;
            y = hi ? ny - 1 : 0;
            x = j / nz;
            z = j - x * nz;
        } else {  // axis 2: in-face coords (x, y)
            j -= 2 * a0 + 2 * a1;
            bool hi = j < a2;
            if (!hi) j -= a2;
            z = hi ? nz - 1 : 0;
            x = j / ny;
            y = j - x * ny;
        }
        out[i] = labels[(x * ny + y) * nz + z];
    }
}

// ---------------------------------------------------------------------------
// [Z][Y][X] -> [X][Y][Z]: what np.ascontiguousarray does to the F-ordered chunk CloudVolume hands the reference's
// tasks (proc/seg/conncomps.py:21, io/backends/cloudvolume.py:6-20), done on the device after the bytes were
// copied as they are.  32 x 32 tiles of one y-plane through (padded) shared memory: reads coalesced along x,
// writes coalesced along z.  grid = (ceil(nx/32), ceil(nz/32), ny), block = (32, 8).
// ---------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_transpose_zyx(const T *__restrict__ src, T *__restrict__ dst, i64 nx, i64 ny,
                                                       i64 nz)
{
    __shared__ T tile[32][33];
    const i64 y = blockIdx.z;
    const i64 x0 = (i64)blockIdx.x * 32, z0 = (i64)blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const i64 z = z0 + threadIdx.y + j, x = x0 + threadIdx.x;
        if (z < nz && x < nx) tile[threadIdx.y + j][threadIdx.x] = src[(z * ny + y) * nx + x];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const i64 x = x0 + threadIdx.y + j, z = z0 + threadIdx.x;
        if (x < nx && z < nz) dst[(x * ny + y) * nz + z] = tile[threadIdx.x][threadIdx.y + j];
    }
}

// ---------------------------------------------------------------------------
// relabel through a dense table (_relabel.cpp:19-38 with the dict as a LUT)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_relabel_lut(u32 *__restrict__ d, i64 n, const u32 *__restrict__ lut,
                                                     i64 lut_len)
{
    const i64 tid = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const i64 nt = (i64)gridDim.x * blockDim.x;
    const i64 n4 = ((reinterpret_cast<uintptr_t>(d) & 15) == 0) ? (n >> 2) : 0;
    uint4 *d4 = reinterpret_cast<uint4 *>(d);
    const u32 l0 = lut_len > 0 ? lut[0] : 0u;  // a dict may map 0 too (`if v in mapping`, every voxel)
    for (i64 i = tid; i < n4; i += nt) {
        uint4 v = d4[i];
        if (l0 == 0 && (v.x | v.y | v.z | v.w) == 0) continue;  // background: nothing to write
        uint4 o = v;
        if ((i64)v.x < lut_len) o.x = lut[v.x];
        if ((i64)v.y < lut_len) o.y = lut[v.y];
        if ((i64)v.z < lut_len) o.z = lut[v.z];
        if ((i64)v.w < lut_len) o.w = lut[v.w];
        if (o.x != v.x || o.y != v.y || o.z != v.z || o.w != v.w) d4[i] = o;
    }
    for (i64 i = n4 * 4 + tid; i < n; i += nt) {
        u32 v = d[i];
        if ((i64)v < lut_len) {
            u32 o = lut[v];
            if (o != v) d[i] = o;
        }
    }
}

__global__ void __launch_bounds__(256) k_max_u32(const u32 *__restrict__ d, i64 n, u32 *out)
{
    u32 m = 0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        m = max(m, d[i]);
    m = __reduce_max_sync(SYN_FULL, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ---------------------------------------------------------------------------
// statistics of an arbitrary label volume with ids <= max_id
// (seg_utils/describe.py:14-80).  Lane <-> voxel of a 32-voxel word; lanes with
// equal labels are merged with match_any before the atomics.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_init_stats(CompStat *stat, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        CompStat s;
        s.size = s.sx = s.sy = s.sz = 0;
        s.minx = s.miny = s.minz = 0xFFFFFFFFu;
        s.maxx = s.maxy = s.maxz = 0;
        s.keep = 0;
        s.row = 0;
        stat[i] = s;
    }
}

__global__ void __launch_bounds__(256) k_label_stats(const u32 *__restrict__ labels, Geom g, i64 max_id,
                                                     CompStat *stat)
{
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 w = warp; w < g.nwords; w += nwarps) {
        u32 row = w / g.wz;
        u32 kw = w - row * g.wz;
        i64 z = (i64)kw * 32 + lane;
        u32 a = 0;
        if (z < g.nz) a = labels[(i64)row * g.nz + z];
        if ((i64)a > max_id) a = 0;
        if (__ballot_sync(SYN_FULL, a != 0) == 0) continue;
        u32 peers = __match_any_sync(SYN_FULL, a);
        if (a == 0) continue;
        int leader = __ffs(peers) - 1;
        if (lane != leader) continue;
        u32 x = row / (u32)g.ny, y = row - x * (u32)g.ny;
        u32 n = __popc(peers);
        u32 z0 = kw * 32;
        CompStat *s = stat + a;
        atomicAdd(&s->size, (u64)n);
        atomicAdd(&s->sx, (u64)n * x);
        atomicAdd(&s->sy, (u64)n * y);
        atomicAdd(&s->sz, (u64)n * z0 + bitpos_sum(peers));
        atomicMin(&s->minx, x);
        atomicMin(&s->miny, y);
        atomicMin(&s->minz, z0 + (u32)__ffs(peers) - 1u);
        atomicMax(&s->maxx, x);
        atomicMax(&s->maxy, y);
        atomicMax(&s->maxz, z0 + 31u - (u32)__clz(peers));
    }
}

__device__ __forceinline__ i64 centroid_f32_aux(u64 sum, u64 size)
{
    float q = __fdiv_rn(__ll2float_rn((i64)sum), __ll2float_rn((i64)size));
    return (i64)roundf(q);
}

__global__ void __launch_bounds__(256) k_stats_to_table(const CompStat *__restrict__ stat, i64 n, i64 *table)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const CompStat s = stat[i];
        i64 *r = table + i * SYN_SEG_NCOLS;
        r[SYN_SEG_ID] = i;
        r[SYN_SEG_SIZE] = (i64)s.size;
        if (s.size == 0) {
            for (int c = SYN_SEG_CX; c < SYN_SEG_NCOLS; ++c) r[c] = 0;
            continue;
        }
        r[SYN_SEG_CX] = centroid_f32_aux(s.sx, s.size);
        r[SYN_SEG_CY] = centroid_f32_aux(s.sy, s.size);
        r[SYN_SEG_CZ] = centroid_f32_aux(s.sz, s.size);
        r[SYN_SEG_BX] = s.minx;
        r[SYN_SEG_BY] = s.miny;
        r[SYN_SEG_BZ] = s.minz;
        r[SYN_SEG_EX] = (i64)s.maxx + 1;
        r[SYN_SEG_EY] = (i64)s.maxy + 1;
        r[SYN_SEG_EZ] = (i64)s.maxz + 1;
    }
}

// ---------------------------------------------------------------------------
// cross-chunk merge primitives (proc/seg/merge/merge_ccs.py:117-128, proc/utils.py:39-86)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_face_pair_edges(const u32 *__restrict__ fa, const u32 *__restrict__ fb,
                                                         i64 n, u32 *__restrict__ edges, i64 max_edges, u64 *n_edges)
{
    const int lane = threadIdx.x & 31;
    for (i64 i0 = (blockIdx.x * (i64)blockDim.x + threadIdx.x) - lane; i0 < n; i0 += (i64)gridDim.x * blockDim.x) {
        i64 i = i0 + lane;
        u32 a = 0, b = 0;
        if (i < n) { a = fa[i]; b = fb[i]; }
        bool act = a && b;
        // drop a pair equal to the previous lane's (faces are raster ordered: long identical runs)
        u32 pa = __shfl_up_sync(SYN_FULL, a, 1), pb = __shfl_up_sync(SYN_FULL, b, 1);
        if (lane > 0 && pa == a && pb == b) act = false;
        u32 m = __ballot_sync(SYN_FULL, act);
        if (!m) continue;
        u64 basei = 0;
        if (lane == 0) basei = atomicAdd(n_edges, (u64)__popc(m));
        basei = __shfl_sync(SYN_FULL, basei, 0);
        if (act) {
            u64 e = basei + __popc(m & mask_lt((u32)lane));
            if ((i64)e < max_edges) { edges[2 * e] = a; edges[2 * e + 1] = b; }
        }
    }
}

__global__ void __launch_bounds__(256) k_iota_u32(u32 *p, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = (u32)i;
}

__global__ void __launch_bounds__(256) k_union_edges(const u32 *__restrict__ edges, const u64 *n_edges,
                                                     i64 max_edges, i64 n_ids, u32 *par)
{
    i64 n = (i64)*n_edges;
    if (n > max_edges) n = max_edges;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        u32 a = edges[2 * i], b = edges[2 * i + 1];
        if ((i64)a < n_ids && (i64)b < n_ids) uf_union(par, a, b);
    }
}

__global__ void __launch_bounds__(256) k_flatten(u32 *par, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        par[i] = uf_find_ro(par, (u32)i);
}


// ---------------------------------------------------------------------------
// bitmask -> uint32 0/1 volume (standalone seg_utils.dilate_by_k on a mask)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_expand_mask(const u32 *__restrict__ mask, Geom g, u32 *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 w = warp; w < g.nwords; w += nwarps) {
        u32 row = w / g.wz;
        u32 kw = w - row * g.wz;
        i64 z = (i64)kw * 32 + lane;
        if (z < g.nz) out[(i64)row * g.nz + z] = (mask[w] >> lane) & 1u;
    }
}

// ---------------------------------------------------------------------------
// distinct values of a uint64 array (np.unique of the base segmentation,
// seg_utils/overlap.py:28): open-addressing hash set, 64-bit CAS.  Consecutive
// equal values (the common case in a segmentation) are collapsed with a ballot
// before the table is touched.  Output order is arbitrary (host sorts the few
// distinct values).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unique_u64(const u64 *__restrict__ vals, i64 n, u64 *table, u32 capmask,
                                                    u64 *__restrict__ out, i64 cap, u64 *count)
{
    const int lane = threadIdx.x & 31;
    const u64 EMPTY = ~0ull;
    for (i64 i0 = (blockIdx.x * (i64)blockDim.x + threadIdx.x) - lane; i0 < n; i0 += (i64)gridDim.x * blockDim.x) {
        i64 i = i0 + lane;
        u64 v = i < n ? __ldcs(vals + i) : EMPTY;
        u64 pv = __shfl_up_sync(SYN_FULL, v, 1);
        bool act = (i < n) && (lane == 0 || pv != v);
        if (!act) continue;
        // EMPTY itself (2^64-1) cannot be stored as a key: report it through slot-less path
        if (v == EMPTY) {
            // claim a dedicated flag in table[capmask + 1]
            if (atomicCAS(table + (u64)capmask + 1, 0ull, 1ull) == 0ull) {
                u64 k = atomicAdd(count, 1ull);
                if ((i64)k < cap) out[k] = v;
            }
            continue;
        }
        u32 h = (u32)mix64(0x51ED27ull, v) & capmask;
        for (u32 probe = 0; probe <= capmask; ++probe) {
            u64 cur = table[h];
            if (cur == v) break;
            if (cur == EMPTY) {
                cur = atomicCAS(table + h, EMPTY, v);
                if (cur == EMPTY) {
                    u64 k = atomicAdd(count, 1ull);
                    if ((i64)k < cap) out[k] = v;
                    break;
                }
                if (cur == v) break;
            }
            h = (h + 1) & capmask;
            if (probe == capmask) atomicAdd(count, (u64)cap + 1);  // table full: force the overflow signal
        }
    }
}

__global__ void __launch_bounds__(256) k_fill_u64(u64 *p, i64 n, u64 v)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}


// ---------------------------------------------------------------------------
// Cross-chunk merge of the seg-info tables (proc/seg/merge/merge_df.py:13-68,
// assign_ids.py:8-42).  rows[i] is the table row of GLOBAL id i+1 (chunks
// concatenated in np.ndindex order); gmap[gid] = smallest id of gid's merged
// component.  Per destination id: size = sum, bbox = min / max (integer atomics,
// order free); centroid = int(sum_i c_i * (s_i / S)) in float64 with the members
// taken in ascending id order and Kahan-compensated like pandas' groupby().sum().
// ---------------------------------------------------------------------------
struct MergeAcc {
    u64 size;
    i64 lo[3], hi[3];
    u32 count;   // fragments
    u32 off;     // start of the member list
    u32 cursor;  // fill position
    u32 outrow;  // row in the compacted output (kept destinations only)
};

__global__ void __launch_bounds__(256) k_merge_init(MergeAcc *acc, i64 n)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= n; i += (i64)gridDim.x * blockDim.x) {
        MergeAcc a;
        a.size = 0;
        a.lo[0] = a.lo[1] = a.lo[2] = 0x7FFFFFFFFFFFFFFFll;
        a.hi[0] = a.hi[1] = a.hi[2] = -0x7FFFFFFFFFFFFFFFll - 1;
        a.count = a.off = a.cursor = a.outrow = 0;
        acc[i] = a;
    }
}

__global__ void __launch_bounds__(256) k_merge_accumulate(const i64 *__restrict__ rows, i64 n,
                                                          const u32 *__restrict__ gmap, MergeAcc *acc)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 *r = rows + i * SYN_SEG_NCOLS;
        MergeAcc *a = acc + gmap[i + 1];
        atomicAdd(&a->size, (u64)r[SYN_SEG_SIZE]);
        atomicAdd(&a->count, 1u);
        for (int k = 0; k < 3; ++k) {
            atomicMin(&a->lo[k], r[SYN_SEG_BX + k]);
            atomicMax(&a->hi[k], r[SYN_SEG_EX + k]);
        }
    }
}

// one block: exclusive scans of the fragment counts (member list offsets) and of the kept flags (output rows)
__global__ void __launch_bounds__(SCAN_THREADS) k_merge_scan(MergeAcc *acc, i64 n, i64 size_thr, i64 *n_merged)
{
    __shared__ u32 sm[33];
    u32 carry_c = 0, carry_k = 0;
    for (i64 b0 = 0; b0 <= n; b0 += SCAN_THREADS) {
        const i64 i = b0 + threadIdx.x;
        u32 c = 0, k = 0;
        if (i <= n && i > 0) {
            c = acc[i].count;
            k = (c > 0) && ((i64)acc[i].size >= size_thr);
        }
        u32 tc, tk;
        u32 ec = block_excl_scan(c, sm, &tc);
        u32 ek = block_excl_scan(k, sm, &tk);
        if (i <= n && i > 0) {
            acc[i].off = carry_c + ec;
            acc[i].outrow = k ? (carry_k + ek) : 0xFFFFFFFFu;
        }
        carry_c += tc;
        carry_k += tk;
    }
    if (threadIdx.x == 0) *n_merged = (i64)carry_k;
}

__global__ void __launch_bounds__(256) k_merge_fill(i64 n, const u32 *__restrict__ gmap, MergeAcc *acc,
                                                    u32 *__restrict__ members, u32 *__restrict__ fmap)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= n; i += (i64)gridDim.x * blockDim.x) {
        if (i == 0) { fmap[0] = 0; continue; }
        MergeAcc *a = acc + gmap[i];
        members[a->off + atomicAdd(&a->cursor, 1u)] = (u32)i;
        fmap[i] = a->outrow != 0xFFFFFFFFu ? gmap[i] : 0u;   // merged segments under the size threshold map to 0
    }
}

__global__ void __launch_bounds__(256) k_merge_emit(const i64 *__restrict__ rows, i64 n, const MergeAcc *__restrict__ acc,
                                                    u32 *members, i64 *__restrict__ out)
{
    for (i64 d = blockIdx.x * (i64)blockDim.x + threadIdx.x + 1; d <= n; d += (i64)gridDim.x * blockDim.x) {
        const MergeAcc a = acc[d];
        if (a.count == 0 || a.outrow == 0xFFFFFFFFu) continue;
        u32 *m = members + a.off;
        for (u32 i = 1; i < a.count; ++i) {   // ascending id order: the fill order is arbitrary
            u32 v = m[i];
            int j = (int)i - 1;
            while (j >= 0 && m[j] > v) { m[j + 1] = m[j]; --j; }
            m[j + 1] = v;
        }
        double sum[3] = {0.0, 0.0, 0.0}, comp[3] = {0.0, 0.0, 0.0};
        const double S = (double)a.size;
        for (u32 i = 0; i < a.count; ++i) {
            const i64 *r = rows + (i64)(m[i] - 1) * SYN_SEG_NCOLS;
            const double w = __ddiv_rn((double)r[SYN_SEG_SIZE], S);
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
        for (int k = 0; k < 3; ++k) {
            o[SYN_SEG_CX + k] = (i64)sum[k];              // astype(int): truncation
            o[SYN_SEG_BX + k] = a.lo[k];
            o[SYN_SEG_EX + k] = a.hi[k];
        }
    }
}

// lut[ids[i * stride]] = base + 1 + i  (chunk-local id -> global id of its table row; assign_ids.py:34-42)
__global__ void __launch_bounds__(256) k_ids_to_lut(const i64 *__restrict__ ids, i64 n, i64 stride, u32 base,
                                                    u32 *__restrict__ lut, i64 lut_len)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 id = ids[i * stride];
        if (id >= 0 && id < lut_len) lut[id] = base + 1u + (u32)i;
    }
}

}  // namespace syn
