// finalize.cuh -- the dense output pass and the overlap counter.
//
//   k_write_labels : bitmask + per-run labels -> uint32 label volume (4 B/voxel written,
//                    the only dense write of the pipeline), optionally fused with the
//                    (label, base id) pair histogram against a uint64 base segmentation
//                    (8 B/voxel read) -- seg_utils/overlap.py:10-61.
//   k_overlap_scan : the same histogram for an arbitrary uint32 label volume
//                    (seg_utils.count_overlaps as a standalone call).
//   ordering       : Counter insertion order (overlap.py:47) == ascending first-occurrence
//                    voxel index; recovered with a bucket sort on the first positions.
#pragma once
#include "ccl.cuh"

namespace syn {


// 32-bit mixing of (label, base): cheap (three 32-bit multiplies) and good enough for open addressing
__device__ __forceinline__ u32 pair_hash(u32 label, u64 base)
{
    u32 h = label * 0x9E3779B1u ^ (u32)base * 0x85EBCA77u ^ (u32)(base >> 32) * 0xC2B2AE3Du;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// Adds `n` occurrences of (label, base) first seen at padded position `pos`.
__device__ __forceinline__ void pair_insert(const PairTable &T, u32 label, u64 base, u32 n, u64 pos)
{
    const PairKey want{base, (u64)label};
    const PairKey empty{~0ull, ~0ull};
    u32 h = pair_hash(label, base) & T.capmask;
    for (u32 probe = 0; probe <= T.capmask; ++probe) {
        // Keys are write-once (empty -> key).  A plain 16 B read can only be stale or torn
        // while the slot is being claimed, and then at least one half still reads all-ones:
        // in that case the 128-bit CAS is the authority.
        u64 cb, cl, cf;                                       // key and (stale) first position: one 32 B read
        [[maybe_unused]] u64 cc;                              // (the count travels along)
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(cb), "=l"(cl), "=l"(cc), "=l"(cf) : "l"(T.slots + h));
        PairKey cur{cb, cl};
        bool hit = cur.base == want.base && cur.label == want.label;
        if (!hit && (cur.base == ~0ull || cur.label == ~0ull)) {
            cur = cas128(&T.slots[h].key, empty, want);
            if (cur.base == empty.base && cur.label == empty.label) {  // claimed
                atomicAdd((u64 *)&T.cnt->v[SYN_CNT_TABLE_SLOTS], 1ull);
                hit = true;
            } else {
                hit = cur.base == want.base && cur.label == want.label;
            }
            cf = ~0ull;
        }
        if (hit) {
            atomicAdd(&T.slots[h].count, (u64)n);
            // the first position only ever decreases: a (possibly stale) value that is already smaller makes the
            // atomic unnecessary
            if (pos < cf) atomicMin(&T.slots[h].first, pos);
            return;
        }
        h = (h + 1) & T.capmask;
    }
    atomicOr((u64 *)&T.cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_TABLE);  // table full
}

__global__ void __launch_bounds__(256) k_pair_table_init(PairTable T)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        T.slots[i] = PairSlot{PairKey{~0ull, ~0ull}, 0ull, ~0ull};
    }
}

// per-lane accumulation of up to 4 consecutive voxels into pair inserts
struct PairAcc {
    u32 lab = 0;
    u64 base = 0;
    u32 n = 0;
    u64 pos = 0;
    __device__ __forceinline__ void add(const PairTable &T, u32 l, u64 b, u64 p)
    {
        if (n && l == lab && b == base) { ++n; return; }
        flush(T);
        lab = l; base = b; n = 1; pos = p;
    }
    __device__ __forceinline__ void flush(const PairTable &T)
    {
        if (n) pair_insert(T, lab, base, n, pos);
        n = 0;
    }
};

// ===========================================================================
// Dense output pass.  One warp per 128-voxel chunk of a row (as k_threshold).
// VEC: lane l owns voxels 4l..4l+3 -> one 16 B label store, two 16 B base loads.
// ===========================================================================
template <bool VEC, bool HAS_BASE, int UNROLL>
__global__ void __launch_bounds__(256) k_write_labels(const u32 *__restrict__ lmask, const u32 *__restrict__ omask,
                                                      RunIdx R,
                                                      const u32 *__restrict__ runlabel, u32 *__restrict__ labels,
                                                      const u64 *__restrict__ base, Geom g, PairTable T, Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u64 bmax = 0;
    for (u32 c0 = warp * UNROLL; c0 < g.nchunks; c0 += nwarps * UNROLL) {
        if (VEC) {
            ulonglong2 b01[UNROLL], b23[UNROLL];
            u32 m[UNROLL], om[UNROLL], widx[UNROLL];
            i64 vox[UNROLL];
            bool ok[UNROLL];
            const int sh = 4 * (lane & 7);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                u32 c = c0 + u;
                ok[u] = false;
                m[u] = om[u] = 0;
                widx[u] = 0;
                vox[u] = 0;
                if (c < g.nchunks) {
                    u32 row = c / g.cpr;
                    u32 k4 = (c - row * g.cpr) * 4;
                    i64 z = (i64)k4 * 32 + 4 * lane;
                    if (z < g.nz) {
                        ok[u] = true;
                        widx[u] = row * g.wz + k4 + (lane >> 3);
                        vox[u] = (i64)row * g.nz + z;
                        m[u] = lmask[widx[u]];
                        om[u] = omask[widx[u]];
                        if (HAS_BASE) {
                            const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(base + vox[u]);
                            b01[u] = __ldcs(bp);
                            b23[u] = __ldcs(bp + 1);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (!ok[u]) continue;
                u32 lab[4] = {0, 0, 0, 0};
                const u32 nib = (om[u] >> sh) & 0xFu;
                if (nib) {
                    const u32 w = widx[u];
                    const u32 kw = w % g.wz;
                    const u32 starts = R.starts(lmask, w, kw);
                    const u32 rb = R.at(w, w / g.wz, kw);
                    u32 prev_run = 0xFFFFFFFFu, prev_lab = 0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if ((nib >> t) & 1u) {
                            u32 run = rb + __popc(starts & mask_le((u32)(sh + t))) - 1u;
                            if (run != prev_run) { prev_run = run; prev_lab = runlabel[run]; }
                            lab[t] = prev_lab;
                        }
                    }
                }
                __stcs(reinterpret_cast<uint4 *>(labels + vox[u]), make_uint4(lab[0], lab[1], lab[2], lab[3]));
                if (HAS_BASE) {
                    const u64 bv[4] = {b01[u].x, b01[u].y, b23[u].x, b23[u].y};
                    bmax = max(bmax, max(max(bv[0], bv[1]), max(bv[2], bv[3])));
                    if (nib) {
                        PairAcc acc;
                        const u64 pos0 = (u64)widx[u] * 32 + sh;
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (lab[t] && bv[t]) acc.add(T, lab[t], bv[t], pos0 + t);
                        acc.flush(T);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int u = 0; u < UNROLL; ++u) {
                u32 c = c0 + u;
                if (c >= g.nchunks) break;
                u32 row = c / g.cpr;
                u32 k4 = (c - row * g.cpr) * 4;
                for (int j = 0; j < 4; ++j) {
                    u32 kw = k4 + j;
                    if (kw >= g.wz) break;
                    i64 z = (i64)kw * 32 + lane;
                    if (z >= g.nz) continue;
                    u32 w = row * g.wz + kw;
                    u32 mm = lmask[w], oo = omask[w];
                    u32 lab = 0;
                    if ((oo >> lane) & 1u) {
                        (void)mm;
                        u32 run = R.at(w, row, kw) + __popc(R.starts(lmask, w, kw) & mask_le((u32)lane)) - 1u;
                        lab = runlabel[run];
                    }
                    i64 v = (i64)row * g.nz + z;
                    labels[v] = lab;
                    if (HAS_BASE) {
                        u64 b = __ldcs(base + v);
                        bmax = max(bmax, b);
                        if (lab && b) pair_insert(T, lab, b, 1, (u64)w * 32 + lane);
                    }
                }
            }
        }
    }
    if (HAS_BASE) {
        for (int d = 16; d; d >>= 1) bmax = max(bmax, __shfl_down_sync(SYN_FULL, bmax, d));
        if (lane == 0 && bmax) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_BASE], bmax);
    }
}


// ===========================================================================
// Output pass, sector mode (nz % 8 == 0: a 32-byte sector of the label volume = 8 voxels
// = one byte of a mask word, never straddling two rows).  k_stream has already zero-filled
// every sector WITHOUT foreground; k_sector_labels writes the sectors that DO hold
// foreground, whole (kept labels or 0), one lane per sector, and feeds the (label, base)
// pair histogram.  The two kernels write disjoint sectors.
// ===========================================================================
// Foreground sectors.  A warp takes a batch of SL_BATCH = 128 active mask words (4 per lane, all loads of the batch in flight together),
// stages the per-word data in its slice of shared memory and writes one entry per non-empty byte into a SECTOR
// LIST; the lanes then walk the list 32 sectors at a time, ONE LANE PER SECTOR, with no search.  A sector holds
// at most 4 pieces of runs: their labels are fetched together, the (at most 8) base ids of the foreground voxels
// too, then the whole 32-byte sector is stored (one 256-bit store) and the (label, base)
// groups go to the pair table.
constexpr int SL_WPL = 4;                   // words per lane per batch
constexpr int SL_BATCH = 32 * SL_WPL;
constexpr int SL_WARPS = 8;

struct SlWarp {                             // per-warp staging area
    u32 w[SL_BATCH], om[SL_BATCH], m[SL_BATCH], st[SL_BATCH], rb[SL_BATCH];
    unsigned short list[SL_BATCH * 4];      // slot | byte << 7
};

static_assert(sizeof(SlWarp) % 16 == 0, "staging area alignment");

template <bool HAS_BASE>
__global__ void __launch_bounds__(32 * SL_WARPS) k_sector_labels(const u32 *__restrict__ lmask,
                                                                const u32 *__restrict__ omask, RunIdx R,
                                                                const u32 *__restrict__ runlabel,
                                                                u32 *__restrict__ labels, const u64 *__restrict__ base,
                                                                Geom g, PairTable T, const u32 *__restrict__ actw,
                                                                const Counts *cnt)
{
    __shared__ SlWarp s_warp[SL_WARPS];
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u32 nact = (u32)cnt->v[SYN_CNT_ACTIVE_WORDS];
    const int lane = threadIdx.x & 31;
    const u32 wib = threadIdx.x >> 5;
    SlWarp &S = s_warp[wib];
    const u32 nbatch = (nact + SL_BATCH - 1) / SL_BATCH;
    for (u32 bi = blockIdx.x * SL_WARPS + wib; bi < nbatch; bi += gridDim.x * SL_WARPS) {
        // ---- stage the batch: slot = q * 32 + lane <-> active word bi * 128 + slot ----
        u32 w[SL_WPL], om[SL_WPL], m[SL_WPL], pv[SL_WPL], rbv[SL_WPL];
        bool ok[SL_WPL];
#pragma unroll
        for (int q = 0; q < SL_WPL; ++q) {
            const u32 i = bi * SL_BATCH + q * 32 + lane;
            ok[q] = i < nact;
            w[q] = ok[q] ? actw[i] : 0u;
        }
#pragma unroll
        for (int q = 0; q < SL_WPL; ++q) {
            om[q] = m[q] = pv[q] = rbv[q] = 0;
            if (ok[q]) {
                u32 row, kw;
                split_word(g, w[q], row, kw);
                om[q] = omask[w[q]];
                m[q] = lmask[w[q]];
                pv[q] = kw ? lmask[w[q] - 1] : 0u;
                rbv[q] = R.at(w[q], row, kw);
            }
        }
        u32 packed = 0;     // non-empty bytes per slot, 8 bits each (a warp total is <= 128)
        u32 f[SL_WPL];
#pragma unroll
        for (int q = 0; q < SL_WPL; ++q) {
            f[q] = (u32)((om[q] & 0xFFu) != 0) | ((u32)((om[q] & 0xFF00u) != 0) << 1) |
                   ((u32)((om[q] & 0xFF0000u) != 0) << 2) | ((u32)((om[q] & 0xFF000000u) != 0) << 3);
            packed |= (u32)__popc(f[q]) << (8 * q);
        }
        const u32 incl = warp_incl_scan(packed, lane);
        const u32 tot = __shfl_sync(SYN_FULL, incl, 31);
        const u32 excl = incl - packed;
        u32 total = 0;
#pragma unroll
        for (int q = 0; q < SL_WPL; ++q) {
            const u32 slot = q * 32 + lane;
            S.w[slot] = w[q];
            S.om[slot] = om[q];
            S.m[slot] = m[q];
            S.st[slot] = run_starts(m[q], pv[q] >> 31);
            S.rb[slot] = rbv[q];
            u32 pos = total + ((excl >> (8 * q)) & 0xFFu);
            u32 rest = f[q];
            while (rest) {
                const u32 t = (u32)__ffs(rest) - 1u;
                rest &= rest - 1;
                S.list[pos++] = (unsigned short)(slot | (t << 7));
            }
            total += (tot >> (8 * q)) & 0xFFu;
        }
        __syncwarp();
        // ---- one lane per sector ----
        for (u32 r0 = 0; r0 < total; r0 += 32) {
            const u32 j = r0 + lane;
            if (j >= total) continue;
            const u32 e = S.list[j];
            const u32 slot = e & 127u;
            const u32 q8 = 8u * (e >> 7);                      // first bit of the sector inside the word
            const u32 w_o = S.w[slot], om_o = S.om[slot], m_o = S.m[slot], st_o = S.st[slot], rb_o = S.rb[slot];
            u32 row, kw;
            split_word(g, w_o, row, kw);
            const i64 vox = (i64)row * g.nz + (i64)kw * 32 + q8;
            const u32 ob = (om_o >> q8) & 0xFFu;               // voxels that receive a label
            const u32 mb = (m_o >> q8) & 0xFFu;                // labelled mask (differs from ob only with dilation)
            // pieces of runs inside the byte (<= 4) and their labels
            const u32 ps = mb & ~(mb << 1) & 0xFFu;
            u32 plab[4];
            {
                u32 rest = ps;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    plab[p] = 0;
                    if (rest) {
                        const u32 t = (u32)__ffs(rest) - 1u;
                        rest &= rest - 1;
                        plab[p] = runlabel[rb_o + __popc(st_o & mask_le(q8 + t)) - 1u];
                    }
                }
            }
            // base ids of the foreground voxels: up to four 16-byte loads -- only for a sector that keeps a label
            // (the sectors of dust components, a large share of all foreground sectors, need none)
            ulonglong2 bq[4];
            if (HAS_BASE) {
                const bool any = (plab[0] | plab[1] | plab[2] | plab[3]) != 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bq[u] = make_ulonglong2(0, 0);
                    if (any && ((ob >> (2 * u)) & 3u)) bq[u] = __ldcs(reinterpret_cast<const ulonglong2 *>(base + vox) + u);
                }
            }
            u32 lab[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const u32 pi = __popc(ps & mask_le((u32)t));   // 1-based piece index of voxel t
                const u32 l = pi <= 1 ? plab[0] : pi == 2 ? plab[1] : pi == 3 ? plab[2] : plab[3];
                lab[t] = ((ob >> t) & 1u) ? l : 0u;
            }
            st_256(labels + vox, lab);
            if (HAS_BASE) {
                const u64 bv[8] = {bq[0].x, bq[0].y, bq[1].x, bq[1].y, bq[2].x, bq[2].y, bq[3].x, bq[3].y};
                const u64 pos0 = (u64)w_o * 32 + q8;
                // groups of voxels with the same (label, base) -> pair cache
                u32 gl = 0, gn = 0;
                u64 gb = 0, gp = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (lab[t] && bv[t]) {
                        if (gn && lab[t] == gl && bv[t] == gb) { ++gn; continue; }
                        if (gn) pair_insert(T, gl, gb, gn, gp);
                        gl = lab[t]; gb = bv[t]; gn = 1; gp = pos0 + t;
                    }
                }
                if (gn) pair_insert(T, gl, gb, gn, gp);
            }
        }
        __syncwarp();   // the staging area is rewritten by the next batch
    }
}

// ===========================================================================
// Standalone overlap scan over an arbitrary uint32 label volume + uint64 base.
// Also reduces max(seg1) and max(seg2) (shape of the COO matrix, overlap.py:41-42).
// Lane <-> voxel within a 32-voxel word; consecutive lanes with the same
// (label, base) pair are merged with a ballot before touching the table.
// ===========================================================================
__global__ void __launch_bounds__(256) k_overlap_scan(const u32 *__restrict__ seg1, const u64 *__restrict__ seg2,
                                                      Geom g, PairTable T, Counts *cnt)
{
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u32 amax = 0;
    u64 bmax = 0;
    for (u32 w = warp; w < g.nwords; w += nwarps) {
        u32 row = w / g.wz;
        u32 kw = w - row * g.wz;
        i64 z = (i64)kw * 32 + lane;
        u32 a = 0;
        u64 b = 0;
        if (z < g.nz) {
            i64 v = (i64)row * g.nz + z;
            a = __ldcs(seg1 + v);
            b = __ldcs(seg2 + v);
        }
        amax = max(amax, a);
        bmax = max(bmax, b);
        bool act = a && b;
        u32 actm = __ballot_sync(SYN_FULL, act);
        if (actm == 0) continue;
        // head of a group = active lane whose predecessor lane holds a different pair
        u32 pa = __shfl_up_sync(SYN_FULL, a, 1);
        u64 pb = __shfl_up_sync(SYN_FULL, b, 1);
        bool head = act && (lane == 0 || pa != a || pb != b);
        u32 headm = __ballot_sync(SYN_FULL, head);
        if (head) {
            // group extends to the next head or the next inactive lane
            u32 stop = (headm | ~actm) & ~mask_le((u32)lane);
            int end = stop ? (__ffs(stop) - 1) : 32;
            pair_insert(T, a, b, (u32)(end - lane), (u64)w * 32 + lane);
        }
    }
    amax = __reduce_max_sync(SYN_FULL, amax);
    for (int d = 16; d; d >>= 1) bmax = max(bmax, __shfl_down_sync(SYN_FULL, bmax, d));
    if (lane == 0) {
        if (amax) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_LABEL], (u64)amax);
        if (bmax) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_BASE], bmax);
    }
}

// ---------------------------------------------------------------------------
// Ordering of the table entries by first occurrence (Counter insertion order, overlap.py:47 == ascending
// first position).  Bucket sort on the position: bucket = position >> PB_SHIFT (4096 voxel positions, so a
// bucket holds a handful of pairs at most), counts -> exclusive scan -> scatter into bucket order -> the rank
// inside a bucket by counting the smaller positions of the same bucket.  Four tiny kernels over the table
// (a few 100 k slots) instead of a bitmap + popcount scan over every voxel position.
// ---------------------------------------------------------------------------
constexpr int PB_SHIFT = 12;

__global__ void __launch_bounds__(256) k_pair_count(PairTable T, u32 *__restrict__ bcount)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        const u64 f = T.slots[i].first;
        if (f != ~0ull) atomicAdd(bcount + (f >> PB_SHIFT), 1u);
    }
}

// bucket counts -> bucket bases: generic single-pass scan (k_scan_1p), in place
struct U32Src {
    const u32 *a;
    u32 len;
    __device__ __forceinline__ u64 n() const { return len; }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const { load_items_u32(a, i0, n, item); }
};
struct U32PrefixDst {
    u32 *a;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
        store_prefix_u32(a, i0, n, ex, item);
    }
};

__global__ void __launch_bounds__(256) k_pair_scatter(PairTable T, const u32 *__restrict__ bbase, u32 *__restrict__ bfill,
                                                      u32 *__restrict__ sorted_slot, i64 max_pairs)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        const u64 f = T.slots[i].first;
        if (f == ~0ull) continue;
        const u32 b = (u32)(f >> PB_SHIFT);
        const u32 j = bbase[b] + atomicAdd(bfill + b, 1u);
        if ((i64)j < max_pairs) sorted_slot[j] = i;
        else atomicOr((u64 *)&T.cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_PAIRS);
    }
}

__global__ void __launch_bounds__(256) k_pair_emit(PairTable T, const u32 *__restrict__ bbase, u32 nb,
                                                   const u32 *__restrict__ sorted_slot, u32 *__restrict__ rows,
                                                   u64 *__restrict__ cols, i64 *__restrict__ vals, i64 max_pairs)
{
    const u32 np = (u32)T.cnt->v[SYN_CNT_PAIRS];
    if ((i64)np > max_pairs) return;     // capacity error already flagged by the scatter
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < np; j += gridDim.x * blockDim.x) {
        const u32 slot = sorted_slot[j];
        const u64 f = T.slots[slot].first;
        const u32 b = (u32)(f >> PB_SHIFT);
        const u32 lo = bbase[b], hi = (b + 1 < nb) ? bbase[b + 1] : np;
        u32 r = lo;
        for (u32 t = lo; t < hi; ++t) r += (T.slots[sorted_slot[t]].first < f);
        const PairKey k = T.slots[slot].key;
        rows[r] = (u32)k.label;
        cols[r] = k.base;
        vals[r] = (i64)T.slots[slot].count;
    }
}

// ---------------------------------------------------------------------------
// merge_overlaps (proc/tasks.py:300-311, overlap/merge/merge_overlaps.py:24-40, overlap.py:15-28): the chunks'
// (label, base id, count) triples, labels already in merged global ids, are summed per pair in the pair table;
// per label the base id with the largest total wins, ties go to the smallest base id (the consolidated matrix is
// sorted by (row, col) and find_max_overlaps keeps the first strict maximum).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mo_insert(const u32 *__restrict__ rows, const u64 *__restrict__ cols,
                                                   const i64 *__restrict__ counts, i64 n, PairTable T)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const u32 r = rows[i];
        const i64 c = counts[i];
        if (r == 0 || c <= 0) continue;          // label 0 = dropped by the merge (size threshold)
        pair_insert(T, r, cols[i], (u32)c, (u64)i);
    }
}

__global__ void __launch_bounds__(256) k_mo_init(u64 *__restrict__ argcol, i64 *__restrict__ maxval, i64 n_ids)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n_ids; i += (i64)gridDim.x * blockDim.x) {
        argcol[i] = ~0ull;
        maxval[i] = 0;
    }
}

__global__ void __launch_bounds__(256) k_mo_rowmax(PairTable T, i64 *__restrict__ maxval, i64 n_ids)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        const PairKey k = T.slots[i].key;
        if (k.label == ~0ull || (i64)k.label >= n_ids) continue;
        atomicMax((u64 *)maxval + k.label, T.slots[i].count);
    }
}

__global__ void __launch_bounds__(256) k_mo_rowarg(PairTable T, const i64 *__restrict__ maxval,
                                                   u64 *__restrict__ argcol, i64 n_ids)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        const PairKey k = T.slots[i].key;
        if (k.label == ~0ull || (i64)k.label >= n_ids) continue;
        if ((i64)T.slots[i].count == maxval[k.label]) atomicMin(argcol + k.label, k.base);
    }
}

}  // namespace syn
