// finalize.cuh -- the dense output pass and the overlap counter.
//
//   k_write_labels : bitmask + per-run labels -> uint32 label volume (4 B/voxel written,
//                    the only dense write of the pipeline), optionally fused with the
//                    (label, base id) pair histogram against a uint64 base segmentation
//                    (8 B/voxel read) -- seg_utils/overlap.py:10-61.
//   k_overlap_scan : the same histogram for an arbitrary uint32 label volume
//                    (seg_utils.count_overlaps as a standalone call).
//   ordering       : Counter insertion order (overlap.py:47) == ascending first-occurrence
//                    voxel index; recovered with a bitmap + popcount scan.
#pragma once
#include "ccl.cuh"

namespace syn {

struct PairTable {
    PairKey *keys;   // [cap] 16-byte keys, all-ones = empty
    u64 *count;      // [cap]
    u64 *first;      // [cap] smallest padded bit position (row * wz * 32 + z) of the pair
    u32 capmask;     // cap - 1 (cap is a power of two)
    Counts *cnt;
};

// Adds `n` occurrences of (label, base) first seen at padded position `pos`.
__device__ __forceinline__ void pair_insert(const PairTable &T, u32 label, u64 base, u32 n, u64 pos)
{
    const PairKey want{base, (u64)label};
    const PairKey empty{~0ull, ~0ull};
    u32 h = (u32)mix64(label, base) & T.capmask;
    for (u32 probe = 0; probe <= T.capmask; ++probe) {
        // Keys are write-once (empty -> key).  A plain 16 B read can only be stale or torn
        // while the slot is being claimed, and then at least one half still reads all-ones:
        // in that case the 128-bit CAS is the authority.
        PairKey cur = T.keys[h];
        bool hit = cur.base == want.base && cur.label == want.label;
        if (!hit && (cur.base == ~0ull || cur.label == ~0ull)) {
            cur = cas128(T.keys + h, empty, want);
            if (cur.base == empty.base && cur.label == empty.label) {  // claimed
                atomicAdd((u64 *)&T.cnt->v[SYN_CNT_TABLE_SLOTS], 1ull);
                hit = true;
            } else {
                hit = cur.base == want.base && cur.label == want.label;
            }
        }
        if (hit) {
            atomicAdd(T.count + h, (u64)n);
            // `first` only ever decreases: a (possibly stale) read that is already below pos makes the atomic useless
            if (pos < __ldcg(T.first + h)) atomicMin(T.first + h, pos);
            return;
        }
        h = (h + 1) & T.capmask;
    }
    atomicOr((u64 *)&T.cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_TABLE);  // table full
}

__global__ void __launch_bounds__(256) k_pair_table_init(PairTable T)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        T.keys[i] = PairKey{~0ull, ~0ull};
        T.count[i] = 0;
        T.first[i] = ~0ull;
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
                                                      const u32 *__restrict__ runbase,
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
                    const u32 carry = kw ? (lmask[w - 1] >> 31) : 0u;
                    const u32 starts = run_starts(m[u], carry);
                    const u32 rb = runbase[w];
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
                        u32 carry = kw ? (lmask[w - 1] >> 31) : 0u;
                        u32 run = runbase[w] + __popc(run_starts(mm, carry) & mask_le((u32)lane)) - 1u;
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
// Foreground sectors.  A warp takes 32 active mask words, prefix-sums their non-empty bytes and then walks the
// batch's sectors 32 at a time, ONE LANE PER SECTOR: owner word by a shuffle binary search over the prefix,
// byte by find-nth-set.  A sector holds at most 4 pieces of runs: their labels are fetched together, the (at
// most 8) base ids of the foreground voxels too, then the whole 32-byte sector is stored.
template <bool HAS_BASE>
__global__ void __launch_bounds__(256) k_sector_labels(const u32 *__restrict__ lmask, const u32 *__restrict__ omask,
                                                       RunIdx R,
                                                       const u32 *__restrict__ runlabel, u32 *__restrict__ labels,
                                                       const u64 *__restrict__ base, Geom g, PairTable T,
                                                       const u32 *__restrict__ actw, const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u32 nact = (u32)cnt->v[SYN_CNT_ACTIVE_WORDS];
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i0 = warp * 32; i0 < nact; i0 += nwarps * 32) {
        const u32 i = i0 + lane;
        u32 w = 0, om = 0, m = 0, starts = 0, rb = 0;
        i64 vox0 = 0;
        if (i < nact) {
            w = actw[i];
            om = omask[w];
            m = lmask[w];
            u32 row, kw;
            split_word(g, w, row, kw);
            const u32 carry = kw ? (lmask[w - 1] >> 31) : 0u;
            starts = run_starts(m, carry);
            rb = R.at(w, row, kw);
            vox0 = (i64)row * g.nz + (i64)kw * 32;
        }
        const u32 f = (u32)((om & 0xFFu) != 0) | ((u32)((om & 0xFF00u) != 0) << 1) |
                      ((u32)((om & 0xFF0000u) != 0) << 2) | ((u32)((om & 0xFF000000u) != 0) << 3);
        const u32 n = __popc(f);
        const u32 incl = warp_incl_scan(n, lane);
        const u32 excl = incl - n;
        const u32 total = __shfl_sync(SYN_FULL, incl, 31);
        for (u32 r0 = 0; r0 < total; r0 += 32) {
            const u32 j = r0 + lane;
            int o = 0;   // owner = last lane whose exclusive prefix is <= j
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                const u32 e = __shfl_sync(SYN_FULL, excl, (o + s) & 31);
                if (e <= j) o += s;
            }
            const u32 e_o = __shfl_sync(SYN_FULL, excl, o);
            const u32 f_o = __shfl_sync(SYN_FULL, f, o);
            const u32 om_o = __shfl_sync(SYN_FULL, om, o);
            const u32 m_o = __shfl_sync(SYN_FULL, m, o);
            const u32 st_o = __shfl_sync(SYN_FULL, starts, o);
            const u32 rb_o = __shfl_sync(SYN_FULL, rb, o);
            const u32 w_o = __shfl_sync(SYN_FULL, w, o);
            const i64 v_o = __shfl_sync(SYN_FULL, vox0, o);
            if (j >= total) continue;
            const u32 q8 = 8u * select_bit(f_o, j - e_o);      // first bit of the sector inside the word
            const u32 ob = (om_o >> q8) & 0xFFu;               // voxels that receive a label
            const u32 mb = (m_o >> q8) & 0xFFu;                // labelled mask (differs from ob only with dilation)
            const i64 vox = v_o + q8;
            // base ids of the foreground voxels: up to four 16-byte loads, issued before the label lookups
            ulonglong2 bq[4];
            if (HAS_BASE) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bq[u] = make_ulonglong2(0, 0);
                    if ((ob >> (2 * u)) & 3u) bq[u] = __ldcs(reinterpret_cast<const ulonglong2 *>(base + vox) + u);
                }
            }
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
            u32 lab[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const u32 pi = __popc(ps & mask_le((u32)t));   // 1-based piece index of voxel t
                const u32 l = pi <= 1 ? plab[0] : pi == 2 ? plab[1] : pi == 3 ? plab[2] : plab[3];
                lab[t] = ((ob >> t) & 1u) ? l : 0u;
            }
            uint4 *lp = reinterpret_cast<uint4 *>(labels + vox);
            lp[0] = make_uint4(lab[0], lab[1], lab[2], lab[3]);
            lp[1] = make_uint4(lab[4], lab[5], lab[6], lab[7]);
            if (HAS_BASE) {
                const u64 bv[8] = {bq[0].x, bq[0].y, bq[1].x, bq[1].y, bq[2].x, bq[2].y, bq[3].x, bq[3].y};
                PairAcc acc;
                const u64 pos0 = (u64)w_o * 32 + q8;
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (lab[t] && bv[t]) acc.add(T, lab[t], bv[t], pos0 + t);
                acc.flush(T);
            }
        }
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
// ordering of the table entries by first occurrence
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_mark(PairTable T, u32 *__restrict__ bitmap)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        u64 f = T.first[i];
        if (f != ~0ull) atomicOr(bitmap + (f >> 5), 1u << (f & 31));
    }
}

struct PopcSrc {
    const u32 *bitmap;
    u32 nwords;
    __device__ __forceinline__ u64 n() const { return nwords; }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
        u32 m[N];
        load_items_u32(bitmap, i0, n, m);
#pragma unroll
        for (int t = 0; t < N; ++t) item[t] = __popc(m[t]);
    }
};
struct WordBaseDst {
    u32 *wordbase;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
        store_prefix_u32(wordbase, i0, n, ex, item);
    }
};

__global__ void __launch_bounds__(256) k_pair_emit(PairTable T, const u32 *__restrict__ bitmap,
                                                   const u32 *__restrict__ wordbase, u32 *__restrict__ rows,
                                                   u64 *__restrict__ cols, i64 *__restrict__ vals, i64 max_pairs)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.capmask; i += gridDim.x * blockDim.x) {
        u64 f = T.first[i];
        if (f == ~0ull) continue;
        u32 w = (u32)(f >> 5), b = (u32)(f & 31);
        i64 r = (i64)wordbase[w] + __popc(bitmap[w] & mask_lt(b));
        if (r >= max_pairs) {
            atomicOr((u64 *)&T.cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_PAIRS);
            continue;
        }
        PairKey k = T.keys[i];
        rows[r] = (u32)k.label;
        cols[r] = k.base;
        vals[r] = (i64)T.count[i];
    }
}

}  // namespace syn
