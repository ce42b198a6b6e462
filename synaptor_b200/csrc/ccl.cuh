// ccl.cuh -- the labeler: threshold -> bitmask -> z-runs -> union-find ->
// canonical (raster-first) numbering -> per-component statistics -> dust filter.
//
// Replaces, on the device:
//   mask = d > thresh                                  synaptor/proc/seg/conncomps.py:17
//   cc3d.connected_components(mask, connectivity=6)    conncomps.py:22  (third-party labeler)
//   seg_utils.dilate_by_k                              seg_utils/_seg_utils.cpp:25-82,180-200
//   segment_sizes / centers_of_mass / bounding_boxes   seg_utils/describe.py:14-80, _describe.cpp:23-64
//   filter_segs_by_size(to_ignore=continuation ids)    seg_utils/filtering.py:7-40, proc/tasks.py:45-51
#pragma once
#include "common.cuh"

namespace syn {

// ===========================================================================
// K1: float32 predictions -> 1 bit per voxel.  One warp per 128-voxel chunk of a
// row; VEC path: one 128-bit streaming load per lane (4 voxels), the 4-bit
// results are OR-reduced over each group of 8 lanes into a 32-bit mask word.
// HBM traffic: 4 B/voxel read, 1/8 B/voxel written.
// ===========================================================================
template <bool VEC, int UNROLL>
__global__ void __launch_bounds__(256) k_threshold(const float *__restrict__ pred, u32 *__restrict__ mask,
                                                   Geom g, float thr, u64 *__restrict__ fg_count)
{
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u32 nfg = 0;
    for (u32 c0 = warp * UNROLL; c0 < g.nchunks; c0 += nwarps * UNROLL) {
        if (VEC) {
            float4 v[UNROLL];
            bool ok[UNROLL];
            u32 widx[UNROLL];
            u32 row = c0 / g.cpr;            // one division per iteration: the UNROLL chunks are consecutive
            u32 kc = c0 - row * g.cpr;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 z = (i64)kc * 128 + 4 * lane;
                ok[u] = (c0 + u < g.nchunks) && z < g.nz;
                widx[u] = row * g.wz + kc * 4 + (lane >> 3);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[u]) v[u] = __ldcs(reinterpret_cast<const float4 *>(pred + (i64)row * g.nz + z));
                if (++kc == g.cpr) { kc = 0; ++row; }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (c0 + u >= g.nchunks) break;  // warp-uniform
                u32 nib = 0;
                if (ok[u]) nib = (u32)(v[u].x > thr) | ((u32)(v[u].y > thr) << 1) | ((u32)(v[u].z > thr) << 2) |
                                 ((u32)(v[u].w > thr) << 3);
                // OR over each group of 8 lanes (a sub-mask redux.sync compiles to a software loop: 3 shuffles instead)
                u32 w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(SYN_FULL, w, 1);
                w |= __shfl_xor_sync(SYN_FULL, w, 2);
                w |= __shfl_xor_sync(SYN_FULL, w, 4);
                // ok[] of the first lane of a group says whether the word exists
                if ((lane & 7) == 0 && ok[u]) {
                    mask[widx[u]] = w;
                    nfg += __popc(w);
                }
            }
        } else {
#pragma unroll 1
            for (int u = 0; u < UNROLL; ++u) {
                u32 c = c0 + u;
                if (c >= g.nchunks) break;
                u32 row = c / g.cpr;
                u32 k4 = (c - row * g.cpr) * 4;
                const float *p = pred + (i64)row * g.nz;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    i64 z = (i64)(k4 + j) * 32 + lane;
                    bool b = false;
                    if (z < g.nz) b = __ldcs(p + z) > thr;
                    u32 w = __ballot_sync(SYN_FULL, b);
                    if (lane == j && k4 + j < g.wz) {
                        mask[row * g.wz + k4 + j] = w;
                        nfg += __popc(w);
                    }
                }
            }
        }
    }
    // voxel count above threshold (diagnostic + fg fraction)
    for (int d = 16; d; d >>= 1) nfg += __shfl_down_sync(SYN_FULL, nfg, d);
    if (lane == 0 && nfg) atomicAdd(fg_count, (u64)nfg);
}

// ===========================================================================
// Run-id prefix.  runbase[w] = number of z-runs that start before mask word w.
// Two representations:
//   global : rb[w] holds the value itself (generic scan, k_scan_1p<RunStartSrc>), tb == nullptr
//   tiled  : rb[w] holds the prefix INSIDE the word's stream tile (TS_CHUNKS chunks) and tb[tile] the
//            number of runs before the tile (k_stream; its last CTA scans the tile counts: no look-back
//            chain in the streaming kernel, no fix-up pass over the words)
// ===========================================================================
constexpr int TS_THREADS = 256;
constexpr int TS_CHUNKS = 128;              // 128 chunks = 16384 voxels = 64 KB of float32 per tile
constexpr int TS_SLOTS = TS_CHUNKS * 4;     // word slots per tile (a slot past the end of a row holds no word)

struct RunIdx {
    const u32 *rb;
    const u32 *tb;
    u32 cpr;
    // multi-label variant (conncomps.py:24-28): runs also break where the base id changes, so the run-start
    // words are stored (startw) instead of derived from the mask, and runs of neighbouring rows are joined
    // only when their base ids (read from mlbase) are equal.  Both null on the plain path.
    const u32 *startw;
    const u64 *mlbase;
    __device__ __forceinline__ u32 at(u32 w, u32 row, u32 k) const
    {
        u32 v = rb[w];
        if (tb) v += tb[(row * cpr + (k >> 2)) / TS_CHUNKS];
        return v;
    }
    // run-start bits of word w (its in-row index is k)
    __device__ __forceinline__ u32 starts(const u32 *mask, u32 w, u32 k) const
    {
        if (startw) return startw[w];
        return run_starts(mask[w], k ? (mask[w - 1] >> 31) : 0u);
    }
};

// a row whose unions with already-visited neighbour rows cross a tile border of k_tile_ccl
__device__ __forceinline__ bool tile_border_row(const TileGeom &tg, u32 x, u32 y, u32 ny)
{
    if (!tg.tx) return false;
    if ((x && (x & (tg.tx - 1)) == 0) || (y && (y & (tg.ty - 1)) == 0)) return true;
    return tg.diag && x && y + 1 < ny && ((y + 1) & (tg.ty - 1)) == 0;      // (x-1, y+1) lies in the next y-tile
}

// ===========================================================================
// K1 (sector mode, nz % 8 == 0): ONE streaming pass over the dense inputs, at HBM speed:
//   pred (4 B/voxel read)   -> mask[w]    the foreground bits
//   base (8 B/voxel read)   -> max(base)  (shape of the overlap matrix)
//   labels (4 B/voxel write)   zero-fill of every 32-byte label sector WITHOUT foreground; the sectors
//                              that hold foreground are written whole by k_sector_labels later
// and, FUSE_SCAN (no dilation), per tile of TS_CHUNKS chunks, from shared memory:
//   rb[w]       run starts before word w INSIDE the tile;  tileagg[tile] = runs of the tile
//   actw[]      compact list of the non-empty words (tile order by one atomic per tile)
//   bndw[]      non-empty words on a tile border of k_tile_ccl (the input of k_union_bnd)
// No dependence between CTAs: 16 B/voxel of pure streaming.
// ===========================================================================
// TMA_BASE / TMA_PRED (need nz % 128 == 0, i.e. chunk c starts at voxel 128 c): the stream does not pass through
// registers.  Thread 0 issues one cp.async.bulk per quarter tile and array (32 KB of base ids, 16 KB of predictions)
// into a ring of NST shared-memory stages, NST - 1 quarter tiles ahead (also across the scan phase and the tile
// boundary); the CTA consumes a stage from shared memory when its mbarrier flips.  The bytes in flight then live
// in shared memory, which is what this kernel is short of (1 / 2 / 3 CTAs per SM = 585 / 406 /
// 364 us with register loads).  Base ids -- two thirds of the bytes read, only needed for a max() -- gain most.
constexpr int TS_QCHUNKS = TS_CHUNKS / 4;                      // chunks per quarter tile (one bulk copy per array)
constexpr int TS_PRED_BYTES = TS_QCHUNKS * 128 * 4;            // 16 KB of predictions
constexpr int TS_BASE_BYTES = TS_QCHUNKS * 128 * 8;            // 32 KB of base ids
__host__ __device__ constexpr size_t ts_stage_bytes(bool tma_pred, bool tma_base)
{
    return (tma_pred ? TS_PRED_BYTES : 0) + (tma_base ? TS_BASE_BYTES : 0);
}

template <bool HAS_BASE, bool FUSE_SCAN, bool TMA_BASE, bool TMA_PRED, int NST>
__global__ void __launch_bounds__(TS_THREADS) k_stream(const float *__restrict__ pred, const u64 *__restrict__ base,
                                                       u32 *__restrict__ labels, u32 *__restrict__ mask,
                                                       u32 *__restrict__ rb, u32 *__restrict__ tileagg,
                                                       u32 *__restrict__ tilebase, u32 *__restrict__ actw,
                                                       u32 *__restrict__ bndw, Geom g, TileGeom tg, float thr,
                                                       Counts *cnt, u32 max_runs)
{
    constexpr bool TMA = TMA_BASE || TMA_PRED;
    constexpr size_t STAGE = ts_stage_bytes(TMA_PRED, TMA_BASE);
    constexpr size_t BASE_OFF = TMA_PRED ? TS_PRED_BYTES : 0;    // stage layout: [predictions][base ids]
    extern __shared__ __align__(128) unsigned char s_stage[];    // TMA: NST stages
    __shared__ __align__(8) u64 s_mbar[NST];
    __shared__ u32 s_m[TS_SLOTS + 1];   // s_m[0]: the word before the tile's first slot; slot s at s_m[s + 1]
    __shared__ u32 sm[33];
    __shared__ u32 s_bcount, s_abase, s_bbase;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 ntiles = (g.nchunks + TS_CHUNKS - 1) / TS_CHUNKS;
    u32 nfg = 0;
    u64 bmax = 0;
    u32 qi = 0;                          // this CTA's quarter-tile counter (TMA)
    // quarter tile #q of this CTA: tile blockIdx.x + (q / 4) * gridDim.x, chunks [.. + (q % 4) * 32, + 32)
    auto issue = [&](u32 q) {            // thread 0 only
        const u64 tile = (u64)blockIdx.x + (u64)(q >> 2) * gridDim.x;
        if (tile >= ntiles) return;
        const u64 chunk = tile * TS_CHUNKS + (u64)(q & 3u) * TS_QCHUNKS;
        if (chunk >= g.nchunks) return;
        const u32 nch = (u32)min((u64)TS_QCHUNKS, (u64)g.nchunks - chunk);
        const u32 stage = q % NST;
        unsigned char *dst = s_stage + (size_t)stage * STAGE;
        mbar_expect_tx(&s_mbar[stage], nch * ((TMA_PRED ? 512u : 0u) + (TMA_BASE ? 1024u : 0u)));
        if (TMA_PRED) bulk_g2s(dst, pred + chunk * 128, nch * 512u, &s_mbar[stage]);
        if (TMA_BASE) bulk_g2s(dst + BASE_OFF, base + chunk * 128, nch * 1024u, &s_mbar[stage]);
    };
    if (TMA) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < NST; ++s) mbar_init(&s_mbar[s], 1);
            mbar_init_fence();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (u32 q = 0; q + 1 < (u32)NST; ++q) issue(q);
    }
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u32 cbase = tile * TS_CHUNKS;
        if (FUSE_SCAN && threadIdx.x == 0) s_bcount = 0;
        // ---- phase 1: the tile's 128 chunks, 16 per warp, 4 chunks (4 x 48 B per lane) in flight ----
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            // warp w handles chunks it * 32 + w * 4 .. + 3 of the tile (a quarter tile per iteration and CTA)
            const u32 cl0 = (u32)it * TS_QCHUNKS + (u32)wid * 4;      // chunk index inside the tile
            const u32 c0 = cbase + cl0;
            if (TMA && threadIdx.x == 0) issue(qi + NST - 1);
            const unsigned char *stg = s_stage + (size_t)(qi % NST) * STAGE;
            const bool qvalid = (u64)cbase + (u64)it * TS_QCHUNKS < g.nchunks;
            if (TMA_PRED && qvalid) mbar_wait(&s_mbar[qi % NST], (qi / NST) & 1u);     // predictions of this quarter tile
            float4 v[4];
            ulonglong2 b01[4], b23[4];
            bool ok[4];
            u32 widx[4];
            i64 vox[4];
            u32 row = c0 / g.cpr;
            u32 kc = c0 - row * g.cpr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const i64 z = (i64)kc * 128 + 4 * lane;
                ok[u] = (c0 + u < g.nchunks) && z < g.nz;
                widx[u] = row * g.wz + kc * 4 + (lane >> 3);
                vox[u] = (i64)row * g.nz + z;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS_BASE && !TMA_BASE) { b01[u] = make_ulonglong2(0, 0); b23[u] = make_ulonglong2(0, 0); }
                if (ok[u]) {
                    if (TMA_PRED) v[u] = *reinterpret_cast<const float4 *>(stg + ((size_t)(wid * 4 + u) * 128 + 4 * lane) * 4);
                    else v[u] = __ldcs(reinterpret_cast<const float4 *>(pred + vox[u]));
                    if (HAS_BASE && !TMA_BASE) {
                        const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(base + vox[u]);
                        b01[u] = __ldcs(bp);
                        b23[u] = __ldcs(bp + 1);
                    }
                }
                if (++kc == g.cpr) { kc = 0; ++row; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                u32 nib = 0;
                if (ok[u]) nib = (u32)(v[u].x > thr) | ((u32)(v[u].y > thr) << 1) | ((u32)(v[u].z > thr) << 2) |
                                 ((u32)(v[u].w > thr) << 3);
                u32 w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(SYN_FULL, w, 1);
                // after the first step w holds exactly this lane pair's sector (8 voxels = 32 B of labels)
                if (ok[u] && w == 0) __stcs(reinterpret_cast<uint4 *>(labels + vox[u]), make_uint4(0, 0, 0, 0));
                w |= __shfl_xor_sync(SYN_FULL, w, 2);
                w |= __shfl_xor_sync(SYN_FULL, w, 4);
                if ((lane & 7) == 0) {
                    if (FUSE_SCAN) s_m[1 + (cl0 + u) * 4 + (lane >> 3)] = w;      // 0 for a slot that holds no word
                    if (ok[u]) {
                        mask[widx[u]] = w;
                        nfg += __popc(w);
                    }
                }
                if (HAS_BASE && !TMA_BASE) bmax = max(bmax, max(max(b01[u].x, b01[u].y), max(b23[u].x, b23[u].y)));
            }
            if (TMA) {
                if (TMA_BASE && qvalid) {
                    if (!TMA_PRED) mbar_wait(&s_mbar[qi % NST], (qi / NST) & 1u);
                    const u64 chunk = (u64)cbase + (u64)it * TS_QCHUNKS;
                    const u32 n16 = (u32)min((u64)TS_QCHUNKS, (u64)g.nchunks - chunk) * 64u;   // 16-byte units
                    const ulonglong2 *sp = reinterpret_cast<const ulonglong2 *>(stg + BASE_OFF);
#pragma unroll 4
                    for (u32 j = threadIdx.x; j < n16; j += TS_THREADS) {
                        const ulonglong2 b = sp[j];
                        bmax = max(bmax, max(b.x, b.y));
                    }
                }
                __syncthreads();     // every thread is done with this stage: it is refilled one iteration later
                ++qi;
            }
        }
        if (!FUSE_SCAN) continue;
        if (threadIdx.x == 0) {
            // top bit of the word before the tile's first slot (same row), recomputed from the prediction
            const u32 row0 = cbase / g.cpr;
            const u32 kc0 = cbase - row0 * g.cpr;
            u32 prev = 0;
            if (kc0) prev = (__ldg(pred + (i64)row0 * g.nz + (i64)kc0 * 128 - 1) > thr) ? 0x80000000u : 0u;
            s_m[0] = prev;
        }
        __syncthreads();
        // ---- phase 2: thread t owns slots 2t, 2t+1: run starts, active flag, tile-border list position ----
        u32 wi[2], st[2], bpos[2];
        bool act[2], exists[2];
        u32 v2 = 0;
        {
            const u32 c = cbase + (threadIdx.x >> 1);
            const bool cok = c < g.nchunks;
            const u32 row = cok ? c / g.cpr : 0;
            const u32 kc = cok ? c - row * g.cpr : 0;
            u32 x, y;
            split_row(g, row, x, y);
            const bool border = tile_border_row(tg, x, y, (u32)g.ny);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const u32 s = 2 * threadIdx.x + q;
                const u32 k = kc * 4 + (s & 3);
                const u32 m = s_m[s + 1];                       // 0 for a slot that holds no word
                const u32 carry = k ? (s_m[s] >> 31) : 0u;
                exists[q] = cok && k < g.wz;
                st[q] = __popc(run_starts(m, carry));
                act[q] = m != 0;
                wi[q] = row * g.wz + k;
                v2 += st[q] + ((u32)act[q] << 16);
                bpos[q] = 0xFFFFFFFFu;
                if (act[q] && border) bpos[q] = atomicAdd(&s_bcount, 1u);
            }
        }
        u32 total;
        const u32 ex = block_excl_scan(v2, sm, &total);
        if (threadIdx.x == 0) {
            tileagg[tile] = total & 0xFFFFu;
            const u32 na = total >> 16, nb = s_bcount;
            s_abase = na ? (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_ACTIVE_WORDS], (u64)na) : 0u;
            s_bbase = nb ? (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_BND_WORDS], (u64)nb) : 0u;
        }
        __syncthreads();
        // ---- phase 3: outputs ----
        {
            u32 r = ex & 0xFFFFu;
            u32 ab = s_abase + (ex >> 16);
            const u32 bb = s_bbase;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (exists[q]) rb[wi[q]] = r;
                r += st[q];
                if (act[q]) actw[ab++] = wi[q];
                if (bpos[q] != 0xFFFFFFFFu) bndw[bb + bpos[q]] = wi[q];
            }
        }
        __syncthreads();   // s_m, s_bcount, s_abase are rewritten in the next round
    }
    for (int d = 16; d; d >>= 1) nfg += __shfl_down_sync(SYN_FULL, nfg, d);
    if (lane == 0 && nfg) atomicAdd((u64 *)&cnt->v[SYN_CNT_FG_VOXELS], (u64)nfg);
    if (HAS_BASE) {
        for (int d = 16; d; d >>= 1) bmax = max(bmax, __shfl_down_sync(SYN_FULL, bmax, d));
        if (lane == 0 && bmax) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_BASE], bmax);
    }
    if (!FUSE_SCAN) return;
    // ---- the CTA that finishes last turns the tiles' run counts into run-id bases (exclusive scan of a few
    // thousand values), publishes the total run count and checks it against the capacity of the per-run arrays ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_abase = (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_TICKET], 1ull);
    __syncthreads();
    if (s_abase != gridDim.x - 1) return;
    __threadfence();
    u64 carry = 0;
    for (u32 t0 = 0; t0 < ntiles; t0 += TS_THREADS * 8) {
        const u32 i0 = t0 + threadIdx.x * 8;
        u32 a[8];
        u32 v = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { a[q] = (i0 + q < ntiles) ? __ldcg(tileagg + i0 + q) : 0u; v += a[q]; }
        u32 total;
        u32 ex = block_excl_scan(v, sm, &total);
        if (carry + total > 0xFFFFFFFFull) {                 // more runs than 32 bits: reported through the count
            carry += total;
            break;
        }
        ex += (u32)carry;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (i0 + q < ntiles) tilebase[i0 + q] = ex;
            ex += a[q];
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        cnt->v[SYN_CNT_RUNS] = (i64)carry;
        if (carry > (u64)max_runs) atomicOr((u64 *)&cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_RUNS);
    }
}

// ===========================================================================
// K1 of the multi-label variant (conncomps.py:24-28: temp[mask] = overlap_seg[mask]; label temp):
// foreground = pred > thr AND base != 0; a z-run also ends where the base id changes.
// Writes the mask word and the run-start word of every 32 voxels (one warp per word).
// ===========================================================================
__global__ void __launch_bounds__(256) k_threshold_ml(const float *__restrict__ pred, const u64 *__restrict__ base,
                                                      u32 *__restrict__ mask, u32 *__restrict__ startw, Geom g,
                                                      float thr, u64 *__restrict__ fg_count)
{
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u32 nfg = 0;
    for (u32 w = warp; w < g.nwords; w += nwarps) {
        u32 row, k;
        split_word(g, w, row, k);
        const i64 z = (i64)k * 32 + lane;
        bool f = false;
        u64 b = 0;
        if (z < g.nz) {
            const i64 v = (i64)row * g.nz + z;
            b = __ldcs(base + v);
            f = (__ldcs(pred + v) > thr) && b != 0;
        }
        // predecessor along z: lane - 1, or (lane 0) the last voxel of the previous word of the row
        u64 pb = __shfl_up_sync(SYN_FULL, b, 1);
        bool pf = __shfl_up_sync(SYN_FULL, (int)f, 1) != 0;
        if (lane == 0) {
            pf = false;
            if (k > 0) {
                const i64 v = (i64)row * g.nz + z - 1;
                pb = base[v];
                pf = (pred[v] > thr) && pb != 0;
            }
        }
        const bool st = f && !(pf && pb == b);
        const u32 mw = __ballot_sync(SYN_FULL, f);
        const u32 sw = __ballot_sync(SYN_FULL, st);
        if (lane == 0) {
            mask[w] = mw;
            startw[w] = sw;
            nfg += __popc(mw);
        }
    }
    if (lane == 0 && nfg) atomicAdd(fg_count, (u64)nfg);
}

// ===========================================================================
// K1d: 2-D L1 dilation by k over axes (y, z) on the bitmask -- what the chamfer
// distance transform + `dist > k -> 0` of _seg_utils.cpp:25-82,180-200 computes
// for a 0/1 mask.  The L1 ball of radius k is the k-fold Minkowski sum of the
// radius-1 diamond, and inside a rectangle every L1-shortest lattice path can be
// chosen inside it, so k passes of the 5-point step below give the clipped
// dilation exactly.  One thread per output word; the bitmask is L2 resident.
// ===========================================================================
__global__ void __launch_bounds__(256) k_dilate_step(const u32 *__restrict__ in, u32 *__restrict__ out, Geom g)
{
    const u32 tail = (g.nz & 31) ? ((1u << (g.nz & 31)) - 1u) : 0xFFFFFFFFu;
    for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < g.nwords; w += gridDim.x * blockDim.x) {
        u32 row, kw, x, y;
        split_word(g, w, row, kw);
        split_row(g, row, x, y);
        const u32 cur = in[w];
        const u32 prev = kw > 0 ? in[w - 1] : 0u;
        const u32 next = kw + 1 < g.wz ? in[w + 1] : 0u;
        u32 acc = cur | (cur << 1) | (prev >> 31) | (cur >> 1) | (next << 31);
        if (y > 0) acc |= in[w - g.wz];
        if (y + 1 < (u32)g.ny) acc |= in[w + g.wz];
        if (kw == g.wz - 1) acc &= tail;
        out[w] = acc;
    }
}

// ===========================================================================
// Generic single-pass exclusive scan of a per-item count (decoupled look-back).
// Tiles of SCAN_TILE items are handed out in order through an atomic ticket, so a
// block that waits on a predecessor's prefix can only wait on a block that is
// already running.  state[tile] packs (flag << 32 | value): flag 1 = the tile's own
// sum, flag 2 = inclusive prefix up to and including the tile.  `state` and `ticket`
// must be zero before the launch.  A thread owns SCAN_ITEMS CONSECUTIVE items:
// `Src` provides n() (item count, may live on the device) and load(i0, n, item[]);
// `Dst` receives store(i0, n, exclusive prefix of item 0, item[]).
// ===========================================================================
__device__ __forceinline__ u32 lookback_u32(volatile u64 *state, u32 tile, u32 total, int lane)
{
    // warp-wide look-back: lane l inspects tile (p - l); tiles before 0 count as a zero prefix
    u32 prefix = 0;
    if (tile != 0) {
        if (lane == 0) state[tile] = (1ull << 32) | total;
        i64 p = (i64)tile - 1;
        while (true) {
            const i64 idx = p - lane;
            const u64 sv = idx >= 0 ? state[idx] : (2ull << 32);
            const u32 flag = (u32)(sv >> 32);
            const u32 m2 = __ballot_sync(SYN_FULL, flag == 2);
            const u32 m0 = __ballot_sync(SYN_FULL, flag == 0);
            const u32 upto = m2 ? mask_le((u32)(__ffs(m2) - 1)) : 0xFFFFFFFFu;   // lanes that count
            if (m0 & upto) continue;                                              // someone not published yet
            u32 val = ((upto >> lane) & 1u) ? (u32)sv : 0u;
#pragma unroll
            for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(SYN_FULL, val, d);
            prefix += val;
            if (m2) break;
            p -= 32;
        }
    }
    if (lane == 0) state[tile] = (2ull << 32) | (u64)(prefix + total);
    return prefix;
}

template <class Src, class Dst, int ITEMS>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_1p(Src src, Dst dst, volatile u64 *state, u32 *ticket,
                                                          i64 *total_out)
{
    constexpr int TILE = SCAN_THREADS * ITEMS;
    __shared__ u32 sm[33];
    __shared__ u32 s_tile, s_prefix;
    const u64 n = src.n();
    const u32 ntiles = (u32)((n + TILE - 1) / TILE);
    if (ntiles == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
        return;
    }
    // When the grid covers all tiles, the CTAs beyond the (device-side) tile count leave at once instead of queueing
    // on the ticket only to find out that there is no work: hundreds of same-address atomics were a fixed ~5 us per
    // scan.  The tickets themselves stay: a CTA then only ever waits on CTAs that are already running, whatever else
    // shares the GPU (a second pipeline on another stream, for instance).
    const bool one_each = gridDim.x >= ntiles;
    if (one_each && blockIdx.x >= ntiles) return;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;  // block-uniform
        u32 item[ITEMS];
        const u64 i0 = (u64)tile * TILE + (u64)threadIdx.x * ITEMS;
        src.load(i0, n, item);
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) v += item[t];
        u32 total;
        u32 ex = block_excl_scan(v, sm, &total);
        if (threadIdx.x < 32) {
            const u32 prefix = lookback_u32(state, tile, total, threadIdx.x);
            if (threadIdx.x == 0) {
                s_prefix = prefix;
                if (tile == ntiles - 1 && total_out) *total_out = (i64)(prefix + total);
            }
        }
        __syncthreads();
        ex += s_prefix;
        dst.store(i0, n, ex, item);
        if (one_each) break;               // one tile per CTA: no second ticket round trip
        __syncthreads();   // s_tile / s_prefix are rewritten in the next round
    }
}

// N consecutive u32 of a 256-byte aligned array (i0 is a multiple of N); zero past n
template <int N>
__device__ __forceinline__ void load_items_u32(const u32 *a, u64 i0, u64 n, u32 (&m)[N])
{
    if (N % 4 == 0 && i0 + N <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(a + i0);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const uint4 v = __ldcg(p + q);
            m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int t = 0; t < N; ++t) m[t] = (i0 + t < n) ? __ldcg(a + i0 + t) : 0u;
    }
}
template <int N>
__device__ __forceinline__ void store_prefix_u32(u32 *a, u64 i0, u64 n, u32 ex, const u32 (&item)[N])
{
    u32 pre[N];
#pragma unroll
    for (int t = 0; t < N; ++t) { pre[t] = ex; ex += item[t]; }
    if (N % 4 == 0 && i0 + N <= n) {
        uint4 *p = reinterpret_cast<uint4 *>(a + i0);
#pragma unroll
        for (int q = 0; q < N / 4; ++q)
            p[q] = make_uint4(pre[4 * q], pre[4 * q + 1], pre[4 * q + 2], pre[4 * q + 3]);
    } else {
#pragma unroll
        for (int t = 0; t < N; ++t)
            if (i0 + t < n) a[i0 + t] = pre[t];
    }
}

// --- scan #1: run starts per mask word -> runbase[w]  (dilation variant only; the plain path fuses this
// scan into the threshold kernel, k_threshold_runs) ------------------------
struct RunStartSrc {
    const u32 *mask;
    u32 nwords, wz;
    const u32 *startw;     // multi-label variant: stored run-start words (else null)
    __device__ __forceinline__ u64 n() const { return nwords; }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
        if (startw) {
            u32 m[N];
            load_items_u32(startw, i0, n, m);
#pragma unroll
            for (int t = 0; t < N; ++t) item[t] = __popc(m[t]);
            return;
        }
        if (i0 >= n) {
#pragma unroll
            for (int t = 0; t < N; ++t) item[t] = 0;
            return;
        }
        u32 m[N];
        load_items_u32(mask, i0, n, m);
        u32 prev = i0 ? __ldcg(mask + i0 - 1) : 0u;
        u32 k = (u32)(i0 % wz);
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const u32 carry = k ? (prev >> 31) : 0u;
            item[t] = __popc(run_starts(m[t], carry));
            prev = m[t];
            if (++k == wz) k = 0;
        }
    }
};
struct RunBaseDst {
    u32 *runbase;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
        store_prefix_u32(runbase, i0, n, ex, item);
    }
};


// ---------------------------------------------------------------------------
// Compaction of the non-empty mask words.  The sparse kernels (unions, statistics,
// foreground sectors) run one thread per ACTIVE word from this list, so every lane
// of a warp has work (3 % foreground means ~10 % of the words are non-empty; a
// thread-per-word grid over all words would run at ~2 active lanes per warp).
// Order inside a block is raster order; blocks append with one atomicAdd each.  Also lists the non-empty words
// on a tile border (the dilation variant; the plain path gets both lists from k_stream).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_compact_active(const u32 *__restrict__ mask, u32 nwords,
                                                                 u32 *__restrict__ actw, u32 *__restrict__ bndw, Geom g,
                                                                 TileGeom tg, Counts *cnt)
{
    __shared__ u32 sm[33];
    __shared__ u32 s_base, s_bcount, s_bbase;
    for (u64 tile = blockIdx.x; tile * SCAN_TILE < nwords; tile += gridDim.x) {
        if (threadIdx.x == 0) s_bcount = 0;
        __syncthreads();
        const u64 i0 = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
        u32 flag[SCAN_ITEMS], bpos[SCAN_ITEMS];
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t) {
            flag[t] = (i0 + t < nwords) && (mask[i0 + t] != 0);
            v += flag[t];
            bpos[t] = 0xFFFFFFFFu;
            if (flag[t] && tg.tx) {      // non-empty word on a tile border of k_tile_ccl -> input of k_union_bnd
                u32 row, k, x, y;
                split_word(g, (u32)(i0 + t), row, k);
                split_row(g, row, x, y);
                if (tile_border_row(tg, x, y, (u32)g.ny)) bpos[t] = atomicAdd(&s_bcount, 1u);
            }
        }
        u32 total;
        u32 ex = block_excl_scan(v, sm, &total);
        if (threadIdx.x == 0) {
            s_base = total ? (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_ACTIVE_WORDS], (u64)total) : 0u;
            const u32 nb = s_bcount;
            s_bbase = nb ? (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_BND_WORDS], (u64)nb) : 0u;
        }
        __syncthreads();
        ex += s_base;
        const u32 bb = s_bbase;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t) {
            if (flag[t]) actw[ex++] = (u32)(i0 + t);
            if (bpos[t] != 0xFFFFFFFFu) bndw[bb + bpos[t]] = (u32)(i0 + t);
        }
        __syncthreads();
    }
}

// validates the run count against capacity and initialises parent[g] = g
__global__ void k_init_parents(u32 *__restrict__ par, Counts *cnt, u32 max_runs)
{
    const u64 nruns = (u64)cnt->v[SYN_CNT_RUNS];
    if (nruns > max_runs) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr((u64 *)&cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_RUNS);
        return;
    }
    for (u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x; g < nruns; g += (u64)gridDim.x * blockDim.x)
        par[g] = (u32)g;
}

// ===========================================================================
// K3: unions between z-runs of neighbouring rows.  One thread per mask word of
// the "current" row; neighbour rows are the already-visited ones in raster order
// ((x, y-1), (x-1, y) for 6-connectivity; plus (x-1, y-1), (x-1, y+1) and the
// +-1 z-shifts for 18 / 26).  Two runs are joined when their z-intervals meet
// (after the shift).  Run id of the run covering bit b of word w:
//     runbase[w] + popc(starts(w) & mask_le(b)) - 1
// ===========================================================================

struct RowView {
    const u32 *mask;     // words of this row
    RunIdx R;            // run-start prefix
    u32 row, wz;
    i64 nz;
    __device__ __forceinline__ u32 word(int k) const { return (k >= 0 && k < (int)wz) ? mask[k] : 0u; }
    // id of the run covering bit b (0..31) of word k; that bit must be set
    __device__ __forceinline__ u32 run_at(int k, int b) const
    {
        const u32 w = row * wz + (u32)k;
        return R.at(w, row, (u32)k) + __popc(R.starts(mask - (size_t)row * wz, w, (u32)k) & mask_le((u32)b)) - 1u;
    }
};

// joins runs of row A (word k, bits `mine`) with runs of row B whose bit (b + shift) is set
template <int SHIFT>
__device__ __forceinline__ void union_shifted(u32 *par, const RowView &A, const RowView &B, int k, u32 mA)
{
    u32 nb;  // bit b set <=> B has bit (b + SHIFT) set
    if (SHIFT == 0) nb = B.word(k);
    else if (SHIFT < 0) nb = (B.word(k) << 1) | (B.word(k - 1) >> 31);
    else nb = (B.word(k) >> 1) | (B.word(k + 1) << 31);
    u32 o = mA & nb;
    if (!o) return;
    u32 os = o & ~(o << 1);  // first bit of every overlap span (a span lies inside one run of A and one of B)
    const bool ml = SHIFT == 0 && A.R.startw != nullptr;
    if (ml)   // spans also break at the run starts of either row (base-id changes)
        os |= o & (A.R.startw[A.row * A.wz + (u32)k] | A.R.startw[B.row * B.wz + (u32)k]);
    while (os) {
        int b = __ffs(os) - 1;
        os &= os - 1;
        if (ml) {
            const i64 z = (i64)k * 32 + b;
            if (A.R.mlbase[(i64)A.row * A.nz + z] != A.R.mlbase[(i64)B.row * A.nz + z]) continue;
        }
        u32 ga = A.run_at(k, b);
        int bb = b + SHIFT, kb = k;
        if (bb < 0) { bb = 31; kb = k - 1; }
        else if (bb > 31) { bb = 0; kb = k + 1; }
        u32 gb = B.run_at(kb, bb);
        uf_union(par, ga, gb);
    }
}

// With tiling enabled (tg.tx != 0) a pair of rows inside one tile was already joined by K2 and is
// skipped here unless that tile was flagged as too dense.
// MODE 0: every active word, borders and flagged tiles (no border list: non-tiled geometry).
// MODE 1: the words of the border list, borders only.
// MODE 2: every active word, flagged tiles only (tiles the shared-memory pass had to leave; usually none).
template <int CONN, int MODE>
__device__ __forceinline__ void union_rows_body(const u32 *__restrict__ mask, const RunIdx &R, u32 *par, const Geom &g,
                                                const TileGeom &tg, const u32 *__restrict__ tileflag,
                                                const u32 *__restrict__ words, const Counts *cnt)
{
    if (MODE == 2 && cnt->v[SYN_CNT_FLAGGED_TILES] == 0) return;
    const u32 nact = (u32)cnt->v[MODE == 1 ? SYN_CNT_BND_WORDS : SYN_CNT_ACTIVE_WORDS];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nact; i += gridDim.x * blockDim.x) {
        const u32 w = words[i];
        u32 m = mask[w];
        u32 row, kk, x, y;
        split_word(g, w, row, kk);
        split_row(g, row, x, y);
        const int k = (int)kk;
        RowView A{mask + (u64)row * g.wz, R, row, g.wz, g.nz};
        bool do_y = true, do_x = true, do_dm = true, do_dp = true;   // (x,y-1) (x-1,y) (x-1,y-1) (x-1,y+1)
        if (tg.tx) {
            const u32 tX = x >> tg.txs, tY = y >> tg.tys;
            const bool dense = tileflag[tX * tg.tiles_y + tY] != 0;
            const bool xb = (x & (tg.tx - 1)) == 0, yb = (y & (tg.ty - 1)) == 0, ye = ((y + 1) & (tg.ty - 1)) == 0;
            // a flagged tile's own rows are all done in MODE 0 / 2; its border rows need nothing extra in MODE 1
            const bool in = MODE != 1 && dense, bd = MODE != 2;
            do_y = in || (bd && yb);
            do_x = in || (bd && xb);
            do_dm = in || (bd && (xb || yb));
            do_dp = in || (bd && (xb || ye));
            if (!(do_y || do_x || do_dm || do_dp)) continue;
        }
        if (y > 0 && do_y) {
            u32 r2 = row - 1;
            RowView B{mask + (u64)r2 * g.wz, R, r2, g.wz, g.nz};
            union_shifted<0>(par, A, B, k, m);
            if (CONN >= 18) { union_shifted<-1>(par, A, B, k, m); union_shifted<1>(par, A, B, k, m); }
        }
        if (x > 0) {
            u32 r2 = row - (u32)g.ny;
            RowView B{mask + (u64)r2 * g.wz, R, r2, g.wz, g.nz};
            if (do_x) {
                union_shifted<0>(par, A, B, k, m);
                if (CONN >= 18) { union_shifted<-1>(par, A, B, k, m); union_shifted<1>(par, A, B, k, m); }
            }
            if (CONN >= 18 && y > 0 && do_dm) {
                u32 r3 = r2 - 1;
                RowView C{mask + (u64)r3 * g.wz, R, r3, g.wz, g.nz};
                union_shifted<0>(par, A, C, k, m);
                if (CONN == 26) { union_shifted<-1>(par, A, C, k, m); union_shifted<1>(par, A, C, k, m); }
            }
            if (CONN >= 18 && y + 1 < (u32)g.ny && do_dp) {
                u32 r3 = r2 + 1;
                RowView C{mask + (u64)r3 * g.wz, R, r3, g.wz, g.nz};
                union_shifted<0>(par, A, C, k, m);
                if (CONN == 26) { union_shifted<-1>(par, A, C, k, m); union_shifted<1>(par, A, C, k, m); }
            }
        }
    }
}

template <int CONN>
__global__ void __launch_bounds__(256) k_union_rows(const u32 *__restrict__ mask, RunIdx R, u32 *par, Geom g, TileGeom tg,
                                                    const u32 *__restrict__ tileflag, const u32 *__restrict__ actw,
                                                    const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    union_rows_body<CONN, 0>(mask, R, par, g, tg, tileflag, actw, cnt);
}


// ===========================================================================
// K2: block-local union-find in shared memory.  A tile = TX x-slabs x TY rows x
// the full z extent (<= TILE_WORDS mask words); one CTA per tile loads the tile's
// mask words and run-id prefixes into shared memory, joins the z-runs of
// neighbouring rows INSIDE the tile with atomicMin on a shared parent array
// (local run ids keep raster order, so local roots are raster-first too), and
// writes parent[run] = global id of the local root.  Only pairs of rows that
// straddle a tile border are left for the global pass (K3).  A tile with more
// than TILE_RUNS runs is flagged and handled entirely by K3.
// ===========================================================================
#ifndef SYN_TILE_WORDS
#define SYN_TILE_WORDS 2048
#endif
constexpr int TILE_WORDS = SYN_TILE_WORDS;      // (A/B builds: -DSYN_TILE_WORDS=1024; the rest scales with it)
constexpr int TILE_RUNS = 2 * TILE_WORDS;
constexpr int TILE_MAXTX = 16;
constexpr int TILE_ROOTS = TILE_WORDS / 4;    // tile-local components with statistics accumulated in shared memory
constexpr int TILE_SLOT = 10;      // u32 per root: size, sum dx, sum dy, sum z, min/max dx, dy, z
constexpr int TILE_THREADS = TILE_WORDS / 4;
constexpr int TILE_WPT = TILE_WORDS / TILE_THREADS;   // mask words per thread (consecutive)
// shared memory of one tile CTA (3 CTAs = 48 warps per SM; the kernel is instruction-issue bound, not
// occupancy bound): mask words, run-start words, packed active list, parents, statistics slots, global ids of
// the local roots, local run-id prefix, x-slab of every run
constexpr size_t TILE_AUX = (size_t)TILE_ROOTS * TILE_SLOT * 4;
constexpr size_t TILE_SMEM = (size_t)TILE_WORDS * 4 * 3 + (size_t)TILE_RUNS * 4 + TILE_AUX + (size_t)TILE_WORDS * 2 +
                             (size_t)TILE_ROOTS * 4 + (size_t)TILE_RUNS;

// Shared parent array: low 16 bits = parent (local run id < TILE_RUNS), high 16 bits = compact index of a root
// (written after the flatten).  During the unions the high half is zero, so values compare like plain ids.
__device__ __forceinline__ u32 sm_find(volatile u32 *p, u32 x)
{
    u32 q = p[x];
    while (q != x) {
        u32 gq = p[q];
        if (gq != q) p[x] = gq;   // path halving (always an ancestor, always smaller)
        x = q;
        q = gq;
    }
    return x;
}

__device__ __forceinline__ u32 sm_find_ro(const volatile u32 *p, u32 x)
{
    u32 q;
    while ((q = p[x]) != x) x = q;
    return x;
}

__device__ __forceinline__ void sm_union(u32 *p, u32 a, u32 b)
{
    while (true) {
        a = sm_find(p, a);
        b = sm_find(p, b);
        if (a == b) return;
        if (a > b) { u32 t = a; a = b; b = t; }
        u32 old = atomicMin(p + b, a);
        if (old == b) return;
        b = old;
    }
}

// position of the (n+1)-th set bit of w when n is small (the usual case: one or two spans per word)
__device__ __forceinline__ u32 select_bit_small(u32 w, u32 n)
{
    for (; n; --n) w &= w - 1;
    return (u32)__ffs(w) - 1u;
}

// s_act entry: word index (11 bits) | k << 11 (10 bits) | yl << 21 (6 bits) | xl << 27
// One pass over the tile's non-empty words for one (neighbour row, z shift) combination, ONE LANE PER
// OVERLAP SPAN: every thread first finds the overlap spans of its word with the neighbour word, a warp
// prefix sum turns the 32 counts into a work list, and the lanes then take the spans one each (owner by
// shuffle binary search, span start by select_bit) and perform the unions with uniform control flow.
// Used for 18 / 26 connectivity (6-connectivity has the combined pass below).
template <bool SHIFTED>
__device__ __forceinline__ void tile_pass_row(const u32 *s_mask, const unsigned short *s_rb, const u32 *s_st,
                                              u32 *s_par, const u32 *s_act, u32 nact, u32 wz, u32 vy, u32 dxl, int dyl,
                                              int off)
{
    const int lane = threadIdx.x & 31;
    const u32 wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    for (u32 a0 = wib * 32; a0 < nact; a0 += nwb * 32) {
        const u32 a = a0 + lane;
        // os0: starts of the direct overlap spans; os1 / os2: the diagonal contacts that are NOT implied by a direct
        // overlap -- a run of this row starting at bit b where a neighbour run ends at b - 1, and a run ending at b
        // where a neighbour run starts at b + 1 (any other z-diagonal contact means the two runs overlap)
        u32 os0 = 0, os1 = 0, os2 = 0, i = 0;
        if (a < nact) {
            const u32 e = s_act[a];
            i = e & 2047u;
            const u32 k = (e >> 11) & 1023u;
            const int yl = (int)((e >> 21) & 63u);
            const u32 xl = e >> 27;
            if (xl >= dxl && yl + dyl >= 0 && yl + dyl < (int)vy) {
                const int iB = (int)i + off;
                const u32 m = s_mask[i], cur = s_mask[iB];
                const u32 o = m & cur;
                os0 = o & ~(o << 1);
                if (SHIFTED) {
                    const bool more = k + 1 < wz;
                    const u32 shl = (cur << 1) | (k > 0 ? (s_mask[iB - 1] >> 31) : 0u);
                    const u32 shr = (cur >> 1) | (more ? (s_mask[iB + 1] << 31) : 0u);
                    const u32 ends = m & ~((m >> 1) | (more ? (s_mask[i + 1] << 31) : 0u));
                    os1 = s_st[i] & shl & ~cur;
                    os2 = ends & shr & ~cur;
                }
            }
        }
        const u32 n = __popc(os0) + __popc(os1) + __popc(os2);
        const u32 incl = warp_incl_scan(n, lane);
        const u32 excl = incl - n;
        const u32 total = __shfl_sync(SYN_FULL, incl, 31);
        for (u32 r0 = 0; r0 < total; r0 += 32) {
            const u32 j = r0 + lane;
            int o = 0;
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                const u32 e = __shfl_sync(SYN_FULL, excl, (o + s) & 31);
                if (e <= j) o += s;
            }
            const u32 e_o = __shfl_sync(SYN_FULL, excl, o);
            const u32 os0_o = __shfl_sync(SYN_FULL, os0, o);
            const u32 os1_o = SHIFTED ? __shfl_sync(SYN_FULL, os1, o) : 0u;
            const u32 os2_o = SHIFTED ? __shfl_sync(SYN_FULL, os2, o) : 0u;
            const u32 i_o = __shfl_sync(SYN_FULL, i, o);
            if (j < total) {
                u32 idx = j - e_o, os_o = os0_o;
                int shift = 0;
                if (SHIFTED) {
                    const u32 n0 = __popc(os0_o), n1 = __popc(os1_o);
                    if (idx >= n0 + n1) { idx -= n0 + n1; os_o = os2_o; shift = 1; }
                    else if (idx >= n0) { idx -= n0; os_o = os1_o; shift = -1; }
                }
                const u32 b = select_bit(os_o, idx);
                const u32 ga = (u32)s_rb[i_o] + __popc(s_st[i_o] & mask_le(b)) - 1u;
                int iB = (int)i_o + off, bb = (int)b + shift;
                if (bb < 0) { bb = 31; --iB; }
                else if (bb > 31) { bb = 0; ++iB; }
                const u32 gb = (u32)s_rb[iB] + __popc(s_st[iB] & mask_le((u32)bb)) - 1u;
                sm_union(s_par, ga, gb);
            }
        }
    }
}

// 6-connectivity: both neighbour rows ((x, y-1) and (x-1, y)) in ONE pass -- the overlap spans of a word with
// its two neighbour words form one work list, so the per-batch set-up is paid once.
__device__ __forceinline__ void tile_pass6(const u32 *s_mask, const unsigned short *s_rb, const u32 *s_st, u32 *s_par,
                                           const u32 *s_act, u32 nact, u32 wz, u32 segw)
{
    const int lane = threadIdx.x & 31;
    const u32 wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    for (u32 a0 = wib * 32; a0 < nact; a0 += nwb * 32) {   // whole batches: the kernel is issue bound, so fewer,
        const u32 a = a0 + lane;                           // fuller batches beat an even split over the warps
        u32 osy = 0, osx = 0, i = 0;
        if (a < nact) {
            const u32 e = s_act[a];
            i = e & 2047u;
            const u32 m = s_mask[i];
            if ((e >> 21) & 63u) { const u32 o = m & s_mask[i - wz]; osy = o & ~(o << 1); }
            if (e >> 27) { const u32 o = m & s_mask[i - segw]; osx = o & ~(o << 1); }
        }
        const u32 n = __popc(osy) + __popc(osx);
        const u32 incl = warp_incl_scan(n, lane);
        const u32 excl = incl - n;
        const u32 total = __shfl_sync(SYN_FULL, incl, 31);
        for (u32 r0 = 0; r0 < total; r0 += 32) {
            const u32 j = r0 + lane;
            int o = 0;
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                const u32 e = __shfl_sync(SYN_FULL, excl, (o + s) & 31);
                if (e <= j) o += s;
            }
            const u32 e_o = __shfl_sync(SYN_FULL, excl, o);
            const u32 osy_o = __shfl_sync(SYN_FULL, osy, o);
            const u32 osx_o = __shfl_sync(SYN_FULL, osx, o);
            const u32 i_o = __shfl_sync(SYN_FULL, i, o);
            if (j < total) {
                u32 idx = j - e_o;
                const u32 nyo = __popc(osy_o);
                const bool isy = idx < nyo;
                if (!isy) idx -= nyo;
                const u32 b = select_bit_small(isy ? osy_o : osx_o, idx);
                const u32 iB = i_o - (isy ? wz : segw);
                const u32 ga = (u32)s_rb[i_o] + __popc(s_st[i_o] & mask_le(b)) - 1u;
                const u32 gb = (u32)s_rb[iB] + __popc(s_st[iB] & mask_le(b)) - 1u;
                sm_union(s_par, ga, gb);
            }
        }
    }
}

// Local run ids come from a block-wide scan of the tile's own run starts (no per-word read of the global
// prefix); only the global id of the first run of each x-slab segment is fetched (<= 16 values per tile).
template <int CONN>
__global__ void __launch_bounds__(TILE_THREADS) k_tile_ccl(const u32 *__restrict__ mask, const u32 *__restrict__ omask,
                                                  RunIdx R, u32 *__restrict__ par,
                                                  u32 *__restrict__ tileflag, TileRec *__restrict__ recs, Geom g,
                                                  TileGeom tg, Counts *cnt, u32 max_runs)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u32 *s_mask = reinterpret_cast<u32 *>(s_raw);                 // [TILE_WORDS]
    u32 *s_st = s_mask + TILE_WORDS;                              // run-start bits of the non-empty words
    u32 *s_act = s_st + TILE_WORDS;                               // packed coordinates of the non-empty words
    u32 *s_par = s_act + TILE_WORDS;                              // [TILE_RUNS] parent | root index << 16
    u32 *s_aux = s_par + TILE_RUNS;
    u32 *s_slot = s_aux;                                          // [TILE_ROOTS][TILE_SLOT] statistics
    u32 *s_rootgid = reinterpret_cast<u32 *>(reinterpret_cast<unsigned char *>(s_aux) + TILE_AUX);   // [TILE_ROOTS]
    unsigned short *s_rb = reinterpret_cast<unsigned short *>(s_rootgid + TILE_ROOTS);            // LOCAL run-id prefix
    unsigned char *s_segof = reinterpret_cast<unsigned char *>(s_rb + TILE_WORDS);                // [TILE_RUNS] x-slab of a run
    __shared__ u32 sm[33];
    __shared__ u32 s_nroots, s_recbase;
    __shared__ u32 s_segoff[TILE_MAXTX + 1];       // local id of the first run of x-slab segment xl
    __shared__ u32 s_delta[TILE_MAXTX];            // global id = local id + delta[xl]
    SYN_TIMER_INIT();
    const u32 tile = blockIdx.x;
    const u32 tX = tile / tg.tiles_y, tY = tile - tX * tg.tiles_y;
    const u32 x0 = tX * tg.tx, y0 = tY * tg.ty;
    const u32 vx = min(tg.tx, (u32)g.nx - x0), vy = min(tg.ty, (u32)g.ny - y0);   // valid rows in this tile
    const u32 segw = vy * g.wz;                // words per x-slab segment (contiguous in memory)
    const u32 ntw = vx * segw;                 // words of this tile (<= TILE_WORDS)

    // ---- load: thread t owns words 4t .. 4t+3 of the tile (segment-major order); all global loads of the
    // prologue are issued before anything waits on them ----
    const u32 i0 = threadIdx.x * TILE_WPT;
    u32 m[TILE_WPT];
    u32 xl0 = 0, jj0 = 0;                      // segment / offset in segment of word i0
    if (i0 < ntw) { xl0 = i0 / segw; jj0 = i0 - xl0 * segw; }
    if ((segw & 3u) == 0 && (g.wz & 3u) == 0) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i0 < ntw) v = *reinterpret_cast<const uint4 *>(mask + ((x0 + xl0) * (u32)g.ny + y0) * g.wz + jj0);
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
    } else {
        u32 xl = xl0, jj = jj0;
#pragma unroll
        for (int q = 0; q < TILE_WPT; ++q) {
            m[q] = 0;
            if (i0 + q < ntw) {
                m[q] = mask[((x0 + xl) * (u32)g.ny + y0) * g.wz + jj];
                if (++jj == segw) { jj = 0; ++xl; }
            }
        }
    }
    // global id of the first run of every segment
    u32 segbase = 0;
    if (threadIdx.x < vx) {
        const u32 row = (x0 + threadIdx.x) * (u32)g.ny + y0;
        segbase = R.at(row * g.wz, row, 0);
    }
    const u64 errflags = (u64)cnt->v[SYN_CNT_ERRFLAGS];
    const u64 nruns_all = (u64)cnt->v[SYN_CNT_RUNS];
    if (threadIdx.x == 0) s_nroots = 0;
    *reinterpret_cast<uint4 *>(s_mask + i0) = make_uint4(m[0], m[1], m[2], m[3]);
    if (errflags & SYN_EF_RUNS) return;
    if (nruns_all > (u64)max_runs) {             // parent array too small: nothing may be written
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr((u64 *)&cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_RUNS);
        return;
    }
    __syncthreads();
    SYN_TIMER_MARK(cnt, 0);   // load
    // ---- run starts (carry from shared memory), local run ids by a block scan ----
    u32 st[TILE_WPT];                          // run-start bits per word
    u32 yl0 = 0, k0 = 0;
    if (i0 < ntw) {
        if (g.wzs < 32) { yl0 = jj0 >> g.wzs; k0 = jj0 & (g.wz - 1u); }
        else { yl0 = jj0 / g.wz; k0 = jj0 - yl0 * g.wz; }
    }
    u32 v = 0;
    {
        u32 k = k0;
#pragma unroll
        for (int q = 0; q < TILE_WPT; ++q) {
            st[q] = 0;
            if (i0 + q < ntw && m[q]) {
                st[q] = run_starts(m[q], k ? (s_mask[i0 + q - 1] >> 31) : 0u);
                v += __popc(st[q]) + (1u << 16);
            }
            if (++k == g.wz) k = 0;
        }
    }
    u32 total;
    const u32 ex = block_excl_scan(v, sm, &total);
    const u32 nloc = total & 0xFFFFu, nact = total >> 16;
    SYN_TIMER_MARK(cnt, 1);   // run starts + scan
    if (nloc == 0) return;                     // no foreground in this tile (block-uniform)
    const bool dense = nloc > TILE_RUNS;       // too dense for the shared-memory pass: K3 / k_stats do the whole tile
    {
        u32 rb = ex & 0xFFFFu, ab = ex >> 16;
        u32 jj = jj0, xl = xl0, yl = yl0, k = k0;
        const bool one_seg = (segw & 3u) == 0;           // the thread's four words lie in one x-slab segment
#pragma unroll
        for (int q = 0; q < TILE_WPT; ++q) {
            if (i0 + q < ntw) {
                if (jj == 0) s_segoff[xl] = rb;            // segment start
                const u32 n = __popc(st[q]);
                if (!dense) {
                    s_rb[i0 + q] = (unsigned short)rb;
                    if (m[q]) {
                        s_st[i0 + q] = st[q];
                        s_act[ab++] = (i0 + q) | (k << 11) | (yl << 21) | (xl << 27);
                        if (!one_seg)
                            for (u32 r = 0; r < n; ++r) { s_segof[rb + r] = (unsigned char)xl; s_par[rb + r] = rb + r; }
                    }
                }
                rb += n;
                if (++k == g.wz) { k = 0; ++yl; }
                if (++jj == segw) { jj = 0; yl = 0; ++xl; }
            }
        }
        if (!dense && one_seg) {
            // parent = self and x-slab of every run of my words: ONE loop over the thread's runs (usually 0-2)
            const u32 r0 = ex & 0xFFFFu;
            for (u32 r = r0; r < rb; ++r) { s_segof[r] = (unsigned char)xl0; s_par[r] = r; }
        }
    }
    if (threadIdx.x == 0) s_segoff[vx] = nloc;
    __syncthreads();
    if (threadIdx.x < vx) s_delta[threadIdx.x] = segbase - s_segoff[threadIdx.x];
    if (dense) {
        __syncthreads();
        // identity parents for the tile's runs (the global pass joins them), then leave
        u32 rb = ex & 0xFFFFu;
        u32 jj = jj0, xl = xl0;
#pragma unroll
        for (int q = 0; q < TILE_WPT; ++q) {
            if (i0 + q < ntw) {
                const u32 d = s_delta[xl];
                const u32 n = __popc(st[q]);
                for (u32 r = 0; r < n; ++r) par[rb + r + d] = rb + r + d;
                rb += n;
                if (++jj == segw) { jj = 0; ++xl; }
            }
        }
    }
    SYN_TIMER_MARK(cnt, 2);   // local tables
    if (dense) {
        if (threadIdx.x == 0) {
            tileflag[tile] = 1;
            atomicAdd((u64 *)&cnt->v[SYN_CNT_FLAGGED_TILES], 1ull);
        }
        return;
    }

    const int wz = (int)g.wz, sg = (int)segw;
    if (CONN == 6) {
        tile_pass6(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, segw);
    } else {
        // one pass per neighbour row: (x, y-1) and (x-1, y) with their z-diagonals, (x-1, y-1) and (x-1, y+1) with
        // (26) or without (18) theirs
        tile_pass_row<true>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 0, -1, -wz);
        tile_pass_row<true>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 0, -sg);
        tile_pass_row<CONN == 26>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, -1, -sg - wz);
        tile_pass_row<CONN == 26>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 1, -sg + wz);
    }
    __syncthreads();
    SYN_TIMER_MARK(cnt, 3);   // unions
    // flatten; write parent[global run] = global id of the local root
    for (u32 i = threadIdx.x; i < nloc; i += blockDim.x) {
        const u32 r = sm_find_ro(s_par, i);    // read-only walk; a concurrent walk through i meets the old
        s_par[i] = r;                          // ancestor or the root: both fine
        par[i + s_delta[s_segof[i]]] = r + s_delta[s_segof[r]];
    }
    __syncthreads();
    // every parent is a root now: give every local root a compact index (one shared atomic per warp), remember
    // its global id, and put the index into the high half of the root's own word (nobody walks any more)
    for (u32 i0r = 0; i0r < nloc; i0r += blockDim.x) {
        const u32 i = i0r + threadIdx.x;
        const bool isroot = i < nloc && s_par[i] == i;
        const u32 bal = __ballot_sync(SYN_FULL, isroot);
        if (bal) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(bal) - 1;
            u32 basei = 0;
            if (lane == leader) basei = atomicAdd(&s_nroots, (u32)__popc(bal));
            basei = __shfl_sync(SYN_FULL, basei, leader) + __popc(bal & mask_lt((u32)lane));
            if (isroot) {
                if (basei < TILE_ROOTS) s_rootgid[basei] = i + s_delta[s_segof[i]];
                s_par[i] = i | (basei << 16);
            }
        }
    }
    __syncthreads();
    SYN_TIMER_MARK(cnt, 4);   // flatten
    // ---- statistics of the tile-local components in shared memory (one record per local root afterwards) ----
    const u32 nroots = s_nroots;
    if (nroots > TILE_ROOTS) {                 // statistics of this tile are left to k_stats
        if (threadIdx.x == 0) {
            tileflag[tile] = 2;
            atomicAdd((u64 *)&cnt->v[SYN_CNT_FLAGGED_TILES], 1ull);
        }
        return;
    }
    for (u32 j = threadIdx.x; j < nroots * TILE_SLOT; j += blockDim.x) {
        const u32 f = j % TILE_SLOT;
        s_slot[j] = (f >= 4 && f <= 6) ? 0xFFFFFFFFu : 0u;
    }
    if (threadIdx.x == 0) s_recbase = (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_TILE_RECORDS], (u64)nroots);
    __syncthreads();
    SYN_TIMER_MARK(cnt, 5);   // slot init
    const bool same_mask = (omask == mask);
    for (u32 a = threadIdx.x; a < nact; a += blockDim.x) {
        const u32 e = s_act[a];
        const u32 i = e & 2047u, k = (e >> 11) & 1023u, yl = (e >> 21) & 63u, xl = e >> 27;
        const u32 mm = s_mask[i];
        const u32 om = same_mask ? mm : omask[((x0 + xl) * (u32)g.ny + y0 + yl) * g.wz + k];
        if (!om) continue;
        const u32 starts = s_st[i], rb = s_rb[i], z0 = k * 32;
        u32 ps = mm & ~(mm << 1);                // piece starts
        while (ps) {
            const int b0 = __ffs(ps) - 1;
            ps &= ps - 1;
            const u32 span = ~(mm >> b0);
            int len = span ? (__ffs(span) - 1) : 32;
            if (len > 32 - b0) len = 32 - b0;
            const u32 pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b0;
            const u32 bits = om & pm;
            if (!bits) continue;
            const u32 li = rb + __popc(starts & mask_le((u32)b0)) - 1u;
            const u32 root = s_par[li] & 0xFFFFu;
            u32 *sl = s_slot + (s_par[root] >> 16) * TILE_SLOT;
            const u32 n = __popc(bits);
            atomicAdd(sl + 0, n);
            atomicAdd(sl + 1, n * xl);
            atomicAdd(sl + 2, n * yl);
            atomicAdd(sl + 3, n * z0 + bitpos_sum(bits));
            // the box only ever grows: a (possibly stale) value that already covers this piece makes the atomic
            // unnecessary -- and neighbouring lanes mostly hit the SAME slot, where atomics serialise
            const u32 zlo = z0 + (u32)__ffs(bits) - 1u, zhi = z0 + 31u - (u32)__clz(bits);
            atomicMin(sl + 4, xl);
            atomicMin(sl + 5, yl);
            atomicMin(sl + 6, zlo);
            atomicMax(sl + 7, xl);
            atomicMax(sl + 8, yl);
            atomicMax(sl + 9, zhi);
        }
    }
    __syncthreads();
    SYN_TIMER_MARK(cnt, 6);   // statistics
    for (u32 idx = threadIdx.x; idx < nroots; idx += blockDim.x) {
        const u32 *sl = s_slot + idx * TILE_SLOT;
        TileRec rec;
        rec.root = s_rootgid[idx];
        rec.size = sl[0];
        rec.sx = (u64)sl[1] + (u64)sl[0] * x0;
        rec.sy = (u64)sl[2] + (u64)sl[0] * y0;
        rec.sz = sl[3];
        rec.minx = sl[4] + x0; rec.miny = sl[5] + y0; rec.minz = sl[6];
        rec.maxx = sl[7] + x0; rec.maxy = sl[8] + y0; rec.maxz = sl[9];
        // pad[0] = 1: the component cannot reach another tile (it touches no tile border that has a neighbour
        // tile), so this record IS the component: k_stats_merge stores it without atomics.  Not with dilation:
        // the box covers original voxels only, the connectivity runs through the dilated ones.
        const bool reach = (sl[4] == 0 && x0 > 0) || (sl[7] + 1 == vx && x0 + vx < (u32)g.nx) ||
                           (sl[5] == 0 && y0 > 0) || (sl[8] + 1 == vy && y0 + vy < (u32)g.ny);
        rec.pad[0] = (same_mask && sl[0] && !reach) ? 1u : 0u;
        rec.pad[1] = 0;
        recs[s_recbase + idx] = rec;           // a local root with no ORIGINAL voxel (dilation) has size 0: harmless
    }
    SYN_TIMER_MARK(cnt, 7);   // records
#ifdef SYN_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd((u64 *)&cnt->v[SYN_CNT_DBG0 + 8], 1ull);
#endif
}

// ===========================================================================
// K3 (6-connectivity): unions across tile borders.  Input: the compact list of non-empty words lying on a
// tile border (bndw, written by k_threshold_runs).  ONE LANE PER OVERLAP SPAN, as in the tile pass; the run
// ids come from the global prefix (runbase) and the unions go to the global parent array.
// ===========================================================================
__global__ void __launch_bounds__(256) k_union_bnd(const u32 *__restrict__ mask, RunIdx R,
                                                   u32 *par, Geom g, TileGeom tg, const u32 *__restrict__ bndw,
                                                   const u32 *__restrict__ tileflag, const u32 *__restrict__ actw,
                                                   const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    // tiles the shared-memory pass had to leave (too many runs; normally none): every row pair inside them
    union_rows_body<6, 2>(mask, R, par, g, tg, tileflag, actw, cnt);
    const u32 nb = (u32)cnt->v[SYN_CNT_BND_WORDS];
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    const u32 xstride = (u32)g.ny * g.wz;
    for (u32 a0 = warp * 32; a0 < nb; a0 += nwarps * 32) {
        const u32 a = a0 + lane;
        u32 w = 0, osy = 0, osx = 0, starts = 0, rb = 0, k = 0, row = 0;
        if (a < nb) {
            w = bndw[a];
            const u32 m = mask[w];
            u32 x, y;
            split_word(g, w, row, k);
            split_row(g, row, x, y);
            if (y && (y & (tg.ty - 1)) == 0) { const u32 o = m & mask[w - g.wz]; osy = o & ~(o << 1); }
            if (x && (x & (tg.tx - 1)) == 0) { const u32 o = m & mask[w - xstride]; osx = o & ~(o << 1); }
            if (osy | osx) {
                const u32 carry = k ? (mask[w - 1] >> 31) : 0u;
                starts = run_starts(m, carry);
                rb = R.at(w, row, k);
            }
        }
        const u32 ny_ = __popc(osy);
        const u32 n = ny_ + __popc(osx);
        const u32 incl = warp_incl_scan(n, lane);
        const u32 excl = incl - n;
        const u32 total = __shfl_sync(SYN_FULL, incl, 31);
        for (u32 r0 = 0; r0 < total; r0 += 32) {
            const u32 j = r0 + lane;
            int o = 0;
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                const u32 e = __shfl_sync(SYN_FULL, excl, (o + s) & 31);
                if (e <= j) o += s;
            }
            const u32 e_o = __shfl_sync(SYN_FULL, excl, o);
            const u32 osy_o = __shfl_sync(SYN_FULL, osy, o);
            const u32 osx_o = __shfl_sync(SYN_FULL, osx, o);
            const u32 w_o = __shfl_sync(SYN_FULL, w, o);
            const u32 st_o = __shfl_sync(SYN_FULL, starts, o);
            const u32 rb_o = __shfl_sync(SYN_FULL, rb, o);
            const u32 k_o = __shfl_sync(SYN_FULL, k, o);
            const u32 row_o = __shfl_sync(SYN_FULL, row, o);
            if (j < total) {
                u32 idx = j - e_o;
                const u32 nyo = __popc(osy_o);
                const bool isy = idx < nyo;
                if (!isy) idx -= nyo;
                const u32 b = select_bit(isy ? osy_o : osx_o, idx);
                const u32 wB = w_o - (isy ? g.wz : xstride);
                const u32 rowB = row_o - (isy ? 1u : (u32)g.ny);
                const u32 mB = mask[wB];
                const u32 carryB = k_o ? (mask[wB - 1] >> 31) : 0u;
                const u32 ga = rb_o + __popc(st_o & mask_le(b)) - 1u;
                const u32 gb = R.at(wB, rowB, k_o) + __popc(run_starts(mB, carryB) & mask_le(b)) - 1u;
                uf_union(par, ga, gb);
            }
        }
    }
}

// 18 / 26 connectivity: the same lane-per-span scheme, one neighbour ROW after the other, the z shifts of a row
// together: rows (x, y-1), (x-1, y) with shifts 0 / -1 / +1, rows (x-1, y-1), (x-1, y+1) with shift 0 (18) and
// 0 / -1 / +1 (26).  A word takes part for a row when that row pair crosses a tile border.
template <int CONN>
__global__ void __launch_bounds__(256) k_union_bnd_diag(const u32 *__restrict__ mask, RunIdx R, u32 *par, Geom g,
                                                        TileGeom tg, const u32 *__restrict__ bndw,
                                                        const u32 *__restrict__ tileflag, const u32 *__restrict__ actw,
                                                        const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    union_rows_body<CONN, 2>(mask, R, par, g, tg, tileflag, actw, cnt);      // flagged tiles (normally none)
    const u32 nb = (u32)cnt->v[SYN_CNT_BND_WORDS];
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    const int ny = (int)g.ny, wz = (int)g.wz;
    for (u32 a0 = warp * 32; a0 < nb; a0 += nwarps * 32) {
        const u32 a = a0 + lane;
        u32 w = 0, m = 0, starts = 0, ends = 0, rb = 0, k = 0, row = 0, x = 0, y = 0;
        bool xb = false, yb = false, ye = false;
        if (a < nb) {
            w = bndw[a];
            m = mask[w];
            split_word(g, w, row, k);
            split_row(g, row, x, y);
            xb = (x & (tg.tx - 1)) == 0;
            yb = (y & (tg.ty - 1)) == 0;
            ye = ((y + 1) & (tg.ty - 1)) == 0;
            starts = run_starts(m, k ? (mask[w - 1] >> 31) : 0u);
            ends = m & ~((m >> 1) | ((int)k + 1 < wz ? (mask[w + 1] << 31) : 0u));
            rb = R.at(w, row, k);
        }
#pragma unroll 1
        for (int nr = 0; nr < 4; ++nr) {
            // neighbour rows: 0 (x, y-1), 1 (x-1, y): z shifts 0, -1, +1;  2 (x-1, y-1), 3 (x-1, y+1): shift 0 (18),
            // 0, -1, +1 (26).  The spans of all shifts of a row go through ONE prefix sum / search.
            int drow;
            bool use;
            if (nr == 0) { drow = -1; use = y > 0 && yb; }
            else if (nr == 1) { drow = -ny; use = x > 0 && xb; }
            else if (nr == 2) { drow = -ny - 1; use = x > 0 && y > 0 && (xb || yb); }
            else { drow = -ny + 1; use = x > 0 && (int)y + 1 < ny && (xb || ye); }
            const bool shifted = CONN == 26 || nr < 2;
            u32 os0 = 0, os1 = 0, os2 = 0;
            if (a < nb && use && m) {
                const u32 *B = mask + (i64)((int)row + drow) * wz;
                const u32 cur = B[k];
                const u32 o = m & cur;
                os0 = o & ~(o << 1);
                if (shifted) {
                    // the diagonal contacts not implied by a direct overlap: a run starting at b where a neighbour
                    // run ends at b - 1, a run ending at b where a neighbour run starts at b + 1
                    os1 = starts & ((cur << 1) | (k > 0 ? (B[k - 1] >> 31) : 0u)) & ~cur;
                    os2 = ends & ((cur >> 1) | ((int)k + 1 < wz ? (B[k + 1] << 31) : 0u)) & ~cur;
                }
            }
            const u32 n = __popc(os0) + __popc(os1) + __popc(os2);
            const u32 incl = warp_incl_scan(n, lane);
            const u32 excl = incl - n;
            const u32 total = __shfl_sync(SYN_FULL, incl, 31);
            for (u32 r0 = 0; r0 < total; r0 += 32) {
                const u32 j = r0 + lane;
                int o = 0;
#pragma unroll
                for (int s = 16; s; s >>= 1) {
                    const u32 e = __shfl_sync(SYN_FULL, excl, (o + s) & 31);
                    if (e <= j) o += s;
                }
                const u32 e_o = __shfl_sync(SYN_FULL, excl, o);
                const u32 os0_o = __shfl_sync(SYN_FULL, os0, o);
                const u32 os1_o = __shfl_sync(SYN_FULL, os1, o);
                const u32 os2_o = __shfl_sync(SYN_FULL, os2, o);
                const u32 st_o = __shfl_sync(SYN_FULL, starts, o);
                const u32 rb_o = __shfl_sync(SYN_FULL, rb, o);
                const u32 k_o = __shfl_sync(SYN_FULL, k, o);
                const u32 row_o = __shfl_sync(SYN_FULL, row, o);
                if (j < total) {
                    u32 idx = j - e_o, os_o = os0_o;
                    int shift = 0;
                    const u32 n0 = __popc(os0_o), n1 = __popc(os1_o);
                    if (idx >= n0 + n1) { idx -= n0 + n1; os_o = os2_o; shift = 1; }
                    else if (idx >= n0) { idx -= n0; os_o = os1_o; shift = -1; }
                    const u32 b = select_bit(os_o, idx);
                    const u32 ga = rb_o + __popc(st_o & mask_le(b)) - 1u;
                    const u32 rowB = (u32)((int)row_o + drow);
                    int bb = (int)b + shift, kb = (int)k_o;
                    if (bb < 0) { bb = 31; --kb; }
                    else if (bb > 31) { bb = 0; ++kb; }
                    const u32 wB = rowB * g.wz + (u32)kb;
                    const u32 mB = mask[wB];
                    const u32 carryB = kb ? (mask[wB - 1] >> 31) : 0u;
                    const u32 gb = R.at(wB, rowB, (u32)kb) + __popc(run_starts(mB, carryB) & mask_le((u32)bb)) - 1u;
                    uf_union(par, ga, gb);
                }
            }
        }
    }
}

// ===========================================================================
// K4: flatten every run to its root, flag roots, rank the roots (scan #2).  The
// rank of a root among all roots in run-id order is the canonical component
// number minus one: run ids are in raster order of the run's first voxel and
// the root is the smallest run of its component.
// ===========================================================================
struct RootFlagSrc {  // flattens while flagging
    u32 *par;
    const Counts *cnt;
    __device__ __forceinline__ u64 n() const
    {
        return (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) ? 0 : (u64)cnt->v[SYN_CNT_RUNS];
    }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
        u32 p[N], x[N], r[N];
        load_items_u32(par, i0, n, p);
        // walk all N chains in lockstep, one hop per round: the loads of a round are independent of each other
        // (most runs are roots or children of a root; a component spanning several tiles needs 2-4 hops)
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const u32 g = (u32)(i0 + t);
            const bool valid = i0 + t < n;
            x[t] = valid ? p[t] : 0u;
            r[t] = (valid && p[t] != g) ? ldcg_u32(par + p[t]) : x[t];
        }
        bool pending = true;
        while (pending) {
            pending = false;
#pragma unroll
            for (int t = 0; t < N; ++t) {
                if (r[t] != x[t]) {
                    x[t] = r[t];
                    r[t] = ldcg_u32(par + x[t]);
                    pending = true;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < N; ++t) {
            item[t] = 0;
            if (i0 + t < n) {
                const u32 g = (u32)(i0 + t);
                if (x[t] != p[t]) par[g] = x[t];    // flatten (a concurrent reader sees the old ancestor or the root)
                item[t] = (x[t] == g);
            }
        }
    }
};
// rank[g] = canonical index for roots; also initialises the root's CompStat record
struct RankDst {
    u32 *rank;
    CompStat *stat;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
        store_prefix_u32(rank, i0, n, ex, item);
#pragma unroll
        for (int t = 0; t < N; ++t) {
            if (item[t]) {
                CompStat s;
                s.size = s.sx = s.sy = s.sz = 0;
                s.minx = s.miny = s.minz = 0xFFFFFFFFu;
                s.maxx = s.maxy = s.maxz = 0;
                s.keep = 0;
                s.row = 0;
                stat[ex] = s;
            }
            ex += item[t];
        }
    }
};

// ===========================================================================
// K5: per-component size / coordinate sums / bounding box.  One thread per mask
// word; every maximal span of set bits inside the word ("piece") belongs to one
// run, hence one component.  In the dilation variant the spans come from the
// dilated mask but only voxels of the original mask are counted
// (conncomps.py:49-50 `ccs[mask == 0] = 0`).
// ===========================================================================
__device__ __forceinline__ void stats_body(const u32 *__restrict__ lmask, const u32 *__restrict__ omask,
                                           const RunIdx &R, const u32 *__restrict__ par,
                                           const u32 *__restrict__ rank, CompStat *stat, const Geom &g,
                                           const TileGeom &tg, const u32 *__restrict__ tileflag,
                                           const u32 *__restrict__ actw, const Counts *cnt)
{
    // with tiling, the tile pass already produced the statistics of every tile it did not flag
    if (tg.tx && cnt->v[SYN_CNT_FLAGGED_TILES] == 0) return;
    const u32 nact = (u32)cnt->v[SYN_CNT_ACTIVE_WORDS];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nact; i += gridDim.x * blockDim.x) {
        const u32 w = actw[i];
        u32 m = lmask[w];
        u32 om = omask[w];
        if (om == 0) continue;
        u32 row, k, x, y;
        split_word(g, w, row, k);
        split_row(g, row, x, y);
        if (tg.tx && tileflag[(x >> tg.txs) * tg.tiles_y + (y >> tg.tys)] == 0) continue;
        u32 starts = R.starts(lmask, w, k);
        u32 rb = R.at(w, row, k);
        u32 ps = (m & ~(m << 1)) | starts;  // piece starts (a run continuing from the previous word starts a piece too)
        while (ps) {
            int b0 = __ffs(ps) - 1;
            ps &= ps - 1;
            u32 span = ~(m >> b0);                       // zero bits above the piece
            int len = span ? (__ffs(span) - 1) : 32;     // span==0 only when b0==0 and m is all ones
            if (len > 32 - b0) len = 32 - b0;
            if (ps) len = min(len, __ffs(ps) - 1 - b0);  // (multi-label) the next run may start right behind
            u32 pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b0;
            u32 bits = om & pm;
            if (!bits) continue;
            u32 gidx = rb + __popc(starts & mask_le((u32)b0)) - 1u;
            u32 c = rank[par[gidx]];
            CompStat *s = stat + c;
            u32 n = __popc(bits);
            u32 z0 = k * 32;
            atomicAdd(&s->size, (u64)n);
            atomicAdd(&s->sx, (u64)n * x);
            atomicAdd(&s->sy, (u64)n * y);
            atomicAdd(&s->sz, (u64)n * z0 + bitpos_sum(bits));
            // the box only ever grows: a (possibly stale) read that already covers this piece makes the
            // atomic unnecessary, which removes most of the min/max traffic
            const u32 zlo = z0 + (u32)__ffs(bits) - 1u, zhi = z0 + 31u - (u32)__clz(bits);
            const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(&s->minx));    // two loads, after the sums: see
            const uint2 hi = __ldcg(reinterpret_cast<const uint2 *>(&s->maxy));    // k_stats_merge
            if (x < lo.x) atomicMin(&s->minx, x);
            if (y < lo.y) atomicMin(&s->miny, y);
            if (zlo < lo.z) atomicMin(&s->minz, zlo);
            if (x > lo.w) atomicMax(&s->maxx, x);
            if (y > hi.x) atomicMax(&s->maxy, y);
            if (zhi > hi.y) atomicMax(&s->maxz, zhi);
        }
    }
}

// Merges the tile records into the per-component statistics (after the canonical ranks are known); tiles whose
// statistics the shared-memory pass could not hold (flagged, normally none) are accumulated per piece.
__global__ void __launch_bounds__(256) k_stats_merge(const TileRec *__restrict__ recs, const u32 *__restrict__ par,
                                                     const u32 *__restrict__ rank, CompStat *stat,
                                                     const u32 *__restrict__ lmask, const u32 *__restrict__ omask,
                                                     RunIdx R, Geom g, TileGeom tg, const u32 *__restrict__ tileflag,
                                                     const u32 *__restrict__ actw, const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u64 n = (u64)cnt->v[SYN_CNT_TILE_RECORDS];
    // A component that spans thousands of tiles (dilation) would serialise thousands of same-address atomics in L2.
    // Each CTA therefore picks ONE candidate -- the component of the largest record among its first 256 -- and sums
    // that component's records in shared memory; everything else goes straight to the global accumulators.
    __shared__ unsigned long long s_best;            // size << 32 | component of the CTA's largest first-round record
    __shared__ unsigned long long s_acc[4];          // size, sx, sy, sz of the candidate
    __shared__ u32 s_box[6];
    if (threadIdx.x == 0) {
        s_best = 0;
        s_acc[0] = s_acc[1] = s_acc[2] = s_acc[3] = 0;
        s_box[0] = s_box[1] = s_box[2] = 0xFFFFFFFFu;
        s_box[3] = s_box[4] = s_box[5] = 0;
    }
    __syncthreads();
    bool first = true;
    u32 cand = 0xFFFFFFFFu;
    for (u64 i0 = blockIdx.x * (u64)blockDim.x; i0 < n; i0 += (u64)gridDim.x * blockDim.x) {
        const u64 i = i0 + threadIdx.x;
        TileRec r;
        r.size = 0;
        u32 c = 0;
        if (i < n) {
            r = recs[i];
            if (r.size) c = rank[par[r.root]];
        }
        if (first) {
            if (r.size >= 64) atomicMax(&s_best, ((unsigned long long)r.size << 32) | c);
            __syncthreads();
            cand = s_best ? (u32)s_best : 0xFFFFFFFFu;
            first = false;
        }
        if (r.size == 0) continue;
        if (r.pad[0]) {
            // a component that lives in one tile: its single record is stored as it is (no atomics; the look-back
            // scan initialised the same slot before this kernel started)
            CompStat s;
            s.size = r.size; s.sx = r.sx; s.sy = r.sy; s.sz = r.sz;
            s.minx = r.minx; s.miny = r.miny; s.minz = r.minz;
            s.maxx = r.maxx; s.maxy = r.maxy; s.maxz = r.maxz;
            s.keep = 0; s.row = 0;
            stat[c] = s;
            continue;
        }
        if (c == cand) {
            atomicAdd(&s_acc[0], (unsigned long long)r.size);
            atomicAdd(&s_acc[1], (unsigned long long)r.sx);
            atomicAdd(&s_acc[2], (unsigned long long)r.sy);
            atomicAdd(&s_acc[3], (unsigned long long)r.sz);
            atomicMin(&s_box[0], r.minx); atomicMin(&s_box[1], r.miny); atomicMin(&s_box[2], r.minz);
            atomicMax(&s_box[3], r.maxx); atomicMax(&s_box[4], r.maxy); atomicMax(&s_box[5], r.maxz);
            continue;
        }
        CompStat *s = stat + c;
        atomicAdd(&s->size, (u64)r.size);
        atomicAdd(&s->sx, r.sx);
        atomicAdd(&s->sy, r.sy);
        atomicAdd(&s->sz, r.sz);
        // the box only ever grows: a (possibly stale) read that already covers the record makes the atomic
        // unnecessary.  The six values are read with TWO vector loads, issued after the sums went out (the reds pull
        // the record's line into L2, the loads hit it): six scalar loads between the atomics were six serialised L2
        // round trips, 55 % of this kernel's stall samples (dense config 64.7 -> 52.9 us; the same two loads placed
        // BEFORE the sums: 78 us -- they miss, and the reds queue behind the fill)
        const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(&s->minx));    // minx miny minz maxx
        const uint2 hi = __ldcg(reinterpret_cast<const uint2 *>(&s->maxy));    // maxy maxz
        if (r.minx < lo.x) atomicMin(&s->minx, r.minx);
        if (r.miny < lo.y) atomicMin(&s->miny, r.miny);
        if (r.minz < lo.z) atomicMin(&s->minz, r.minz);
        if (r.maxx > lo.w) atomicMax(&s->maxx, r.maxx);
        if (r.maxy > hi.x) atomicMax(&s->maxy, r.maxy);
        if (r.maxz > hi.y) atomicMax(&s->maxz, r.maxz);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_acc[0]) {
        CompStat *s = stat + (u32)s_best;
        atomicAdd(&s->size, (u64)s_acc[0]);
        atomicAdd(&s->sx, (u64)s_acc[1]);
        atomicAdd(&s->sy, (u64)s_acc[2]);
        atomicAdd(&s->sz, (u64)s_acc[3]);
        atomicMin(&s->minx, s_box[0]); atomicMin(&s->miny, s_box[1]); atomicMin(&s->minz, s_box[2]);
        atomicMax(&s->maxx, s_box[3]); atomicMax(&s->maxy, s_box[4]); atomicMax(&s->maxz, s_box[5]);
    }
    stats_body(lmask, omask, R, par, rank, stat, g, tg, tileflag, actw, cnt);
}

// ===========================================================================
// K6: dust decision + compaction of the segment table (scan #3 over components).
// keep = size >= dust  OR  the component touches one of the six chunk faces
// (continuations are never filtered: proc/tasks.py:45-51).
// ===========================================================================
struct KeepSrc {
    const CompStat *stat;
    const Counts *cnt;
    i64 dust;
    u32 mx, my, mz;  // nx-1, ny-1, nz-1
    __device__ __forceinline__ u64 n() const
    {
        return (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) ? 0 : (u64)cnt->v[SYN_CNT_COMPONENTS];
    }
    template <int N>
    __device__ __forceinline__ void load(u64 i0, u64 n, u32 (&item)[N]) const
    {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            item[t] = 0;
            if (i0 + t < n) {
                const CompStat *s = stat + i0 + t;
                const u64 size = __ldcg(&s->size);
                const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(&s->minx));   // minx miny minz maxx
                const uint2 hi = __ldcg(reinterpret_cast<const uint2 *>(&s->maxy));   // maxy maxz
                const bool face = lo.x == 0 || lo.y == 0 || lo.z == 0 || lo.w == mx || hi.x == my || hi.y == mz;
                item[t] = (size > 0) && (face || (i64)size >= dust);
            }
        }
    }
};

// centroid exactly as _describe.cpp:54-61: float32(sum) / float32(size), round half away
__device__ __forceinline__ i64 centroid_f32(u64 sum, u64 size)
{
    float q = __fdiv_rn(__ll2float_rn((i64)sum), __ll2float_rn((i64)size));
    return (i64)roundf(q);
}

// complab[c] = label of component c: c + 1 if kept, else 0 (`rel`: its table row + 1 instead -- the chunk-relative
// id the cross-chunk merge works with); kept components get a table row
struct TableDst {
    const CompStat *stat;
    u32 *complab;
    i64 *table;
    i64 max_segs;
    i64 ox, oy, oz;
    bool rel;
    template <int N>
    __device__ __forceinline__ void store(u64 i0, u64 n, u32 ex, const u32 (&item)[N]) const
    {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            const u64 c = i0 + t;
            if (c < n) complab[c] = item[t] ? (rel ? ex + 1u : (u32)c + 1u) : 0u;
            if (item[t]) {
                const u32 rowi = ex;
                if ((i64)rowi < max_segs) {
                    const CompStat s = stat[c];
                    i64 *r = table + (i64)rowi * SYN_SEG_NCOLS;
                    r[SYN_SEG_ID] = (i64)c + 1;
                    r[SYN_SEG_SIZE] = (i64)s.size;
                    r[SYN_SEG_CX] = centroid_f32(s.sx, s.size) + ox;
                    r[SYN_SEG_CY] = centroid_f32(s.sy, s.size) + oy;
                    r[SYN_SEG_CZ] = centroid_f32(s.sz, s.size) + oz;
                    r[SYN_SEG_BX] = (i64)s.minx + ox;
                    r[SYN_SEG_BY] = (i64)s.miny + oy;
                    r[SYN_SEG_BZ] = (i64)s.minz + oz;
                    r[SYN_SEG_EX] = (i64)s.maxx + 1 + ox;
                    r[SYN_SEG_EY] = (i64)s.maxy + 1 + oy;
                    r[SYN_SEG_EZ] = (i64)s.maxz + 1 + oz;
                }
            }
            ex += item[t];
        }
    }
};

// final per-run label: the component's label if kept else 0; also tracks the largest label.  With `lut` (the
// deferred pass: complab holds table rows + 1) the label goes through the cross-chunk id map first:
// label = lut[lut_base[0] + row + 1] (remap_ids_task, proc/tasks.py:314-333, folded into the label write).
__global__ void __launch_bounds__(256) k_run_labels(const u32 *__restrict__ par, const u32 *__restrict__ rank,
                                                    const u32 *__restrict__ complab, u32 *__restrict__ runlabel,
                                                    Counts *cnt, i64 max_segs, const u32 *__restrict__ lut,
                                                    const i64 *__restrict__ lut_base, i64 lut_len)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u64 nruns = (u64)cnt->v[SYN_CNT_RUNS];
    const i64 lb = lut ? lut_base[0] : 0;
    u32 mx = 0;
    // 4 consecutive runs per thread: 16-byte loads of the (flattened) parents, 16-byte stores of the labels
    for (u64 g0 = (blockIdx.x * (u64)blockDim.x + threadIdx.x) * 4; g0 < nruns; g0 += (u64)gridDim.x * blockDim.x * 4) {
        u32 p[4], lab[4];
        if (g0 + 4 <= nruns) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(par + g0));
            p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) p[t] = (g0 + t < nruns) ? __ldcg(par + g0 + t) : 0u;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) lab[t] = (g0 + t < nruns) ? __ldcg(rank + p[t]) : 0u;
#pragma unroll
        for (int t = 0; t < 4; ++t) lab[t] = (g0 + t < nruns) ? __ldcg(complab + lab[t]) : 0u;
        if (lut) {
#pragma unroll
            for (int t = 0; t < 4; ++t) lab[t] = (lab[t] && lb + (i64)lab[t] < lut_len) ? __ldg(lut + lb + lab[t]) : 0u;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) mx = max(mx, lab[t]);
        if (g0 + 4 <= nruns) *reinterpret_cast<uint4 *>(runlabel + g0) = make_uint4(lab[0], lab[1], lab[2], lab[3]);
        else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (g0 + t < nruns) runlabel[g0 + t] = lab[t];
        }
    }
    mx = __reduce_max_sync(SYN_FULL, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_LABEL], (u64)mx);
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt->v[SYN_CNT_KEPT] > max_segs)
        atomicOr((u64 *)&cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_SEGS);
}

}  // namespace syn
