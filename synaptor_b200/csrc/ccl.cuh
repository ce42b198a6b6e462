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
// K1d: 2-D L1 dilation by k over axes (y, z) on the bitmask -- what the chamfer
// distance transform + `dist > k -> 0` of _seg_utils.cpp:25-82,180-200 computes
// for a 0/1 mask.  out(y, z) = OR_{|dy| <= k} zdil(in(y + dy), k - |dy|).
// One thread per output word; the bitmask is L2 resident.  k <= 32.
// ===========================================================================
__device__ __forceinline__ u32 zdil_word(u32 prev, u32 cur, u32 next, int r)
{
    // bits of `cur` row-word after smearing every set bit by +-r along z
    u32 out = cur;
    u64 lo = ((u64)cur << 32) | prev;   // cur in the high half: (lo >> (32 - s)) == (cur << s) | (prev >> (32 - s))
    u64 hi = ((u64)next << 32) | cur;   // (hi >> s) low half == (cur >> s) | (next << (32 - s))
    for (int s = 1; s <= r; ++s) {
        out |= (u32)(lo >> (32 - s));
        out |= (u32)(hi >> s);
    }
    return out;
}

__global__ void __launch_bounds__(256) k_dilate(const u32 *__restrict__ in, u32 *__restrict__ out, Geom g, int k)
{
    const u32 tail = (g.nz & 31) ? ((1u << (g.nz & 31)) - 1u) : 0xFFFFFFFFu;
    for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < g.nwords; w += gridDim.x * blockDim.x) {
        u32 row = w / g.wz;
        u32 kw = w - row * g.wz;
        u32 y = row % (u32)g.ny;
        u32 acc = 0;
        for (int dy = -k; dy <= k; ++dy) {
            int yy = (int)y + dy;
            if (yy < 0 || yy >= (int)g.ny) continue;
            const u32 *r = in + (i64)(row + dy) * g.wz;  // same x, row index shifts by dy
            u32 cur = r[kw];
            u32 prev = kw > 0 ? r[kw - 1] : 0u;
            u32 next = kw + 1 < g.wz ? r[kw + 1] : 0u;
            if ((cur | prev | next) == 0) continue;
            int rr = k - (dy < 0 ? -dy : dy);
            acc |= zdil_word(prev, cur, next, rr);
        }
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
// must be zero before the launch.  `Src` provides n() (item count, may live on the
// device) and get(i); `Dst` receives put(i, exclusive prefix, item value).
// ===========================================================================
template <class Src, class Dst>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_1p(Src src, Dst dst, volatile u64 *state, u32 *ticket,
                                                          i64 *total_out)
{
    __shared__ u32 sm[33];
    __shared__ u32 s_tile, s_prefix;
    const u64 n = src.n();
    const u32 ntiles = (u32)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (ntiles == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
        return;
    }
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;  // block-uniform
        u32 item[SCAN_ITEMS];
        u32 v = 0;
        const u64 i0 = (u64)tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t) {
            item[t] = (i0 + t < n) ? src.get(i0 + t) : 0;
            v += item[t];
        }
        u32 total;
        u32 ex = block_excl_scan(v, sm, &total);
        if (threadIdx.x < 32) {
            // warp-wide look-back: lane l inspects tile (p - l); tiles before 0 count as a zero prefix
            const int lane = threadIdx.x;
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
            if (lane == 0) {
                state[tile] = (2ull << 32) | (u64)(prefix + total);
                s_prefix = prefix;
                if (tile == ntiles - 1 && total_out) *total_out = (i64)(prefix + total);
            }
        }
        __syncthreads();
        ex += s_prefix;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t) {
            if (i0 + t < n) dst.put(i0 + t, ex, item[t]);
            ex += item[t];
        }
        __syncthreads();   // s_tile / s_prefix are rewritten in the next round
    }
}

// --- scan #1: run starts per mask word -> runbase[w] ------------------------
struct RunStartSrc {
    const u32 *mask;
    u32 nwords, wz;
    __device__ __forceinline__ u64 n() const { return nwords; }
    __device__ __forceinline__ u32 get(u64 w) const
    {
        u32 m = mask[w];
        if (m == 0) return 0;
        u32 k = (u32)w % wz;
        u32 carry = k ? (mask[w - 1] >> 31) : 0u;
        return __popc(run_starts(m, carry));
    }
};
struct RunBaseDst {
    u32 *runbase;
    __device__ __forceinline__ void put(u64 w, u32 ex, u32) const { runbase[w] = ex; }
};


// ---------------------------------------------------------------------------
// Compaction of the non-empty mask words.  The sparse kernels (unions, statistics,
// foreground sectors) run one thread per ACTIVE word from this list, so every lane
// of a warp has work (3 % foreground means ~10 % of the words are non-empty; a
// thread-per-word grid over all words would run at ~2 active lanes per warp).
// Order inside a block is raster order; blocks append with one atomicAdd each.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_compact_active(const u32 *__restrict__ mask, u32 nwords,
                                                                 u32 *__restrict__ actw, Counts *cnt)
{
    __shared__ u32 sm[33];
    __shared__ u32 s_base;
    for (u64 tile = blockIdx.x; tile * SCAN_TILE < nwords; tile += gridDim.x) {
        const u64 i0 = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
        u32 flag[SCAN_ITEMS];
        u32 v = 0;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t) {
            flag[t] = (i0 + t < nwords) && (mask[i0 + t] != 0);
            v += flag[t];
        }
        u32 total;
        u32 ex = block_excl_scan(v, sm, &total);
        if (threadIdx.x == 0) s_base = total ? (u32)atomicAdd((u64 *)&cnt->v[SYN_CNT_ACTIVE_WORDS], (u64)total) : 0u;
        __syncthreads();
        ex += s_base;
#pragma unroll
        for (int t = 0; t < SCAN_ITEMS; ++t)
            if (flag[t]) actw[ex++] = (u32)(i0 + t);
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
struct TileGeom {
    u32 tx, ty;          // rows per tile along x and y, powers of two (tx == 0: tiling disabled)
    u32 tiles_x, tiles_y;
    u32 ntiles;
    u32 txs, tys;        // log2(tx), log2(ty)
};

struct RowView {
    const u32 *mask;     // words of this row
    const u32 *runbase;  // run-start prefix of this row's words
    u32 wz;
    __device__ __forceinline__ u32 word(int k) const { return (k >= 0 && k < (int)wz) ? mask[k] : 0u; }
    // id of the run covering bit b (0..31) of word k; that bit must be set
    __device__ __forceinline__ u32 run_at(int k, int b) const
    {
        u32 m = mask[k];
        u32 carry = k ? (mask[k - 1] >> 31) : 0u;
        return runbase[k] + __popc(run_starts(m, carry) & mask_le((u32)b)) - 1u;
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
    while (os) {
        int b = __ffs(os) - 1;
        os &= os - 1;
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
template <int CONN>
__global__ void __launch_bounds__(256) k_union_rows(const u32 *__restrict__ mask, const u32 *__restrict__ runbase,
                                                    u32 *par, Geom g, TileGeom tg, const u32 *__restrict__ tileflag,
                                                    const u32 *__restrict__ actw, const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u32 nact = (u32)cnt->v[SYN_CNT_ACTIVE_WORDS];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nact; i += gridDim.x * blockDim.x) {
        const u32 w = actw[i];
        u32 m = mask[w];
        u32 row, kk, x, y;
        split_word(g, w, row, kk);
        split_row(g, row, x, y);
        const int k = (int)kk;
        RowView A{mask + (u64)row * g.wz, runbase + (u64)row * g.wz, g.wz};
        bool do_y = true, do_x = true, do_dm = true, do_dp = true;   // (x,y-1) (x-1,y) (x-1,y-1) (x-1,y+1)
        if (tg.tx) {
            const u32 tX = x >> tg.txs, tY = y >> tg.tys;
            const bool dense = tileflag[tX * tg.tiles_y + tY] != 0;
            const bool xb = (x & (tg.tx - 1)) == 0, yb = (y & (tg.ty - 1)) == 0, ye = ((y + 1) & (tg.ty - 1)) == 0;
            do_y = dense || yb;
            do_x = dense || xb;
            do_dm = dense || xb || yb;
            do_dp = dense || xb || ye;
            if (!(do_y || do_x || do_dm || do_dp)) continue;
        }
        if (y > 0 && do_y) {
            u32 r2 = row - 1;
            RowView B{mask + (u64)r2 * g.wz, runbase + (u64)r2 * g.wz, g.wz};
            union_shifted<0>(par, A, B, k, m);
            if (CONN >= 18) { union_shifted<-1>(par, A, B, k, m); union_shifted<1>(par, A, B, k, m); }
        }
        if (x > 0) {
            u32 r2 = row - (u32)g.ny;
            RowView B{mask + (u64)r2 * g.wz, runbase + (u64)r2 * g.wz, g.wz};
            if (do_x) {
                union_shifted<0>(par, A, B, k, m);
                if (CONN >= 18) { union_shifted<-1>(par, A, B, k, m); union_shifted<1>(par, A, B, k, m); }
            }
            if (CONN >= 18 && y > 0 && do_dm) {
                u32 r3 = r2 - 1;
                RowView C{mask + (u64)r3 * g.wz, runbase + (u64)r3 * g.wz, g.wz};
                union_shifted<0>(par, A, C, k, m);
                if (CONN == 26) { union_shifted<-1>(par, A, C, k, m); union_shifted<1>(par, A, C, k, m); }
            }
            if (CONN >= 18 && y + 1 < (u32)g.ny && do_dp) {
                u32 r3 = r2 + 1;
                RowView C{mask + (u64)r3 * g.wz, runbase + (u64)r3 * g.wz, g.wz};
                union_shifted<0>(par, A, C, k, m);
                if (CONN == 26) { union_shifted<-1>(par, A, C, k, m); union_shifted<1>(par, A, C, k, m); }
            }
        }
    }
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
constexpr int TILE_WORDS = 2048;
constexpr int TILE_RUNS = 4096;
constexpr int TILE_MAXTX = 16;
constexpr int TILE_ROOTS = 512;    // tile-local components with statistics accumulated in shared memory
constexpr int TILE_SLOT = 10;      // u32 per root: size, sum dx, sum dy, sum z, min/max dx, dy, z
constexpr size_t TILE_SMEM = (size_t)TILE_WORDS * 4 * 3 + (size_t)TILE_WORDS * 2 + (size_t)TILE_RUNS * 4 +
                             (size_t)TILE_RUNS * 2 + (size_t)TILE_ROOTS * 2 + (size_t)TILE_ROOTS * TILE_SLOT * 4;


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

// One pass over the tile's non-empty words for one (neighbour row, z shift) combination, ONE LANE PER
// OVERLAP SPAN: every thread first finds the overlap spans of its word with the neighbour word, a warp
// prefix sum turns the 32 counts into a work list, and the lanes then take the spans one each (owner by
// shuffle binary search, span start by select_bit) and perform the unions with uniform control flow.
// s_act entry: word index (11 bits) | k << 11 (10 bits) | yl << 21 (6 bits) | xl << 27.
template <int SHIFT>
__device__ __forceinline__ void tile_pass(const u32 *s_mask, const unsigned short *s_rb, const u32 *s_st, u32 *s_par,
                                          const u32 *s_act, u32 nact, u32 wz, u32 vy, u32 dxl, int dyl, int off)
{
    const int lane = threadIdx.x & 31;
    const u32 wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    for (u32 a0 = wib * 32; a0 < nact; a0 += nwb * 32) {
        const u32 a = a0 + lane;
        u32 os = 0, i = 0;
        if (a < nact) {
            const u32 e = s_act[a];
            i = e & 2047u;
            const u32 k = (e >> 11) & 1023u;
            const int yl = (int)((e >> 21) & 63u);
            const u32 xl = e >> 27;
            if (xl >= dxl && yl + dyl >= 0 && yl + dyl < (int)vy) {
                const int iB = (int)i + off;
                u32 nb;
                if (SHIFT == 0) nb = s_mask[iB];
                else if (SHIFT < 0) nb = (s_mask[iB] << 1) | (k > 0 ? (s_mask[iB - 1] >> 31) : 0u);
                else nb = (s_mask[iB] >> 1) | (k + 1 < wz ? (s_mask[iB + 1] << 31) : 0u);
                const u32 o = s_mask[i] & nb;
                os = o & ~(o << 1);
            }
        }
        const u32 n = __popc(os);
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
            const u32 os_o = __shfl_sync(SYN_FULL, os, o);
            const u32 i_o = __shfl_sync(SYN_FULL, i, o);
            if (j < total) {
                const u32 b = select_bit(os_o, j - e_o);
                const u32 ga = (u32)s_rb[i_o] + __popc(s_st[i_o] & mask_le(b)) - 1u;
                int iB = (int)i_o + off, bb = (int)b + SHIFT;
                if (bb < 0) { bb = 31; --iB; }
                else if (bb > 31) { bb = 0; ++iB; }
                const u32 gb = (u32)s_rb[iB] + __popc(s_st[iB] & mask_le((u32)bb)) - 1u;
                sm_union(s_par, ga, gb);
            }
        }
    }
}

constexpr int TILE_THREADS = 512;   // 3 CTAs/SM by shared memory -> 48 warps/SM

template <int CONN>
__global__ void __launch_bounds__(TILE_THREADS) k_tile_ccl(const u32 *__restrict__ mask, const u32 *__restrict__ omask,
                                                  const u32 *__restrict__ runbase, u32 *__restrict__ par,
                                                  u32 *__restrict__ tileflag, TileRec *__restrict__ recs, Geom g,
                                                  TileGeom tg, Counts *cnt)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    u32 *s_mask = reinterpret_cast<u32 *>(s_raw);                 // [TILE_WORDS]
    u32 *s_st = s_mask + TILE_WORDS;                              // run-start bits of the non-empty words
    u32 *s_act = s_st + TILE_WORDS;                               // packed coordinates of the non-empty words
    u32 *s_par = s_act + TILE_WORDS;                              // [TILE_RUNS]
    u32 *s_slot = s_par + TILE_RUNS;                              // [TILE_ROOTS][TILE_SLOT] statistics
    unsigned short *s_rb = reinterpret_cast<unsigned short *>(s_slot + TILE_ROOTS * TILE_SLOT);   // LOCAL run-id prefix
    unsigned short *s_rootidx = s_rb + TILE_WORDS;                // [TILE_RUNS] compact index of a local root
    unsigned short *s_rootrun = s_rootidx + TILE_RUNS;            // [TILE_ROOTS] local run id of root #idx
    __shared__ u32 s_nact, s_nroots, s_recbase;
    __shared__ u32 s_segoff[TILE_MAXTX + 1];       // local id of the first run of x-slab segment xl
    __shared__ u32 s_delta[TILE_MAXTX];            // global id = local id + delta[xl]
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u32 nruns = (u32)cnt->v[SYN_CNT_RUNS];
    const u32 tile = blockIdx.x;
    const u32 tX = tile / tg.tiles_y, tY = tile - tX * tg.tiles_y;
    const u32 x0 = tX * tg.tx, y0 = tY * tg.ty;
    const u32 vx = min(tg.tx, (u32)g.nx - x0), vy = min(tg.ty, (u32)g.ny - y0);   // valid rows in this tile
    const u32 segw = vy * g.wz;                // words per x-slab segment (contiguous in memory)

    if (threadIdx.x < vx) {
        u32 first = ((x0 + threadIdx.x) * (u32)g.ny + y0) * g.wz;
        u32 end = first + segw;
        u32 b0 = runbase[first];
        u32 b1 = end < g.nwords ? runbase[end] : nruns;
        s_segoff[threadIdx.x + 1] = b1 - b0;   // run count, turned into a prefix below
        s_delta[threadIdx.x] = b0;
    }
    if (threadIdx.x == 0) { s_nact = 0; s_nroots = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 acc = 0;
        s_segoff[0] = 0;
        for (u32 i = 0; i < vx; ++i) {
            u32 c = s_segoff[i + 1];
            s_delta[i] -= acc;                 // global = local - segoff + segbase
            acc += c;
            s_segoff[i + 1] = acc;
        }
    }
    __syncthreads();
    const u32 nloc = s_segoff[vx];
    if (nloc == 0) return;                     // no foreground in this tile
    if (nloc > TILE_RUNS) {                    // too dense for the shared-memory pass: K3 / k_stats do the whole tile
        if (threadIdx.x == 0) {
            tileflag[tile] = 1;
            atomicAdd((u64 *)&cnt->v[SYN_CNT_FLAGGED_TILES], 1ull);
        }
        return;
    }
    // load the tile, x-slab segment by segment (each is contiguous in memory)
    for (u32 xl = 0; xl < vx; ++xl) {
        const u32 gw0 = ((x0 + xl) * (u32)g.ny + y0) * g.wz;
        const u32 d = s_delta[xl];
        for (u32 jj = threadIdx.x; jj < segw; jj += blockDim.x) {   // segw need not be a multiple of 32
            const u32 i = xl * segw + jj;
            const u32 m = mask[gw0 + jj];
            s_mask[i] = m;
            s_rb[i] = (unsigned short)(runbase[gw0 + jj] - d);
            u32 yl, k;
            if (g.wzs < 32) { yl = jj >> g.wzs; k = jj & (g.wz - 1u); }
            else { yl = jj / g.wz; k = jj - yl * g.wz; }
            const u32 act = __activemask();
            const u32 bal = __ballot_sync(act, m != 0);
            if (m) {
                const u32 carry = k ? (mask[gw0 + jj - 1] >> 31) : 0u;
                s_st[i] = run_starts(m, carry);
            }
            if (bal) {
                const int lane = threadIdx.x & 31;
                const int leader = __ffs(bal) - 1;
                u32 basei = 0;
                if (lane == leader) basei = atomicAdd(&s_nact, (u32)__popc(bal));
                basei = __shfl_sync(act, basei, leader);
                if (m) s_act[basei + __popc(bal & mask_lt((u32)lane))] = i | (k << 11) | (yl << 21) | (xl << 27);
            }
        }
    }
    for (u32 i = threadIdx.x; i < nloc; i += blockDim.x) s_par[i] = i;
    __syncthreads();

    const u32 nact = s_nact;
    const int wz = (int)g.wz, sg = (int)segw;
    // (x, y-1) and (x-1, y): face neighbours; the remaining rows / shifts only for 18 / 26
    tile_pass<0>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 0, -1, -wz);
    tile_pass<0>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 0, -sg);
    if (CONN >= 18) {
        tile_pass<-1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 0, -1, -wz);
        tile_pass<1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 0, -1, -wz);
        tile_pass<-1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 0, -sg);
        tile_pass<1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 0, -sg);
        tile_pass<0>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, -1, -sg - wz);
        tile_pass<0>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 1, -sg + wz);
    }
    if (CONN == 26) {
        tile_pass<-1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, -1, -sg - wz);
        tile_pass<1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, -1, -sg - wz);
        tile_pass<-1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 1, -sg + wz);
        tile_pass<1>(s_mask, s_rb, s_st, s_par, s_act, nact, g.wz, vy, 1, 1, -sg + wz);
    }
    __syncthreads();
    // flatten; write parent[global run] = global id of the local root; give every local root a compact index
    for (u32 i = threadIdx.x; i < nloc; i += blockDim.x) {
        u32 r = sm_find_ro(s_par, i);          // read-only walk: a halving store could undo a neighbour's flatten
        s_par[i] = r;
        // x-slab segment of a local run id: last xl with segoff[xl] <= id (vx <= 16: 4 steps)
        u32 xi = 0, xr = 0;
#pragma unroll
        for (u32 s2 = 8; s2; s2 >>= 1) {
            if (xi + s2 < vx && s_segoff[xi + s2] <= i) xi += s2;
            if (xr + s2 < vx && s_segoff[xr + s2] <= r) xr += s2;
        }
        par[i + s_delta[xi]] = r + s_delta[xr];
        if (r == i) {
            const u32 idx = atomicAdd(&s_nroots, 1u);
            s_rootidx[i] = (unsigned short)idx;
            if (idx < TILE_ROOTS) s_rootrun[idx] = (unsigned short)i;
        }
    }
    __syncthreads();
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
    const bool same_mask = (omask == mask);
    for (u32 a = threadIdx.x; a < nact; a += blockDim.x) {
        const u32 e = s_act[a];
        const u32 i = e & 2047u, k = (e >> 11) & 1023u, yl = (e >> 21) & 63u, xl = e >> 27;
        const u32 m = s_mask[i];
        const u32 om = same_mask ? m : omask[((x0 + xl) * (u32)g.ny + y0 + yl) * g.wz + k];
        if (!om) continue;
        const u32 starts = s_st[i], rb = s_rb[i], z0 = k * 32;
        u32 ps = m & ~(m << 1);                // piece starts
        while (ps) {
            const int b0 = __ffs(ps) - 1;
            ps &= ps - 1;
            const u32 span = ~(m >> b0);
            int len = span ? (__ffs(span) - 1) : 32;
            if (len > 32 - b0) len = 32 - b0;
            const u32 pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b0;
            const u32 bits = om & pm;
            if (!bits) continue;
            const u32 li = rb + __popc(starts & mask_le((u32)b0)) - 1u;
            u32 *sl = s_slot + (u32)s_rootidx[s_par[li]] * TILE_SLOT;
            const u32 n = __popc(bits);
            atomicAdd(sl + 0, n);
            atomicAdd(sl + 1, n * xl);
            atomicAdd(sl + 2, n * yl);
            atomicAdd(sl + 3, n * z0 + bitpos_sum(bits));
            atomicMin(sl + 4, xl);
            atomicMin(sl + 5, yl);
            atomicMin(sl + 6, z0 + (u32)__ffs(bits) - 1u);
            atomicMax(sl + 7, xl);
            atomicMax(sl + 8, yl);
            atomicMax(sl + 9, z0 + 31u - (u32)__clz(bits));
        }
    }
    __syncthreads();
    for (u32 idx = threadIdx.x; idx < nroots; idx += blockDim.x) {
        const u32 *sl = s_slot + idx * TILE_SLOT;
        const u32 r = s_rootrun[idx];
        u32 xr = 0;
#pragma unroll
        for (u32 s2 = 8; s2; s2 >>= 1)
            if (xr + s2 < vx && s_segoff[xr + s2] <= r) xr += s2;
        TileRec rec;
        rec.root = r + s_delta[xr];
        rec.size = sl[0];
        rec.sx = (u64)sl[1] + (u64)sl[0] * x0;
        rec.sy = (u64)sl[2] + (u64)sl[0] * y0;
        rec.sz = sl[3];
        rec.minx = sl[4] + x0; rec.miny = sl[5] + y0; rec.minz = sl[6];
        rec.maxx = sl[7] + x0; rec.maxy = sl[8] + y0; rec.maxz = sl[9];
        rec.pad[0] = rec.pad[1] = 0;
        recs[s_recbase + idx] = rec;           // a local root with no ORIGINAL voxel (dilation) has size 0: harmless
    }
}

// merges the tile records into the per-component statistics (after the canonical ranks are known)
__global__ void __launch_bounds__(256) k_stats_merge(const TileRec *__restrict__ recs, const u32 *__restrict__ par,
                                                     const u32 *__restrict__ rank, CompStat *stat, const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u64 n = (u64)cnt->v[SYN_CNT_TILE_RECORDS];
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const TileRec r = recs[i];
        if (r.size == 0) continue;
        CompStat *s = stat + rank[par[r.root]];
        atomicAdd(&s->size, (u64)r.size);
        atomicAdd(&s->sx, r.sx);
        atomicAdd(&s->sy, r.sy);
        atomicAdd(&s->sz, r.sz);
        if (r.minx < ldcg_u32(&s->minx)) atomicMin(&s->minx, r.minx);
        if (r.miny < ldcg_u32(&s->miny)) atomicMin(&s->miny, r.miny);
        if (r.minz < ldcg_u32(&s->minz)) atomicMin(&s->minz, r.minz);
        if (r.maxx > ldcg_u32(&s->maxx)) atomicMax(&s->maxx, r.maxx);
        if (r.maxy > ldcg_u32(&s->maxy)) atomicMax(&s->maxy, r.maxy);
        if (r.maxz > ldcg_u32(&s->maxz)) atomicMax(&s->maxz, r.maxz);
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
    __device__ __forceinline__ u32 get(u64 g) const
    {
        u32 r = uf_find_ro(par, (u32)g);
        par[g] = r;
        return r == (u32)g;
    }
};
// rank[g] = canonical index for roots; also initialises the root's CompStat record
struct RankDst {
    u32 *rank;
    CompStat *stat;
    __device__ __forceinline__ void put(u64 g, u32 ex, u32 isroot) const
    {
        rank[g] = ex;
        if (isroot) {
            CompStat s;
            s.size = s.sx = s.sy = s.sz = 0;
            s.minx = s.miny = s.minz = 0xFFFFFFFFu;
            s.maxx = s.maxy = s.maxz = 0;
            s.keep = 0;
            s.row = 0;
            stat[ex] = s;
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
__global__ void __launch_bounds__(256) k_stats(const u32 *__restrict__ lmask, const u32 *__restrict__ omask,
                                               const u32 *__restrict__ runbase, const u32 *__restrict__ par,
                                               const u32 *__restrict__ rank, CompStat *stat, Geom g, TileGeom tg,
                                               const u32 *__restrict__ tileflag,
                                               const u32 *__restrict__ actw, const Counts *cnt)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
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
        u32 carry = k ? (lmask[w - 1] >> 31) : 0u;
        u32 starts = run_starts(m, carry);
        u32 rb = runbase[w];
        u32 ps = m & ~(m << 1);  // piece starts (no carry: a continuing run still starts a piece here)
        while (ps) {
            int b0 = __ffs(ps) - 1;
            ps &= ps - 1;
            u32 span = ~(m >> b0);                       // zero bits above the piece
            int len = span ? (__ffs(span) - 1) : 32;     // span==0 only when b0==0 and m is all ones
            if (len > 32 - b0) len = 32 - b0;
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
            if (x < ldcg_u32(&s->minx)) atomicMin(&s->minx, x);
            if (y < ldcg_u32(&s->miny)) atomicMin(&s->miny, y);
            if (zlo < ldcg_u32(&s->minz)) atomicMin(&s->minz, zlo);
            if (x > ldcg_u32(&s->maxx)) atomicMax(&s->maxx, x);
            if (y > ldcg_u32(&s->maxy)) atomicMax(&s->maxy, y);
            if (zhi > ldcg_u32(&s->maxz)) atomicMax(&s->maxz, zhi);
        }
    }
}

// ===========================================================================
// K6: dust decision + compaction of the segment table (scan #3 over components).
// keep = size >= dust  OR  the component touches one of the six chunk faces
// (continuations are never filtered: proc/tasks.py:45-51).
// ===========================================================================
struct KeepSrc {
    CompStat *stat;
    const Counts *cnt;
    i64 dust;
    u32 mx, my, mz;  // nx-1, ny-1, nz-1
    __device__ __forceinline__ u64 n() const
    {
        return (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) ? 0 : (u64)cnt->v[SYN_CNT_COMPONENTS];
    }
    __device__ __forceinline__ u32 get(u64 c) const
    {
        const CompStat s = stat[c];
        bool face = s.minx == 0 || s.miny == 0 || s.minz == 0 || s.maxx == mx || s.maxy == my || s.maxz == mz;
        u32 keep = (s.size > 0) && (face || (i64)s.size >= dust);
        stat[c].keep = keep;
        return keep;
    }
};

// centroid exactly as _describe.cpp:54-61: float32(sum) / float32(size), round half away
__device__ __forceinline__ i64 centroid_f32(u64 sum, u64 size)
{
    float q = __fdiv_rn(__ll2float_rn((i64)sum), __ll2float_rn((i64)size));
    return (i64)roundf(q);
}

struct TableDst {
    CompStat *stat;
    i64 *table;
    i64 max_segs;
    i64 ox, oy, oz;
    __device__ __forceinline__ void put(u64 c, u32 rowi, u32 keep) const
    {
        if (!keep) return;
        stat[c].row = rowi;
        if ((i64)rowi >= max_segs) return;
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
};

// final per-run label: canonical id if kept else 0; also tracks the largest kept id
__global__ void __launch_bounds__(256) k_run_labels(const u32 *__restrict__ par, const u32 *__restrict__ rank,
                                                    const CompStat *__restrict__ stat, u32 *__restrict__ runlabel,
                                                    Counts *cnt, i64 max_segs)
{
    if (cnt->v[SYN_CNT_ERRFLAGS] & SYN_EF_RUNS) return;
    const u64 nruns = (u64)cnt->v[SYN_CNT_RUNS];
    u32 mx = 0;
    for (u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x; g < nruns; g += (u64)gridDim.x * blockDim.x) {
        u32 c = rank[par[g]];
        u32 lab = stat[c].keep ? c + 1 : 0;
        runlabel[g] = lab;
        mx = max(mx, lab);
    }
    mx = __reduce_max_sync(SYN_FULL, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax((u64 *)&cnt->v[SYN_CNT_MAX_LABEL], (u64)mx);
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt->v[SYN_CNT_KEPT] > max_segs)
        atomicOr((u64 *)&cnt->v[SYN_CNT_ERRFLAGS], (u64)SYN_EF_SEGS);
}

}  // namespace syn
