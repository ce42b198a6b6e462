// common.cuh -- shared device helpers for libsynaptor_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/synaptor_b200.h"

#ifndef __CUDA_ARCH__
#define SYN_HOST 1
#endif

#define SYN_FULL 0xFFFFFFFFu

// error flag bits (counts[SYN_CNT_ERRFLAGS])
#define SYN_EF_RUNS 1
#define SYN_EF_SEGS 2
#define SYN_EF_TABLE 4
#define SYN_EF_PAIRS 8

namespace syn {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

// ---------------------------------------------------------------------------
// Volume geometry.  The 1-bit-per-voxel masks are stored row by row: a "row" is
// one (x, y) pair, `wz` 32-bit words per row, bit b of word k <-> z = 32k + b;
// bits at z >= nz are always zero.  Word index = row * wz + k fits in 32 bits
// (checked on the host).
// ---------------------------------------------------------------------------
struct Geom {
    i64 nx, ny, nz;
    i64 nvox;
    u32 rows;    // nx * ny
    u32 wz;      // ceil(nz / 32)
    u32 nwords;  // rows * wz
    u32 cpr;     // 128-voxel chunks per row = ceil(wz / 4)
    u32 nchunks; // rows * cpr
    u32 wzs;     // log2(wz) when wz is a power of two, else 32
    u32 nys;     // log2(ny) when ny is a power of two, else 32
};

// word index -> (row, word in row); row -> (x, y).  Shifts when the extents are powers of two.
__device__ __forceinline__ void split_word(const Geom &g, u32 w, u32 &row, u32 &k)
{
    if (g.wzs < 32) { row = w >> g.wzs; k = w & (g.wz - 1u); }
    else { row = w / g.wz; k = w - row * g.wz; }
}
__device__ __forceinline__ void split_row(const Geom &g, u32 row, u32 &x, u32 &y)
{
    if (g.nys < 32) { x = row >> g.nys; y = row & ((u32)g.ny - 1u); }
    else { x = row / (u32)g.ny; y = row - x * (u32)g.ny; }
}

// tiling of the (x, y) row grid for the shared-memory labelling pass (k_tile_ccl)
struct TileGeom {
    u32 tx, ty;          // rows per tile along x and y, powers of two (tx == 0: tiling disabled)
    u32 tiles_x, tiles_y;
    u32 ntiles;
    u32 txs, tys;        // log2(tx), log2(ty)
    u32 diag;            // 18 / 26 connectivity: the last row of a tile (y + 1 on a tile border) is a border row too
};

// per-component accumulator, one 64-byte record (two 32 B sectors)
struct __align__(16) CompStat {
    u64 size, sx, sy, sz;
    u32 minx, miny, minz, maxx, maxy, maxz;
    u32 keep;   // dust decision
    u32 row;    // row in the compacted segment table
};
static_assert(sizeof(CompStat) == 64, "CompStat must be 64 bytes");

// partial statistics of one tile-local component (written by the tile pass, merged per canonical component)
struct __align__(16) TileRec {
    u32 root;   // global run id of the tile-local root
    u32 size;
    u64 sx, sy, sz;
    u32 minx, miny, minz, maxx, maxy, maxz;
    u32 pad[2];
};
static_assert(sizeof(TileRec) == 64, "TileRec must be 64 bytes");

// slot of the (label, base id) pair table: 16-byte key claimed with one 128-bit CAS
struct __align__(16) PairKey {
    u64 base;
    u64 label;  // 0xFFFF...F = empty
};

// device-side counters at the head of the workspace
struct Counts {
    i64 v[SYN_NCOUNTS];
};

// phase timers of instrumented builds (make TIMING=1): thread 0 of a CTA adds the cycles since the previous
// mark to counts[SYN_CNT_DBG0 + slot]
#ifdef SYN_PHASE_TIMING
#define SYN_TIMER_INIT() long long syn_t0__ = clock64()
#define SYN_TIMER_MARK(cnt, slot)                                                              \
    do {                                                                                       \
        if (threadIdx.x == 0) {                                                                \
            const long long t__ = clock64();                                                   \
            atomicAdd((u64 *)&(cnt)->v[SYN_CNT_DBG0 + (slot)], (u64)(t__ - syn_t0__));         \
            syn_t0__ = t__;                                                                    \
        }                                                                                      \
    } while (0)
#else
#define SYN_TIMER_INIT() do { } while (0)
#define SYN_TIMER_MARK(cnt, slot) do { } while (0)
#endif

// open-addressing table of (label, base id) pairs in HBM (finalize.cuh).  One 32-byte slot = one L2 sector:
// the 16-byte key (claimed with one 128-bit CAS), the count and the first position, so that a probe reads the
// key AND the current first position with one request.
struct __align__(32) PairSlot {
    PairKey key;     // all-ones = empty
    u64 count;
    u64 first;       // smallest padded bit position (row * wz * 32 + z) of the pair
};
static_assert(sizeof(PairSlot) == 32, "PairSlot must be 32 bytes");

struct PairTable {
    PairSlot *slots; // [cap]
    u32 capmask;     // cap - 1 (cap is a power of two)
    Counts *cnt;
};

__host__ __device__ __forceinline__ u32 div_up_u32(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

// bits 0..b inclusive
__device__ __forceinline__ u32 mask_le(u32 b) { return (2u << b) - 1u; }
// bits 0..b-1
__device__ __forceinline__ u32 mask_lt(u32 b) { return (1u << b) - 1u; }

// run-start bits of a mask word given the top bit of the previous word in the row
__device__ __forceinline__ u32 run_starts(u32 m, u32 carry) { return m & ~((m << 1) | carry); }

// position of the (n+1)-th set bit of w (n < popc(w)): 5-step binary search on popcounts
__device__ __forceinline__ u32 select_bit(u32 w, u32 n)
{
    u32 pos = 0;
#pragma unroll
    for (u32 s = 16; s; s >>= 1) {
        const u32 c = __popc(w & ((1u << (pos + s)) - 1u));   // set bits below pos+s
        if (c <= n) pos += s;
    }
    return pos;
}

// sum of the positions of the set bits of w
__device__ __forceinline__ u32 bitpos_sum(u32 w)
{
    return __popc(w & 0xAAAAAAAAu) + (__popc(w & 0xCCCCCCCCu) << 1) + (__popc(w & 0xF0F0F0F0u) << 2) +
           (__popc(w & 0xFF00FF00u) << 3) + (__popc(w & 0xFFFF0000u) << 4);
}

__device__ __forceinline__ u32 ldcg_u32(const u32 *p) { return __ldcg(p); }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int d)
{
    return ((u64)__shfl_xor_sync(SYN_FULL, (u32)(v >> 32), d) << 32) | __shfl_xor_sync(SYN_FULL, (u32)v, d);
}

// ---------------------------------------------------------------------------
// Lock-free union-find on a u32 parent array in global memory.  Links always go
// from the larger to the smaller index (atomicMin), so the root of a set is its
// smallest member: with run ids in raster order that is the raster-first run,
// which is what makes the final numbering canonical.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 uf_find(u32 *par, u32 x)
{
    u32 p = ldcg_u32(par + x);
    while (p != x) {
        u32 gp = ldcg_u32(par + p);
        if (gp != p) par[x] = gp;  // path halving; always an ancestor, always smaller
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ u32 uf_find_ro(const u32 *par, u32 x)
{
    u32 p = ldcg_u32(par + x);
    while (p != x) {
        x = p;
        p = ldcg_u32(par + x);
    }
    return x;
}

__device__ __forceinline__ void uf_union(u32 *par, u32 a, u32 b)
{
    while (true) {
        a = uf_find(par, a);
        b = uf_find(par, b);
        if (a == b) return;
        if (a > b) { u32 t = a; a = b; b = t; }
        u32 old = atomicMin(par + b, a);
        if (old == b) return;
        b = old;  // b was no longer a root: keep merging a with its (former) parent
    }
}

// ---------------------------------------------------------------------------
// TMA bulk copy global -> shared with mbarrier completion (cp.async.bulk, sm_90+): the in-flight bytes live in
// shared memory instead of registers, so a streaming kernel can keep far more of them outstanding.
// ---------------------------------------------------------------------------
// 256-bit global store (sm_100: STG.256).  The address must be 32-byte aligned.
__device__ __forceinline__ void st_256(u32 *p, const u32 (&v)[8])
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    u32 done;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------
// 128-bit compare-and-swap (atom.global.cas.b128, sm_90+)
// ---------------------------------------------------------------------------
__device__ __forceinline__ PairKey cas128(PairKey *addr, PairKey cmp, PairKey val)
{
    PairKey old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.global.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.base), "=l"(old.label)
        : "l"(cmp.base), "l"(cmp.label), "l"(val.base), "l"(val.label), "l"(addr)
        : "memory");
    return old;
}

__device__ __forceinline__ u64 mix64(u64 a, u64 b)
{
    u64 h = b * 0x9E3779B97F4A7C15ull;
    h ^= a * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan (blockDim.x == SCAN_THREADS)
// ---------------------------------------------------------------------------
// 256 threads x 16 consecutive items: 8 CTAs per SM are resident, so a whole 512^3 mask (4.2 M words = 1024
// tiles) is scanned in ONE round of CTAs and a block waiting in the look-back only stalls its own 8 warps
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u32 warp_incl_scan(u32 v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(SYN_FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// returns the exclusive prefix of v over the block; *total = block sum.
// smem: 33 u32 of shared scratch.  Ends with a __syncthreads().
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *smem, u32 *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    u32 inc = warp_incl_scan(v, lane);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u32 w = lane < nw ? smem[lane] : 0;
        u32 wi = warp_incl_scan(w, lane);
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    u32 res = smem[wid] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

}  // namespace syn
