// cabi.cu -- extern "C" entry points of libsynaptor_b200.so (see include/synaptor_b200.h)
// and the host-side orchestration of the kernel pipeline.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdint>

#include "aux.cuh"
#include "ccl.cuh"
#include "finalize.cuh"
#include "merge.cuh"

using namespace syn;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(SYN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

std::atomic<long long> g_launches{0};

#define CK_LAUNCH()            \
    do {                       \
        g_launches.fetch_add(1, std::memory_order_relaxed); \
        CK(cudaGetLastError()); \
    } while (0)

// ---- optional per-kernel profiling with CUDA events on the launching stream -------------------
// While enabled, syn_chunk_ccs records an event after every kernel; interval i is named after the kernel that
// ended at mark i + 1.
constexpr int PROF_MAX = 40;
bool g_prof_on = false;
cudaEvent_t g_prof_ev[PROF_MAX];
const char *g_prof_name[PROF_MAX];
int g_prof_n = 0;
bool g_prof_created = false;
bool g_prof_valid = false;

std::mutex g_prof_mu;     // the profiling hooks are for one measuring thread; the lock only keeps the state sane

int prof_mark(const char *name, cudaStream_t st)
{
    if (!g_prof_on) return SYN_OK;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_n >= PROF_MAX) return SYN_OK;
    g_prof_name[g_prof_n] = name;
    CK(cudaEventRecord(g_prof_ev[g_prof_n], st));
    ++g_prof_n;
    return SYN_OK;
}
#define PROF(name)                                   \
    do {                                             \
        int prc__ = prof_mark((name), st);           \
        if (prc__ != SYN_OK) return prc__;           \
    } while (0)

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int make_geom(i64 nx, i64 ny, i64 nz, Geom *g)
{
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(SYN_ERR_INVALID, "empty or negative shape (%lld,%lld,%lld)",
                                                  (long long)nx, (long long)ny, (long long)nz);
    i64 wz = (nz + 31) / 32;
    i64 rows = nx * ny;
    i64 nwords = rows * wz;
    i64 cpr = (wz + 3) / 4;
    if (rows >= (1ll << 31) || nwords >= (1ll << 32) - 8192 || rows * cpr >= (1ll << 32) - 8192 ||
        nx >= (1ll << 31) || ny >= (1ll << 31) || nz >= (1ll << 31))
        return fail(SYN_ERR_INVALID, "chunk too large for 32-bit word indexing (%lld,%lld,%lld)", (long long)nx,
                    (long long)ny, (long long)nz);
    g->nx = nx; g->ny = ny; g->nz = nz;
    g->nvox = nx * ny * nz;
    g->rows = (u32)rows;
    g->wz = (u32)wz;
    g->nwords = (u32)nwords;
    g->cpr = (u32)cpr;
    g->nchunks = (u32)(rows * cpr);
    g->wzs = 32;
    g->nys = 32;
    for (u32 b = 0; b < 31; ++b) {
        if ((i64)1 << b == wz) g->wzs = b;
        if ((i64)1 << b == ny) g->nys = b;
    }
    return SYN_OK;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

u32 pow2_at_least(i64 v)
{
    u32 p = 1024;
    while ((i64)p < v && p < (1u << 31)) p <<= 1;
    return p;
}

// workspace carving (must match syn_ccs_workspace_bytes)
TileGeom make_tiles(const Geom &g)
{
    TileGeom t{0, 0, 0, 0, 0, 0, 0, 0};
    if (g.wz > TILE_WORDS / 2) return t;          // one row already fills a tile: no shared-memory pass
    u32 rpt = TILE_WORDS / g.wz;                    // rows per tile, >= 2
    u32 ty = 1;
    while (ty * 2 <= 64 && (ty * 2) * (ty * 2) <= 2 * rpt) ty *= 2;
    u32 tx = 1;                                      // power of two as well: tile coordinates by shifts
    while (tx * 2 <= TILE_MAXTX && tx * 2 * ty <= rpt) tx *= 2;
    t.tx = tx;
    t.ty = ty;
    for (u32 b = 0; b < 8; ++b) {
        if (1u << b == tx) t.txs = b;
        if (1u << b == ty) t.tys = b;
    }
    t.tiles_x = (u32)((g.nx + tx - 1) / tx);
    t.tiles_y = (u32)((g.ny + ty - 1) / ty);
    t.ntiles = t.tiles_x * t.tiles_y;
    return t;
}

struct Workspace {
    TileGeom tg;
    u32 *tileflag;
    Counts *cnt;
    u32 *lmask, *omask, *runbase;
    u32 *startw;                           // stored run-start words (multi-label variant only)
    u32 *actw;                             // compact list of non-empty mask words
    u32 *bndw;                             // non-empty words on a tile border
    u64 *scanstate;                        // 5 look-back state arrays + tickets, zeroed per call
    u32 *scan_tickets;
    size_t scan_off[5];
    size_t zero_bytes;                     // head of the workspace that every call zeroes
    u32 *tileagg, *tilebase;               // run counts / run-id bases of the stream tiles (k_stream)
    size_t ts_cap;
    u32 *par, *rank, *runlabel, *complab;
    CompStat *stat;
    TileRec *recs;
    PairSlot *pslots;
    u32 *bcount, *bfill, *sorted_slot;     // pair ordering: bucket counts (nb + 1), fill cursors (nb), bucket order
    u32 nbuckets;
    u32 paircap;
    size_t total;
};


void carve(char *base, const Geom &g, i64 max_runs, i64 max_pairs, bool dilate, bool ml, Workspace *w)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += al(bytes); return p; };
    const size_t nw = (size_t)g.nwords + 8;
    // ---- everything that must be zero at the start of a call sits at the head of the workspace: one memset ----
    w->cnt = (Counts *)take(sizeof(Counts));
    // look-back state of the five scan instances (slot 0: mask words, 16 items; 1: runs, 8; 2: components, 2;
    // 3: pair buckets, 8; 4: face voxels, 8) + their tickets
    const u32 nb = max_pairs > 0 ? (u32)((((u64)g.nwords * 32) >> PB_SHIFT) + 1) : 0;
    const size_t nface = (size_t)(2 * (g.ny * g.nz + g.nx * g.nz + g.nx * g.ny));
    const size_t cap[5] = {nw / (SCAN_THREADS * 16) + 2, (size_t)max_runs / (SCAN_THREADS * 8) + 2,
                           (size_t)max_runs / (SCAN_THREADS * 2) + 2, (size_t)nb / (SCAN_THREADS * 8) + 2,
                           nface / (SCAN_THREADS * 8) + 2};
    size_t so = 0;
    for (int i = 0; i < 5; ++i) { w->scan_off[i] = so; so += cap[i]; }
    w->scanstate = (u64 *)take((so + 8) * 8);          // + 16 u32 tickets (two per slot)
    w->scan_tickets = w->scanstate ? (u32 *)(w->scanstate + so) : nullptr;
    w->tg = make_tiles(g);
    w->tileflag = (u32 *)take(((size_t)w->tg.ntiles + 8) * 4);
    w->paircap = 0;
    w->pslots = nullptr; w->bcount = w->bfill = w->sorted_slot = nullptr;
    w->nbuckets = nb;
    if (max_pairs > 0) {
        w->bcount = (u32 *)take(((size_t)nb + 1) * 4 + (size_t)nb * 4);   // bucket counts + fill cursors
        w->bfill = w->bcount ? w->bcount + nb + 1 : nullptr;
    }
    w->zero_bytes = off;
    // ---- the rest is fully written before it is read ----
    w->lmask = (u32 *)take(nw * 4);
    w->omask = dilate ? (u32 *)take(nw * 4) : w->lmask;
    w->startw = ml ? (u32 *)take(nw * 4) : nullptr;
    w->runbase = (u32 *)take(nw * 4);
    w->actw = (u32 *)take(nw * 4);
    w->bndw = (u32 *)take(nw * 4);
    w->ts_cap = (size_t)g.nchunks / TS_CHUNKS + 2;
    w->tileagg = (u32 *)take(w->ts_cap * 4);
    w->tilebase = (u32 *)take(w->ts_cap * 4);
    w->par = (u32 *)take((size_t)max_runs * 4);
    w->rank = (u32 *)take((size_t)max_runs * 4);
    w->runlabel = (u32 *)take((size_t)max_runs * 4);
    w->complab = (u32 *)take((size_t)max_runs * 4);
    w->stat = (CompStat *)take((size_t)max_runs * sizeof(CompStat));
    w->recs = (TileRec *)take((size_t)max_runs * sizeof(TileRec));
    if (max_pairs > 0) {
        w->paircap = pow2_at_least(2 * max_pairs);
        w->pslots = (PairSlot *)take((size_t)w->paircap * sizeof(PairSlot));
        w->sorted_slot = (u32 *)take((size_t)max_pairs * 4);
    }
    w->total = off;
}

// grid = (resident CTAs per SM for this kernel) x (SM count): one full wave, no tail.
// Function attributes and occupancy are per DEVICE: the cache is keyed by (device, kernel, threads, smem), so a
// process that drives several GPUs sets the opt-in shared-memory size on each of them.
std::mutex g_occ_mu;
struct OccEntry { int dev; const void *fn; int threads; size_t smem; int blocks; };
OccEntry g_occ[512];
int g_nocc = 0;

template <class K>
int occ_blocks(K kernel, int threads, size_t dyn_smem)
{
    const void *fn = (const void *)kernel;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_occ_mu);
    for (int i = 0; i < g_nocc; ++i)
        if (g_occ[i].dev == dev && g_occ[i].fn == fn && g_occ[i].threads == threads && g_occ[i].smem == dyn_smem)
            return g_occ[i].blocks;
    if (dyn_smem > 48 * 1024)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, dyn_smem) != cudaSuccess || nb < 1) nb = 1;
    if (g_nocc < 512) g_occ[g_nocc++] = OccEntry{dev, fn, threads, dyn_smem, nb};
    return nb;
}

template <class K>
int wave_grid(K kernel, int threads, i64 max_blocks, size_t dyn_smem = 0)
{
    i64 g = (i64)occ_blocks(kernel, threads, dyn_smem) * sm_count();
    if (max_blocks < 1) max_blocks = 1;
    return (int)(g < max_blocks ? g : max_blocks);
}

int grid_for(i64 items, int threads, int per_sm)
{
    i64 need = (items + threads - 1) / threads;
    i64 cap = (i64)sm_count() * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// single-pass scan #slot (0..4) with its own look-back state inside the workspace
template <int ITEMS, class Src, class Dst>
int launch_scan(const Workspace &W, int slot, Src src, Dst dst, i64 max_items, i64 *total_out, cudaStream_t st)
{
    // (the capacity of slot `slot` in carve() assumes exactly this ITEMS)
    const i64 ntiles = (max_items + SCAN_THREADS * ITEMS - 1) / (SCAN_THREADS * ITEMS);
    u64 *state = W.scanstate + W.scan_off[slot];
    u32 *ticket = W.scan_tickets + 2 * slot;
    k_scan_1p<Src, Dst, ITEMS><<<wave_grid(k_scan_1p<Src, Dst, ITEMS>, SCAN_THREADS, ntiles), SCAN_THREADS, 0, st>>>(
        src, dst, state, ticket, total_out);
    CK_LAUNCH();
    return SYN_OK;
}

// orders the pair table entries by first occurrence and emits them into the output arrays
int order_and_emit_pairs(const Geom &g, const Workspace &W, PairTable T, u32 *rows, u64 *cols, i64 *vals,
                         i64 max_pairs, cudaStream_t st)
{
    (void)g;
    k_pair_count<<<grid_for(W.paircap, 256, 8), 256, 0, st>>>(T, W.bcount);
    CK_LAUNCH();
    PROF("pair_count");
    int rc = launch_scan<8>(W, 3, U32Src{W.bcount, W.nbuckets}, U32PrefixDst{W.bcount}, W.nbuckets,
                            &W.cnt->v[SYN_CNT_PAIRS], st);
    if (rc != SYN_OK) return rc;
    PROF("pair_bucket_scan");
    k_pair_scatter<<<grid_for(W.paircap, 256, 8), 256, 0, st>>>(T, W.bcount, W.bfill, W.sorted_slot, max_pairs);
    CK_LAUNCH();
    PROF("pair_scatter");
    k_pair_emit<<<grid_for(max_pairs, 256, 8), 256, 0, st>>>(T, W.bcount, W.nbuckets, W.sorted_slot, rows, cols, vals,
                                                            max_pairs);
    CK_LAUNCH();
    PROF("pair_emit");
    return SYN_OK;
}

int finish_counts(const Workspace &W, i64 *counts_host, cudaStream_t st)
{
    if (!counts_host) return SYN_OK;
    CK(cudaMemcpyAsync(counts_host, W.cnt, sizeof(Counts), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    i64 ef = counts_host[SYN_CNT_ERRFLAGS];
    if (ef)
        return fail(SYN_ERR_CAPACITY,
                    "capacity exceeded (flags=%lld): runs=%lld kept=%lld table_slots=%lld pairs=%lld", (long long)ef,
                    (long long)counts_host[SYN_CNT_RUNS], (long long)counts_host[SYN_CNT_KEPT],
                    (long long)counts_host[SYN_CNT_TABLE_SLOTS], (long long)counts_host[SYN_CNT_PAIRS]);
    return SYN_OK;
}

template <int CONN>
int launch_union(const Workspace &W, const Geom &g, RunIdx R, bool bnd_list, i64 max_runs, cudaStream_t st)
{
    if (W.tg.tx) {
        (void)occ_blocks(k_tile_ccl<CONN>, TILE_THREADS, TILE_SMEM);     // opt-in shared memory, once per device
        k_tile_ccl<CONN><<<W.tg.ntiles, TILE_THREADS, TILE_SMEM, st>>>(W.lmask, W.omask, R, W.par, W.tileflag, W.recs, g,
                                                            W.tg, W.cnt, (u32)max_runs);
        CK_LAUNCH();
        PROF("tile_ccl");
    }
    if (CONN == 6 && bnd_list) {
        // tile borders from the compact border-word list (+ the tiles the shared-memory pass had to leave)
        k_union_bnd<<<wave_grid(k_union_bnd, 256, div_up_u32(g.nwords, 256 * 4)), 256, 0, st>>>(
            W.lmask, R, W.par, g, W.tg, W.bndw, W.tileflag, W.actw, W.cnt);
        CK_LAUNCH();
        PROF("union_bnd");
    } else if (bnd_list) {
        // 18 / 26 connectivity: the border-word list (incl. the rows whose diagonal neighbours cross a tile border)
        k_union_bnd_diag<CONN == 6 ? 18 : CONN><<<wave_grid(k_union_bnd_diag<CONN == 6 ? 18 : CONN>, 256,
                                                            div_up_u32(g.nwords, 256 * 4)), 256, 0, st>>>(
            W.lmask, R, W.par, g, W.tg, W.bndw, W.tileflag, W.actw, W.cnt);
        CK_LAUNCH();
        PROF("union_bnd_diag");
    } else {
        k_union_rows<CONN><<<wave_grid(k_union_rows<CONN>, 256, div_up_u32(g.nwords, 256)), 256, 0, st>>>(
            W.lmask, R, W.par, g, W.tg, W.tileflag, W.actw, W.cnt);
        CK_LAUNCH();
        PROF("union_rows");
    }
    return SYN_OK;
}

}  // namespace

extern "C" {

int syn_version(void) { return 100; }

const char *syn_last_error(void) { return g_err; }

int64_t syn_ccs_workspace_bytes(int64_t nx, int64_t ny, int64_t nz, int64_t max_runs, int64_t max_pairs)
{
    Geom g;
    if (make_geom(nx, ny, nz, &g) != SYN_OK) return SYN_ERR_INVALID;
    if (max_runs < 1) max_runs = 1;
    Workspace w;
    carve(nullptr, g, max_runs, max_pairs, true, true, &w);
    return (int64_t)w.total;
}

int syn_read_counts(const void *workspace, int64_t *counts_host, void *stream)
{
    if (!workspace || !counts_host) return fail(SYN_ERR_INVALID, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(counts_host, workspace, sizeof(Counts), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SYN_OK;
}

}  // extern "C"

namespace {

// One chunk_ccs call, split into the part that needs only the chunk itself (ccs_begin: everything up to the
// segment table) and the part that writes the foreground labels and counts the overlaps (ccs_finish).  In between
// a caller may run the cross-chunk merge and hand the final id map to ccs_finish, so that the label volume is
// written ONCE, already in merged global ids (proc/tasks.py:314-333 remap_ids_task folded into the label write).
struct CcsCall {
    // arguments
    const float *pred;
    float thresh;
    int connectivity, dil_k;
    i64 dust;
    i64 off[3];
    u32 *labels;
    i64 *seg_table;
    i64 max_segs;
    const u64 *base;
    u32 *prow;
    u64 *pcol;
    i64 *pval;
    i64 max_pairs;
    void *workspace;
    i64 ws_bytes, max_runs;
    i64 *counts_host;
    cudaStream_t st;
    const u64 *ml_base;     // base segmentation of the multi-label variant (cc_task(overlap_seg=...)), else null
    bool defer;             // table-relative labels (row + 1) instead of component numbers; run labels in finish
    bool share_sm;          // the caller runs other chunks' kernels next to this one: k_stream takes 2 CTAs per SM
    const u32 *lut;         // finish, deferred: final label = lut[lut_base[0] + row + 1] (null: row + 1)
    const i64 *lut_base;
    i64 lut_len;
    // derived
    Geom g;
    Workspace W;
    bool has_base, dilate, ml, vec_in, sector_out, fused_scan;
    RunIdx R;
    PairTable T;
};

int ccs_setup(CcsCall &c, int64_t nx, int64_t ny, int64_t nz, bool need_pairs_out)
{
    int rc = make_geom(nx, ny, nz, &c.g);
    if (rc != SYN_OK) return rc;
    if (!c.pred || !c.labels || !c.workspace) return fail(SYN_ERR_INVALID, "null pointer argument");
    if (c.connectivity != 6 && c.connectivity != 18 && c.connectivity != 26)
        return fail(SYN_ERR_INVALID, "connectivity must be 6, 18 or 26 (got %d)", c.connectivity);
    if (c.dil_k < 0 || c.dil_k > 32) return fail(SYN_ERR_INVALID, "dil_k must be in [0, 32] (got %d)", c.dil_k);
    if (c.max_runs < 1 || c.max_runs >= (1ll << 32) - 8192) return fail(SYN_ERR_INVALID, "bad max_runs");
    if (c.max_segs < 0 || (c.max_segs > 0 && !c.seg_table)) return fail(SYN_ERR_INVALID, "bad seg table arguments");
    c.has_base = c.base != nullptr;
    if (c.has_base && (c.max_pairs < 1 || (need_pairs_out && (!c.prow || !c.pcol || !c.pval))))
        return fail(SYN_ERR_INVALID, "base_seg given without pair output arrays");
    if (!c.has_base) c.max_pairs = 0;
    c.dilate = c.dil_k > 0;
    c.ml = c.ml_base != nullptr;
    if (c.ml && (c.dilate || c.connectivity != 6 || c.has_base))
        return fail(SYN_ERR_INVALID, "the multi-label variant is 6-connected, without dilation or pair counting");
    const Geom &g = c.g;
    Workspace &W = c.W;
    carve((char *)c.workspace, g, c.max_runs, c.max_pairs, c.dilate, c.ml, &W);
    if (c.ml) W.tg = TileGeom{0, 0, 0, 0, 0, 0, 0, 0};   // generic global path: no shared-memory tile pass
    W.tg.diag = c.connectivity > 6;                      // rows with diagonal neighbours across a tile border
    if ((i64)W.total > c.ws_bytes)
        return fail(SYN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %lld", W.total, (long long)c.ws_bytes);
    // Sector mode (nz % 8 == 0, aligned buffers): ONE streaming kernel reads the predictions and the base volume
    // and zero-fills the label sectors without foreground (16 B/voxel, HBM bound); without dilation it also
    // produces the tiled run-id prefix and the active / border word lists.  Otherwise: plain threshold kernel,
    // and the generic output pass at the end.
    c.vec_in = (nz % 4 == 0) && ((reinterpret_cast<uintptr_t>(c.pred) & 15) == 0);
    c.sector_out = !c.ml && (nz % 8 == 0) && c.vec_in && ((reinterpret_cast<uintptr_t>(c.labels) & 31) == 0) &&
                   (!c.has_base || (reinterpret_cast<uintptr_t>(c.base) & 15) == 0);
    static const bool no_fuse = getenv("SYN_NO_FUSED_SCAN") != nullptr;
    c.fused_scan = c.sector_out && !c.dilate && !no_fuse;
    c.R = RunIdx{W.runbase, c.fused_scan ? W.tilebase : nullptr, g.cpr, c.ml ? W.startw : nullptr, c.ml_base};
    c.T = PairTable{W.pslots, W.paircap ? W.paircap - 1 : 0, W.cnt};
    return SYN_OK;
}

int ccs_begin(CcsCall &c)
{
    const Geom &g = c.g;
    Workspace &W = c.W;
    cudaStream_t st = c.st;
    const float *pred = c.pred;
    u32 *labels_out = c.labels;
    const float thresh = c.thresh;
    const i64 max_runs = c.max_runs;
    const bool has_base = c.has_base, dilate = c.dilate, ml = c.ml, vec_in = c.vec_in, sector_out = c.sector_out,
               fused_scan = c.fused_scan;
    const i64 nz = g.nz;
    int rc;

    CK(cudaMemsetAsync(c.workspace, 0, W.zero_bytes, st));     // counters, scan state, tile flags, pair buckets
    g_prof_n = 0;
    PROF("start");

    const u64 *bs = c.base;
    if (sector_out) {
        const i64 ntiles = ((i64)g.nchunks + TS_CHUNKS - 1) / TS_CHUNKS;
        static const int stream_ctas = getenv("SYN_STREAM_CTAS") ? atoi(getenv("SYN_STREAM_CTAS")) : 0;   // per SM; 0 = occupancy
        // co-running (another chunk's labelling kernels on a second stream): two resident CTAs per SM instead of three
        // leave 100 KB of shared memory for a k_tile_ccl CTA next to them (a lone chunk loses 6 % with that)
        const int ctas_per_sm = stream_ctas ? stream_ctas : (c.share_sm && has_base ? 2 : 0);
        static const bool no_tma = getenv("SYN_NO_TMA") != nullptr;
        static const bool all_tma = getenv("SYN_ALL_TMA") != nullptr;    // A/B: predictions by TMA on the fused path too
        // bulk copies into shared memory need chunk c to start at voxel 128 c (nz % 128 == 0)
        const bool tma = !no_tma && (nz % 128 == 0);
#define SYN_STREAM(HB, FS, TB, TP, NST)                                                                            \
    k_stream<HB, FS, TB, TP, NST><<<ctas_per_sm > 0 ? (int)std::min<i64>((i64)std::min(ctas_per_sm,                \
                                        occ_blocks(k_stream<HB, FS, TB, TP, NST>, TS_THREADS,                       \
                                                   NST * ts_stage_bytes(TP, TB))) * sm_count(), ntiles)            \
                                                   : wave_grid(k_stream<HB, FS, TB, TP, NST>, TS_THREADS, ntiles,  \
                                                               NST * ts_stage_bytes(TP, TB)),                      \
                                   TS_THREADS, NST * ts_stage_bytes(TP, TB), st>>>(                                \
        pred, bs, labels_out, W.omask, W.runbase, W.tileagg, W.tilebase, W.actw, W.bndw, g, W.tg, thresh, W.cnt,   \
        (u32)max_runs)
        if (has_base && tma && all_tma) { if (fused_scan) SYN_STREAM(true, true, true, true, 2); else SYN_STREAM(true, false, true, true, 2); }
        else if (has_base && tma) { if (fused_scan) SYN_STREAM(true, true, true, false, 2); else SYN_STREAM(true, false, true, false, 2); }
        else if (has_base) { if (fused_scan) SYN_STREAM(true, true, false, false, 1); else SYN_STREAM(true, false, false, false, 1); }
        else if (tma && fused_scan) {
            // predictions through a deeper ring on large volumes: 6 x 16 KB stages, 2 CTAs/SM (512^3: 206 -> 198 us,
            // 1024^3: 1560 -> 1469 us); small volumes keep 4 stages and 3 CTAs/SM (256^3: 34 vs 39 us)
            static const int nb_nst = getenv("SYN_NB_NST") ? atoi(getenv("SYN_NB_NST")) : 0;      // A/B override
            const int nst = nb_nst ? nb_nst : (ntiles >= 4096 ? 6 : 4);
            if (nst == 6) SYN_STREAM(false, true, false, true, 6);
            else SYN_STREAM(false, true, false, true, 4);
        }
        else if (tma) SYN_STREAM(false, false, false, true, 4);
        else { if (fused_scan) SYN_STREAM(false, true, false, false, 1); else SYN_STREAM(false, false, false, false, 1); }
#undef SYN_STREAM
        CK_LAUNCH();
        PROF("stream");
    } else if (ml) {
        k_threshold_ml<<<wave_grid(k_threshold_ml, 256, div_up_u32(g.nwords, 8)), 256, 0, st>>>(
            pred, c.ml_base, W.omask, W.startw, g, thresh, (u64 *)&W.cnt->v[SYN_CNT_FG_VOXELS]);
        CK_LAUNCH();
        PROF("threshold_ml");
    } else {
        const i64 need = div_up_u32((u64)g.nchunks, 8 * 4);
        if (vec_in)
            k_threshold<true, 4><<<wave_grid(k_threshold<true, 4>, 256, need), 256, 0, st>>>(
                pred, W.omask, g, thresh, (u64 *)&W.cnt->v[SYN_CNT_FG_VOXELS]);
        else
            k_threshold<false, 4><<<wave_grid(k_threshold<false, 4>, 256, need), 256, 0, st>>>(
                pred, W.omask, g, thresh, (u64 *)&W.cnt->v[SYN_CNT_FG_VOXELS]);
        CK_LAUNCH();
        PROF("threshold");
    }
    if (dilate) {
        // k radius-1 passes, ping-pong between lmask and the (not yet used) run-prefix array; the last lands in lmask
        const u32 *src = W.omask;
        for (int i = 1; i <= c.dil_k; ++i) {
            u32 *dst = ((c.dil_k - i) & 1) ? W.runbase : W.lmask;
            k_dilate_step<<<grid_for(g.nwords, 256, 8), 256, 0, st>>>(src, dst, g);
            CK_LAUNCH();
            src = dst;
        }
        PROF("dilate");
    }
    const RunIdx &R = c.R;
    if (has_base) {      // (folding this into the streaming kernel cost it 10 % of its bandwidth: measured, reverted)
        k_pair_table_init<<<grid_for(W.paircap, 256, 8), 256, 0, st>>>(c.T);
        CK_LAUNCH();
        PROF("pair_init");
    }

    // --- scan #1: run ids (already done by the streaming kernel on the plain path) ----------------------
    if (!fused_scan) {
        RunStartSrc src{W.lmask, g.nwords, g.wz, ml ? W.startw : nullptr};
        rc = launch_scan<16>(W, 0, src, RunBaseDst{W.runbase}, g.nwords, &W.cnt->v[SYN_CNT_RUNS], st);
        if (rc != SYN_OK) return rc;
        PROF("run_scan");
        const int ntiles = (int)((g.nwords + SCAN_TILE - 1) / SCAN_TILE);
        k_compact_active<<<wave_grid(k_compact_active, SCAN_THREADS, ntiles), SCAN_THREADS, 0, st>>>(
            W.lmask, g.nwords, W.actw, W.bndw, g, W.tg, W.cnt);
        CK_LAUNCH();
        PROF("compact_active");
    }
    if (!W.tg.tx) {   // with tiling the tile pass writes every parent (and checks the run capacity)
        k_init_parents<<<wave_grid(k_init_parents, 256, div_up_u32(max_runs, 256)), 256, 0, st>>>(W.par, W.cnt,
                                                                                              (u32)max_runs);
        CK_LAUNCH();
    }

    // --- unions ------------------------------------------------------------------------------
    const bool bnd_list = W.tg.tx != 0;   // both producers of the active-word list also list the border words
    if (c.connectivity == 6) rc = launch_union<6>(W, g, R, bnd_list, max_runs, st);
    else if (c.connectivity == 18) rc = launch_union<18>(W, g, R, bnd_list, max_runs, st);
    else rc = launch_union<26>(W, g, R, bnd_list, max_runs, st);
    if (rc != SYN_OK) return rc;

    // --- scan #2: flatten + canonical rank -----------------------------------------------------
    {
        RootFlagSrc src{W.par, W.cnt};
        rc = launch_scan<8>(W, 1, src, RankDst{W.rank, W.stat}, max_runs, &W.cnt->v[SYN_CNT_COMPONENTS], st);
        if (rc != SYN_OK) return rc;
        PROF("flatten_rank");
    }

    // --- statistics: tile records -> components (flagged tiles, normally none, per piece) -----------------
    k_stats_merge<<<wave_grid(k_stats_merge, 256, div_up_u32(max_runs, 256)), 256, 0, st>>>(
        W.recs, W.par, W.rank, W.stat, W.lmask, W.omask, R, g, W.tg, W.tileflag, W.actw, W.cnt);
    CK_LAUNCH();
    PROF("stats_merge");

    // --- scan #3: dust decision + table ----------------------------------------------------------
    {
        KeepSrc src{W.stat, W.cnt, c.dust, (u32)(g.nx - 1), (u32)(g.ny - 1), (u32)(g.nz - 1)};
        rc = launch_scan<2>(W, 2, src,
                            TableDst{W.stat, W.complab, c.seg_table, c.max_segs, c.off[0], c.off[1], c.off[2], c.defer},
                            max_runs, &W.cnt->v[SYN_CNT_KEPT], st);
        if (rc != SYN_OK) return rc;
        PROF("dust_table");
    }
    return SYN_OK;
}

int ccs_run_labels(CcsCall &c)
{
    Workspace &W = c.W;
    cudaStream_t st = c.st;
    k_run_labels<<<wave_grid(k_run_labels, 256, div_up_u32(c.max_runs, 256 * 4)), 256, 0, st>>>(
        W.par, W.rank, W.complab, W.runlabel, W.cnt, c.max_segs, c.defer ? c.lut : nullptr, c.lut_base, c.lut_len);
    CK_LAUNCH();
    PROF("run_labels");
    return SYN_OK;
}

int ccs_finish(CcsCall &c)
{
    const Geom &g = c.g;
    Workspace &W = c.W;
    cudaStream_t st = c.st;
    const bool has_base = c.has_base;
    const RunIdx &R = c.R;
    const u64 *bs = c.base;
    u32 *labels_out = c.labels;
    int rc;
    // --- foreground sectors (+ overlaps) -------------------------------------------------------------
    if (c.sector_out) {
        const i64 needw = div_up_u32(g.nwords, SL_BATCH * SL_WARPS);   // one warp per batch of active words, at most
        if (has_base)
            k_sector_labels<true><<<wave_grid(k_sector_labels<true>, 256, needw), 256, 0, st>>>(
                W.lmask, W.omask, R, W.runlabel, labels_out, bs, g, c.T, W.actw, W.cnt);
        else
            k_sector_labels<false><<<wave_grid(k_sector_labels<false>, 256, needw), 256, 0, st>>>(
                W.lmask, W.omask, R, W.runlabel, labels_out, bs, g, c.T, W.actw, W.cnt);
        CK_LAUNCH();
        PROF("sector_labels");
    } else {
        const int grid = grid_for((i64)g.nchunks * 32 / 2 + 1, 256, 8);
        if (has_base)
            k_write_labels<false, true, 2><<<grid, 256, 0, st>>>(W.lmask, W.omask, R, W.runlabel, labels_out,
                                                                bs, g, c.T, W.cnt);
        else
            k_write_labels<false, false, 2><<<grid, 256, 0, st>>>(W.lmask, W.omask, R, W.runlabel, labels_out,
                                                                 bs, g, c.T, W.cnt);
        CK_LAUNCH();
        PROF("write_labels");
    }
    if (has_base) {
        rc = order_and_emit_pairs(g, W, c.T, c.prow, c.pcol, c.pval, c.max_pairs, st);
        if (rc != SYN_OK) return rc;
    }
    if (g_prof_on) g_prof_valid = true;
    return finish_counts(W, c.counts_host, st);
}

CcsCall make_call(const float *pred, float thresh, int connectivity, int dil_k, int64_t dust_thresh,
                  const int64_t *offset_xyz, uint32_t *labels_out, int64_t *seg_table, int64_t max_segs,
                  const uint64_t *base_seg, uint32_t *pair_rows, uint64_t *pair_cols, int64_t *pair_counts,
                  int64_t max_pairs, void *workspace, int64_t workspace_bytes, int64_t max_runs, int64_t *counts_host,
                  void *stream, const uint64_t *ml_base)
{
    CcsCall c{};
    c.pred = pred; c.thresh = thresh; c.connectivity = connectivity; c.dil_k = dil_k; c.dust = dust_thresh;
    for (int k = 0; k < 3; ++k) c.off[k] = offset_xyz ? offset_xyz[k] : 0;
    c.labels = labels_out; c.seg_table = (i64 *)seg_table; c.max_segs = max_segs;
    c.base = (const u64 *)base_seg; c.prow = pair_rows; c.pcol = (u64 *)pair_cols; c.pval = (i64 *)pair_counts;
    c.max_pairs = max_pairs; c.workspace = workspace; c.ws_bytes = workspace_bytes; c.max_runs = max_runs;
    c.counts_host = (i64 *)counts_host; c.st = (cudaStream_t)stream; c.ml_base = (const u64 *)ml_base;
    c.defer = false; c.share_sm = false; c.lut = nullptr; c.lut_base = nullptr; c.lut_len = 0;
    return c;
}

int chunk_ccs_impl(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int connectivity, int dil_k,
                   int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out, int64_t *seg_table,
                   int64_t max_segs, const uint64_t *base_seg, uint32_t *pair_rows, uint64_t *pair_cols,
                   int64_t *pair_counts, int64_t max_pairs, void *workspace, int64_t workspace_bytes,
                   int64_t max_runs, int64_t *counts_host, void *stream, const uint64_t *ml_base)
{
    CcsCall c = make_call(pred, thresh, connectivity, dil_k, dust_thresh, offset_xyz, labels_out, seg_table, max_segs,
                          base_seg, pair_rows, pair_cols, pair_counts, max_pairs, workspace, workspace_bytes, max_runs,
                          counts_host, stream, ml_base);
    int rc = ccs_setup(c, nx, ny, nz, true);
    if (rc != SYN_OK) return rc;
    rc = ccs_begin(c);
    if (rc != SYN_OK) return rc;
    rc = ccs_run_labels(c);
    if (rc != SYN_OK) return rc;
    return ccs_finish(c);
}

}  // namespace

extern "C" {

int syn_chunk_ccs(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int connectivity, int dil_k,
                  int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out, int64_t *seg_table,
                  int64_t max_segs, const uint64_t *base_seg, uint32_t *pair_rows, uint64_t *pair_cols,
                  int64_t *pair_counts, int64_t max_pairs, void *workspace, int64_t workspace_bytes,
                  int64_t max_runs, int64_t *counts_host, void *stream)
{
    return chunk_ccs_impl(pred, nx, ny, nz, thresh, connectivity, dil_k, dust_thresh, offset_xyz, labels_out, seg_table,
                          max_segs, base_seg, pair_rows, pair_cols, pair_counts, max_pairs, workspace, workspace_bytes,
                          max_runs, counts_host, stream, nullptr);
}

int syn_chunk_ccs_multilabel(const float *pred, const uint64_t *overlap_seg, int64_t nx, int64_t ny, int64_t nz,
                             float thresh, int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out,
                             int64_t *seg_table, int64_t max_segs, void *workspace, int64_t workspace_bytes,
                             int64_t max_runs, int64_t *counts_host, void *stream)
{
    if (!overlap_seg) return fail(SYN_ERR_INVALID, "null overlap_seg");
    return chunk_ccs_impl(pred, nx, ny, nz, thresh, 6, 0, dust_thresh, offset_xyz, labels_out, seg_table, max_segs,
                          nullptr, nullptr, nullptr, nullptr, 0, workspace, workspace_bytes, max_runs, counts_host,
                          stream, overlap_seg);
}

/* ---- the chunk pass in two halves, with the cross-chunk merge in between (see CcsCall) ---- */

int64_t syn_packed_slot_bytes(int64_t cap_rows, int64_t cap_entries)
{
    if (cap_rows < 1 || cap_entries < 1) return SYN_ERR_INVALID;
    return (int64_t)pk_slot_bytes(cap_rows, cap_entries);
}

int syn_chunk_ccs_begin(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int connectivity,
                        int dil_k, int64_t dust_thresh, const int64_t *offset_xyz, uint32_t *labels_out,
                        const uint64_t *base_seg, int64_t max_pairs, void *packed_slot, int64_t cap_rows,
                        int64_t cap_entries, int face_mask, void *workspace, int64_t workspace_bytes, int64_t max_runs,
                        void *stream)
{
    if (!packed_slot || cap_rows < 1 || cap_entries < 1) return fail(SYN_ERR_INVALID, "bad packed slot arguments");
    if ((reinterpret_cast<uintptr_t>(packed_slot) & 15) != 0) return fail(SYN_ERR_INVALID, "packed slot must be 16-byte aligned");
    i64 *hdr = (i64 *)packed_slot;
    CcsCall c = make_call(pred, thresh, connectivity, dil_k, dust_thresh, offset_xyz, labels_out, (int64_t *)(hdr + PK_HDR), cap_rows,
                          base_seg, nullptr, nullptr, nullptr, max_pairs, workspace, workspace_bytes, max_runs, nullptr,
                          stream, nullptr);
    c.defer = true;
    c.share_sm = (face_mask & SYN_BEGIN_SHARE_SM) != 0;
    int rc = ccs_setup(c, nx, ny, nz, false);
    if (rc != SYN_OK) return rc;
    rc = ccs_begin(c);
    if (rc != SYN_OK) return rc;
    // the chunk's face voxels in table-relative ids, ordered (face, position): what merge_continuations needs
    // of the label volume (proc/seg/continuation.py:87-133), without the label volume
    cudaStream_t st = c.st;
    const Geom &g = c.g;
    CK(cudaMemsetAsync(hdr, 0, PK_HDR * 8, st));
    FaceCtx F{c.W.lmask, c.W.omask, c.R, c.W.par, c.W.rank, c.W.complab, c.W.cnt, g, {0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}};
    const i64 area3[3] = {g.ny * g.nz, g.nx * g.nz, g.nx * g.ny};
    i64 nface = 0;
    for (int f = 0; f < 6; ++f) {
        if (area3[f / 2] >= (1ll << 31)) return fail(SYN_ERR_INVALID, "chunk face too large");
        F.start[f] = (u32)nface;
        F.area[f] = ((face_mask >> f) & 1) ? (u32)area3[f / 2] : 0u;
        nface += F.area[f];
    }
    if (nface >= (1ll << 32)) return fail(SYN_ERR_INVALID, "chunk faces too large");
    F.start[6] = (u32)nface;
    uint2 *entries = reinterpret_cast<uint2 *>(hdr + PK_HDR + cap_rows * SYN_SEG_NCOLS);
    if (nface == 0) {       // no wanted face (a single-chunk grid): only the row count travels
        k_pack_header_only<<<1, 32, 0, st>>>(hdr, c.W.cnt, cap_rows);
        CK_LAUNCH();
        return SYN_OK;
    }
    rc = launch_scan<8>(c.W, 4, FaceSrc{F}, FaceDst{F, hdr, entries, cap_entries, cap_rows}, nface, &hdr[PK_NENT], st);
    if (rc != SYN_OK) return rc;
    PROF("face_entries");
    return SYN_OK;
}

int syn_chunk_ccs_finish(const float *pred, int64_t nx, int64_t ny, int64_t nz, int connectivity, int dil_k,
                         uint32_t *labels_out, const uint64_t *base_seg, uint32_t *pair_rows, uint64_t *pair_cols,
                         int64_t *pair_counts, int64_t max_pairs, int64_t cap_rows, const uint32_t *final_lut,
                         const int64_t *final_lut_base_dev, int64_t final_lut_len, void *workspace,
                         int64_t workspace_bytes, int64_t max_runs, int64_t *counts_host, void *stream)
{
    if (final_lut && (!final_lut_base_dev || final_lut_len < 1)) return fail(SYN_ERR_INVALID, "bad final lut arguments");
    CcsCall c = make_call(pred, 0.f, connectivity, dil_k, 0, nullptr, labels_out, nullptr, 0, base_seg, pair_rows,
                          pair_cols, pair_counts, max_pairs, workspace, workspace_bytes, max_runs, counts_host, stream,
                          nullptr);
    c.defer = true;
    c.max_segs = 0;
    c.lut = final_lut; c.lut_base = (const i64 *)final_lut_base_dev; c.lut_len = final_lut_len;
    int rc = ccs_setup(c, nx, ny, nz, true);
    if (rc != SYN_OK) return rc;
    c.max_segs = cap_rows;
    rc = ccs_run_labels(c);
    if (rc != SYN_OK) return rc;
    return ccs_finish(c);
}

int64_t syn_overlap_workspace_bytes(int64_t nx, int64_t ny, int64_t nz, int64_t max_pairs)
{
    return syn_ccs_workspace_bytes(nx, ny, nz, 1, max_pairs < 1 ? 1 : max_pairs);
}

int syn_count_overlaps(const uint32_t *seg1, const uint64_t *seg2, int64_t nx, int64_t ny, int64_t nz,
                       uint32_t *pair_rows, uint64_t *pair_cols, int64_t *pair_counts, int64_t max_pairs,
                       void *workspace, int64_t workspace_bytes, int64_t *counts_host, void *stream)
{
    Geom g;
    int rc = make_geom(nx, ny, nz, &g);
    if (rc != SYN_OK) return rc;
    if (!seg1 || !seg2 || !workspace || !pair_rows || !pair_cols || !pair_counts || max_pairs < 1)
        return fail(SYN_ERR_INVALID, "null pointer / bad capacity");
    Workspace W;
    carve((char *)workspace, g, 1, max_pairs, true, true, &W);
    if ((i64)W.total > workspace_bytes)
        return fail(SYN_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %lld", W.total,
                    (long long)workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(workspace, 0, W.zero_bytes, st));
    PairTable T{W.pslots, W.paircap - 1, W.cnt};
    k_pair_table_init<<<grid_for(W.paircap, 256, 8), 256, 0, st>>>(T);
    CK_LAUNCH();
    k_overlap_scan<<<grid_for((i64)g.nwords * 32, 256, 8), 256, 0, st>>>(seg1, (const u64 *)seg2, g, T, W.cnt);
    CK_LAUNCH();
    rc = order_and_emit_pairs(g, W, T, pair_rows, (u64 *)pair_cols, (i64 *)pair_counts, max_pairs, st);
    if (rc != SYN_OK) return rc;
    return finish_counts(W, (i64 *)counts_host, st);
}

int syn_extract_faces(const uint32_t *labels, int64_t nx, int64_t ny, int64_t nz, uint32_t *faces_out, void *stream)
{
    if (!labels || !faces_out || nx <= 0 || ny <= 0 || nz <= 0) return fail(SYN_ERR_INVALID, "bad arguments");
    const i64 total = 2 * (ny * nz + nx * nz + nx * ny);
    k_extract_faces<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(labels, nx, ny, nz, faces_out);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_transpose_zyx(const void *src_zyx, void *dst_xyz, int64_t nx, int64_t ny, int64_t nz, int elem_bytes,
                      void *stream)
{
    if (!src_zyx || !dst_xyz || nx <= 0 || ny <= 0 || nz <= 0) return fail(SYN_ERR_INVALID, "bad arguments");
    if (elem_bytes != 4 && elem_bytes != 8) return fail(SYN_ERR_INVALID, "element size must be 4 or 8 bytes");
    if (ny > 65535 || (nz + 31) / 32 > 65535) return fail(SYN_ERR_INVALID, "shape too large for the transpose grid");
    const dim3 grid((unsigned)((nx + 31) / 32), (unsigned)((nz + 31) / 32), (unsigned)ny), block(32, 8);
    if (elem_bytes == 4)
        k_transpose_zyx<u32><<<grid, block, 0, (cudaStream_t)stream>>>((const u32 *)src_zyx, (u32 *)dst_xyz, nx, ny, nz);
    else
        k_transpose_zyx<u64><<<grid, block, 0, (cudaStream_t)stream>>>((const u64 *)src_zyx, (u64 *)dst_xyz, nx, ny, nz);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_relabel_lut(uint32_t *labels, int64_t n, const uint32_t *lut, int64_t lut_len, void *stream)
{
    if (!labels || n < 0 || (lut_len > 0 && !lut)) return fail(SYN_ERR_INVALID, "bad arguments");
    if (n == 0 || lut_len <= 0) return SYN_OK;
    k_relabel_lut<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(labels, n, lut, lut_len);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_max_u32(const uint32_t *labels, int64_t n, uint32_t *out_dev, void *stream)
{
    if (!labels || !out_dev || n < 0) return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(out_dev, 0, 4, st));
    if (n == 0) return SYN_OK;
    k_max_u32<<<grid_for(n, 256, 8), 256, 0, st>>>(labels, n, out_dev);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_segment_stats(const uint32_t *labels, int64_t nx, int64_t ny, int64_t nz, int64_t max_id, int64_t *stats,
                      void *stream)
{
    Geom g;
    int rc = make_geom(nx, ny, nz, &g);
    if (rc != SYN_OK) return rc;
    if (!labels || !stats || max_id < 0) return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CompStat *stat = nullptr;
    CK(cudaMallocAsync(&stat, (size_t)(max_id + 1) * sizeof(CompStat), st));
    k_init_stats<<<grid_for(max_id + 1, 256, 8), 256, 0, st>>>(stat, max_id + 1);
    CK_LAUNCH();
    k_label_stats<<<grid_for((i64)g.nwords * 32, 256, 8), 256, 0, st>>>(labels, g, max_id, stat);
    CK_LAUNCH();
    k_stats_to_table<<<grid_for(max_id + 1, 256, 8), 256, 0, st>>>(stat, max_id + 1, (i64 *)stats);
    CK_LAUNCH();
    CK(cudaFreeAsync(stat, st));
    return SYN_OK;
}

int syn_face_pair_edges(const uint32_t *face_a, const uint32_t *face_b, int64_t n, uint32_t *edges_out,
                        int64_t max_edges, unsigned long long *n_edges_dev, void *stream)
{
    if (!face_a || !face_b || !edges_out || !n_edges_dev || n < 0 || max_edges < 0)
        return fail(SYN_ERR_INVALID, "bad arguments");
    if (n == 0) return SYN_OK;
    k_face_pair_edges<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(face_a, face_b, n, edges_out, max_edges,
                                                                          n_edges_dev);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_union_min_map(const uint32_t *edges, const unsigned long long *n_edges_dev, int64_t max_edges, int64_t n_ids,
                      uint32_t *map_out, void *stream)
{
    if (!edges || !n_edges_dev || !map_out || n_ids < 0 || max_edges < 0) return fail(SYN_ERR_INVALID, "bad arguments");
    if (n_ids == 0) return SYN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_iota_u32<<<grid_for(n_ids, 256, 8), 256, 0, st>>>(map_out, n_ids);
    CK_LAUNCH();
    if (max_edges > 0) {
        k_union_edges<<<grid_for(max_edges, 256, 8), 256, 0, st>>>(edges, n_edges_dev, max_edges, n_ids, map_out);
        CK_LAUNCH();
    }
    k_flatten<<<grid_for(n_ids, 256, 8), 256, 0, st>>>(map_out, n_ids);
    CK_LAUNCH();
    return SYN_OK;
}


int syn_threshold_dilate(const float *pred, int64_t nx, int64_t ny, int64_t nz, float thresh, int dil_k,
                         uint32_t *out01, void *stream)
{
    Geom g;
    int rc = make_geom(nx, ny, nz, &g);
    if (rc != SYN_OK) return rc;
    if (!pred || !out01) return fail(SYN_ERR_INVALID, "null pointer argument");
    if (dil_k < 0 || dil_k > 32) return fail(SYN_ERR_INVALID, "dil_k must be in [0, 32] (got %d)", dil_k);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nw = (size_t)g.nwords + 8;
    u32 *buf = nullptr;
    CK(cudaMallocAsync(&buf, 2 * nw * 4 + 64, st));
    u32 *m0 = buf, *m1 = buf + nw;
    u64 *fg = (u64 *)(buf + 2 * nw);
    CK(cudaMemsetAsync(fg, 0, 8, st));
    const bool vec_in = (nz % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0);
    const int grid = grid_for((i64)g.nchunks * 32 / 4 + 1, 256, 8);
    if (vec_in) k_threshold<true, 4><<<grid, 256, 0, st>>>(pred, m0, g, thresh, fg);
    else k_threshold<false, 4><<<grid, 256, 0, st>>>(pred, m0, g, thresh, fg);
    CK_LAUNCH();
    const u32 *res = m0;
    for (int i = 1; i <= dil_k; ++i) {      // k radius-1 passes, ping-pong
        u32 *dst = (res == m0) ? m1 : m0;
        k_dilate_step<<<grid_for(g.nwords, 256, 8), 256, 0, st>>>(res, dst, g);
        CK_LAUNCH();
        res = dst;
    }
    k_expand_mask<<<grid_for((i64)g.nwords * 32, 256, 8), 256, 0, st>>>(res, g, out01);
    CK_LAUNCH();
    CK(cudaFreeAsync(buf, st));
    return SYN_OK;
}

int syn_unique_u64(const uint64_t *vals, int64_t n, uint64_t *out, int64_t cap, unsigned long long *count_dev,
                   void *stream)
{
    if (!vals || !out || !count_dev || n < 0 || cap < 1) return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(count_dev, 0, 8, st));
    if (n == 0) return SYN_OK;
    const u32 tcap = pow2_at_least(2 * cap);
    u64 *table = nullptr;
    CK(cudaMallocAsync(&table, ((size_t)tcap + 1) * 8, st));
    k_fill_u64<<<grid_for(tcap, 256, 8), 256, 0, st>>>(table, tcap, ~0ull);
    CK_LAUNCH();
    CK(cudaMemsetAsync(table + tcap, 0, 8, st));
    k_unique_u64<<<grid_for(n, 256, 8), 256, 0, st>>>((const u64 *)vals, n, table, tcap - 1, (u64 *)out, cap,
                                                     count_dev);
    CK_LAUNCH();
    CK(cudaFreeAsync(table, st));
    return SYN_OK;
}


int syn_merge_ccs_grid(const uint32_t *faces_all, const int64_t *tables_all, const int64_t *kept,
                       const int64_t *slot_of_chunk, const int64_t *grid_xyz, const int64_t *shape_xyz, int64_t maxk,
                       int64_t size_thr, uint32_t *edges_scratch, int64_t max_edges, unsigned long long *n_edges_dev,
                       uint32_t *gmap_out, int64_t *rows_scratch, int64_t *merged_out, uint32_t *fmap_out,
                       int64_t *n_merged_dev, void *stream)
{
    if (!faces_all || !tables_all || !kept || !slot_of_chunk || !grid_xyz || !shape_xyz || !edges_scratch ||
        !n_edges_dev || !gmap_out || !fmap_out || !n_merged_dev || maxk < 1 || max_edges < 1)
        return fail(SYN_ERR_INVALID, "bad arguments");
    const i64 gx = grid_xyz[0], gy = grid_xyz[1], gz = grid_xyz[2];
    const i64 nx = shape_xyz[0], ny = shape_xyz[1], nz = shape_xyz[2];
    if (gx < 1 || gy < 1 || gz < 1 || nx < 1 || ny < 1 || nz < 1) return fail(SYN_ERR_INVALID, "bad grid / shape");
    cudaStream_t st = (cudaStream_t)stream;
    // face buffer of one chunk: (axis 0 hi, axis 0 lo, axis 1 hi, axis 1 lo, axis 2 hi, axis 2 lo), continuation.py:57-59
    const i64 fsz[3] = {ny * nz, nx * nz, nx * ny};
    i64 foff[6], facebuf = 0;
    for (int f = 0; f < 6; ++f) { foff[f] = facebuf; facebuf += fsz[f / 2]; }
    const i64 nchunks = gx * gy * gz;
    i64 total = 0;
    for (i64 i = 0; i < nchunks; ++i) {
        if (kept[i] < 0 || kept[i] > maxk) return fail(SYN_ERR_INVALID, "kept[%lld] out of range", (long long)i);
        total += kept[i];
    }
    CK(cudaMemsetAsync(n_edges_dev, 0, 8, st));
    // edges of every interior face pair: the hi face of a chunk against the lo face of its +1 neighbour
    // (merge_ccs.py:69-114 visits exactly these pairs)
    for (i64 cx = 0; cx < gx; ++cx)
        for (i64 cy = 0; cy < gy; ++cy)
            for (i64 cz = 0; cz < gz; ++cz) {
                const i64 here = (cx * gy + cy) * gz + cz;
                const i64 idx[3] = {cx, cy, cz}, g3[3] = {gx, gy, gz};
                for (int axis = 0; axis < 3; ++axis) {
                    if (idx[axis] == g3[axis] - 1) continue;
                    i64 o[3] = {cx, cy, cz};
                    o[axis] += 1;
                    const i64 there = (o[0] * gy + o[1]) * gz + o[2];
                    const u32 *fa = faces_all + slot_of_chunk[here] * facebuf + foff[2 * axis];        // hi face
                    const u32 *fb = faces_all + slot_of_chunk[there] * facebuf + foff[2 * axis + 1];   // lo face
                    k_face_pair_edges<<<grid_for(fsz[axis], 256, 8), 256, 0, st>>>(fa, fb, fsz[axis], edges_scratch,
                                                                                  max_edges, n_edges_dev);
                    CK_LAUNCH();
                }
            }
    // union-find to the smallest id
    const i64 n_ids = total + 1;
    k_iota_u32<<<grid_for(n_ids, 256, 8), 256, 0, st>>>(gmap_out, n_ids);
    CK_LAUNCH();
    k_union_edges<<<grid_for(max_edges, 256, 8), 256, 0, st>>>(edges_scratch, n_edges_dev, max_edges, n_ids, gmap_out);
    CK_LAUNCH();
    k_flatten<<<grid_for(n_ids, 256, 8), 256, 0, st>>>(gmap_out, n_ids);
    CK_LAUNCH();
    // the table rows of all chunks in global id order (assign_ids.py:8-25: chunk after chunk, rows in table order)
    if (total > 0 && !rows_scratch) return fail(SYN_ERR_INVALID, "rows_scratch missing");
    i64 base = 0;
    for (i64 i = 0; i < nchunks; ++i) {
        if (kept[i])
            CK(cudaMemcpyAsync(rows_scratch + base * SYN_SEG_NCOLS, tables_all + slot_of_chunk[i] * maxk * SYN_SEG_NCOLS,
                               (size_t)kept[i] * SYN_SEG_NCOLS * 8, cudaMemcpyDeviceToDevice, st));
        base += kept[i];
    }
    return syn_merge_seg_tables(rows_scratch, total, gmap_out, size_thr, merged_out, fmap_out, n_merged_dev, stream);
}

namespace {
struct MergeWs {
    u32 *edges, *par, *gmap, *members;
    MergeAcc2 *acc;
    u64 *scan_state, *msize;
    u32 *scan_ticket, *member_cursor, *block_sums;
    size_t scan_tiles, cap_total;
};
size_t merge_ws_carve(char *p0, int64_t nchunks, int64_t cap_rows, int64_t max_edges, MergeWs *w)
{
    char *p = p0;
    const size_t cap_total = (size_t)nchunks * (size_t)cap_rows;
    w->cap_total = cap_total;
    w->edges = (u32 *)p; p += al((size_t)max_edges * 8);
    w->par = (u32 *)p; p += al((cap_total + 1) * 4);
    w->gmap = (u32 *)p; p += al((cap_total + 1) * 4);
    w->acc = (MergeAcc2 *)p; p += al((cap_total + 1) * sizeof(MergeAcc2));
    w->members = (u32 *)p; p += al(cap_total * 4);
    w->msize = (u64 *)p; p += al((cap_total + 1) * 8);
    // look-back state of the kept-flag scan: [tiles] + ticket (2 u32) + member cursor (u32), then per-CTA sums
    w->scan_tiles = (cap_total + 1) / (SCAN_THREADS * 8) + 2;
    w->scan_state = (u64 *)p;
    w->scan_ticket = (u32 *)(w->scan_state + w->scan_tiles);
    w->member_cursor = w->scan_ticket + 2;
    w->block_sums = (u32 *)(w->scan_state + w->scan_tiles + 2);
    p += al((w->scan_tiles + 2) * 8 + 4096 * 4);
    return (size_t)(p - p0);
}
int merge_args_ok(const void *slots_all, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                  const int32_t *plan_dev, int64_t nchunks, int64_t npairs, int64_t max_edges, void *workspace,
                  int64_t workspace_bytes)
{
    if (!slots_all || !plan_dev || !workspace || npairs < 0) return fail(SYN_ERR_INVALID, "bad arguments");
    const int64_t need = syn_merge_packed_workspace_bytes(nchunks, cap_rows, max_edges);
    if (need < 0) return fail(SYN_ERR_INVALID, "bad nchunks / cap_rows / max_edges");
    if (need > workspace_bytes)
        return fail(SYN_ERR_WORKSPACE, "merge workspace too small: need %lld bytes, got %lld", (long long)need,
                    (long long)workspace_bytes);
    if (slot_bytes != (int64_t)pk_slot_bytes(cap_rows, cap_entries)) return fail(SYN_ERR_INVALID, "slot_bytes mismatch");
    return SYN_OK;
}
}  // namespace

int64_t syn_merge_packed_workspace_bytes(int64_t nchunks, int64_t cap_rows, int64_t max_edges)
{
    if (nchunks < 1 || nchunks > MG_MAXCHUNKS || cap_rows < 1 || max_edges < 1) return SYN_ERR_INVALID;
    if ((size_t)nchunks * (size_t)cap_rows >= (1ull << 31)) return SYN_ERR_INVALID;
    MergeWs w;
    return (int64_t)merge_ws_carve(nullptr, nchunks, cap_rows, max_edges, &w);
}

int syn_merge_ccs_table(const void *slots_all, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                        const int32_t *plan_dev, int64_t nchunks, int64_t size_thr, int64_t max_edges,
                        void *workspace, int64_t workspace_bytes, const int64_t *id_base, int64_t *merged_out,
                        int64_t *info, int owner_rank, int owner_world, void *stream)
{
    if (owner_world < 1 || owner_rank < 0 || owner_rank >= owner_world) return fail(SYN_ERR_INVALID, "bad owner rank / world");
    int rc = merge_args_ok(slots_all, slot_bytes, cap_rows, cap_entries, plan_dev, nchunks, 0, max_edges, workspace,
                           workspace_bytes);
    if (rc != SYN_OK) return rc;
    if (!id_base || !merged_out || !info) return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MergeWs w;
    merge_ws_carve((char *)workspace, nchunks, cap_rows, max_edges, &w);
    const Packed P{(const unsigned char *)slots_all, slot_bytes, cap_rows, cap_entries};
    const int *plan = (const int *)plan_dev;
    const i64 *base = (const i64 *)id_base;
    i64 *inf = (i64 *)info;
    const int wide = grid_for((i64)w.cap_total + 1, 256, 4);
    k_mg_tab_init<<<wide, 256, 0, st>>>(inf, w.acc, w.scan_state, (u32)w.scan_tiles + 2);
    CK_LAUNCH();
    k_mg_tab_acc<<<wide, 256, 0, st>>>(P, plan, base, inf, w.gmap, w.acc, owner_rank, owner_world);
    CK_LAUNCH();
    {
        const i64 ntiles = ((i64)w.cap_total + 1 + SCAN_THREADS * 8 - 1) / (SCAN_THREADS * 8);
        k_scan_1p<MgKeepSrc, MgKeepDst, 8><<<wave_grid(k_scan_1p<MgKeepSrc, MgKeepDst, 8>, SCAN_THREADS, ntiles),
                                             SCAN_THREADS, 0, st>>>(MgKeepSrc{w.acc, inf, size_thr},
                                                                    MgKeepDst{w.acc, w.member_cursor}, w.scan_state,
                                                                    w.scan_ticket, &inf[MG_NMERGED]);
        CK_LAUNCH();
    }
    k_mg_tab_fill<<<wide, 256, 0, st>>>(inf, w.gmap, w.acc, w.members);
    CK_LAUNCH();
    k_mg_emit<<<wide, 256, 0, st>>>(P, plan, base, inf, w.acc, w.members, (i64 *)merged_out);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_merge_ccs_packed(const void *slots_all, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                         const int32_t *plan_dev, int64_t nchunks, int64_t npairs, int64_t size_thr, int64_t max_edges,
                         void *workspace, int64_t workspace_bytes, int64_t *id_base_out, int64_t *merged_out,
                         uint32_t *fmap_out, int64_t *info_out, void *stream)
{
    int rc = merge_args_ok(slots_all, slot_bytes, cap_rows, cap_entries, plan_dev, nchunks, npairs, max_edges,
                           workspace, workspace_bytes);
    if (rc != SYN_OK) return rc;
    if (!id_base_out || !fmap_out || !info_out) return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MergeWs w;
    merge_ws_carve((char *)workspace, nchunks, cap_rows, max_edges, &w);
    Packed P{(const unsigned char *)slots_all, slot_bytes, cap_rows, cap_entries};
    const int *plan = (const int *)plan_dev;
    i64 *base = (i64 *)id_base_out, *info = (i64 *)info_out;
    // the id map: ONE cooperative kernel, its phases separated by grid barriers (few CTAs: a barrier costs more the
    // more CTAs take part)
    static const int per_sm = getenv("SYN_MERGE_CTAS") ? atoi(getenv("SYN_MERGE_CTAS")) : 2;
    const int nblk = (int)std::min<i64>((i64)std::min(per_sm, occ_blocks(k_mg_ids, 256, 0)) * sm_count(),
                                        std::max<i64>(1, ((i64)w.cap_total + 256) / 256));
    i64 me = max_edges, thr = size_thr;
    void *args[] = {(void *)&P, (void *)&plan, (void *)&base, (void *)&info, (void *)&w.par, (void *)&w.gmap,
                    (void *)&w.msize, (void *)&w.edges, (void *)&me, (void *)&thr, (void *)&fmap_out};
    CK(cudaLaunchCooperativeKernel((const void *)k_mg_ids, dim3(nblk), dim3(256), args, 0, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!merged_out) return SYN_OK;            // the caller runs syn_merge_ccs_table itself (another stream)
    return syn_merge_ccs_table(slots_all, slot_bytes, cap_rows, cap_entries, plan_dev, nchunks, size_thr, max_edges,
                               workspace, workspace_bytes, id_base_out, merged_out, info_out, 0, 1, stream);
}

int syn_push_slots(const void *my_slots, int64_t n_slots, int64_t slot_bytes, int64_t cap_rows, int64_t cap_entries,
                   void *const *peer_dst, int n_peers, void *multicast_dst, void *stream)
{
    if (!my_slots || n_slots < 1 || n_peers < 0 || n_peers > 16 || (!multicast_dst && (!peer_dst || n_peers < 1)))
        return fail(SYN_ERR_INVALID, "bad arguments");
    if (slot_bytes != (int64_t)pk_slot_bytes(cap_rows, cap_entries) || (slot_bytes & 15))
        return fail(SYN_ERR_INVALID, "slot_bytes mismatch");
    PeerPtrs P{};
    for (int q = 0; q < n_peers; ++q) P.p[q] = peer_dst ? (unsigned char *)peer_dst[q] : nullptr;
    const unsigned gx = (unsigned)std::max<i64>(8, std::min<i64>(128, 592 / n_slots));
    k_push_slots<<<dim3(gx, (unsigned)n_slots), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char *)my_slots, slot_bytes, cap_rows, cap_entries, P, n_peers, (unsigned char *)multicast_dst);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_merge_overlaps(const uint32_t *rows, const uint64_t *cols, const int64_t *counts, int64_t n, int64_t n_ids,
                       uint64_t *argcol_out, int64_t *maxval_out, void *stream)
{
    if (n < 0 || n_ids < 0 || (n > 0 && (!rows || !cols || !counts)) || (n_ids > 0 && (!argcol_out || !maxval_out)))
        return fail(SYN_ERR_INVALID, "bad arguments");
    if (n_ids == 0) return SYN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_mo_init<<<grid_for(n_ids, 256, 8), 256, 0, st>>>((u64 *)argcol_out, (i64 *)maxval_out, n_ids);
    CK_LAUNCH();
    if (n == 0) return SYN_OK;
    const u32 cap = pow2_at_least(2 * n);
    char *buf = nullptr;
    CK(cudaMallocAsync(&buf, (size_t)cap * sizeof(PairSlot) + sizeof(Counts), st));
    PairTable T{(PairSlot *)buf, cap - 1, (Counts *)(buf + (size_t)cap * sizeof(PairSlot))};
    CK(cudaMemsetAsync(T.cnt, 0, sizeof(Counts), st));
    k_pair_table_init<<<grid_for(cap, 256, 8), 256, 0, st>>>(T);
    CK_LAUNCH();
    k_mo_insert<<<grid_for(n, 256, 8), 256, 0, st>>>(rows, (const u64 *)cols, (const i64 *)counts, n, T);
    CK_LAUNCH();
    k_mo_rowmax<<<grid_for(cap, 256, 8), 256, 0, st>>>(T, (i64 *)maxval_out, n_ids);
    CK_LAUNCH();
    k_mo_rowarg<<<grid_for(cap, 256, 8), 256, 0, st>>>(T, (const i64 *)maxval_out, (u64 *)argcol_out, n_ids);
    CK_LAUNCH();
    CK(cudaFreeAsync(buf, st));
    return SYN_OK;
}

long long syn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int syn_profile_enable(int on)
{
    if (on && !g_prof_created) {
        for (int i = 0; i < PROF_MAX; ++i) CK(cudaEventCreate(&g_prof_ev[i]));
        g_prof_created = true;
    }
    g_prof_on = on != 0;
    g_prof_valid = false;
    g_prof_n = 0;
    return SYN_OK;
}

int syn_profile_read(float *stage_ms, int n)
{
    if (!stage_ms || n < 0) return fail(SYN_ERR_INVALID, "bad arguments");
    if (!g_prof_created || !g_prof_valid || g_prof_n < 2)
        return fail(SYN_ERR_INVALID, "no profiled syn_chunk_ccs call recorded");
    const int k = g_prof_n - 1;
    if (n < k) return fail(SYN_ERR_INVALID, "need room for %d intervals", k);
    CK(cudaEventSynchronize(g_prof_ev[k]));
    for (int i = 0; i < k; ++i) CK(cudaEventElapsedTime(&stage_ms[i], g_prof_ev[i], g_prof_ev[i + 1]));
    return k;
}

const char *syn_profile_name(int i)
{
    if (i < 0 || i + 1 >= g_prof_n) return "";
    return g_prof_name[i + 1];
}


int syn_ids_to_lut(const int64_t *ids, int64_t n, int64_t stride, uint32_t base, uint32_t *lut, int64_t lut_len,
                   void *stream)
{
    if (!lut || n < 0 || lut_len < 0 || (n > 0 && !ids) || stride < 1) return fail(SYN_ERR_INVALID, "bad arguments");
    if (n == 0) return SYN_OK;
    k_ids_to_lut<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const i64 *)ids, n, stride, base, lut, lut_len);
    CK_LAUNCH();
    return SYN_OK;
}

int syn_merge_seg_tables(const int64_t *rows, int64_t n_rows, const uint32_t *gmap, int64_t size_thr,
                         int64_t *merged_out, uint32_t *fmap_out, int64_t *n_merged_dev, void *stream)
{
    if (n_rows < 0 || !fmap_out || !n_merged_dev || (n_rows > 0 && (!rows || !gmap || !merged_out)))
        return fail(SYN_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) {
        CK(cudaMemsetAsync(n_merged_dev, 0, 8, st));
        CK(cudaMemsetAsync(fmap_out, 0, 4, st));
        return SYN_OK;
    }
    MergeAcc *acc = nullptr;
    u32 *members = nullptr;
    CK(cudaMallocAsync(&acc, (size_t)(n_rows + 1) * sizeof(MergeAcc), st));
    CK(cudaMallocAsync(&members, (size_t)n_rows * 4, st));
    k_merge_init<<<grid_for(n_rows + 1, 256, 8), 256, 0, st>>>(acc, n_rows);
    CK_LAUNCH();
    k_merge_accumulate<<<grid_for(n_rows, 256, 8), 256, 0, st>>>((const i64 *)rows, n_rows, gmap, acc);
    CK_LAUNCH();
    k_merge_scan<<<1, SCAN_THREADS, 0, st>>>(acc, n_rows, size_thr, (i64 *)n_merged_dev);
    CK_LAUNCH();
    k_merge_fill<<<grid_for(n_rows + 1, 256, 8), 256, 0, st>>>(n_rows, gmap, acc, members, fmap_out);
    CK_LAUNCH();
    k_merge_emit<<<grid_for(n_rows, 256, 8), 256, 0, st>>>((const i64 *)rows, n_rows, acc, members, (i64 *)merged_out);
    CK_LAUNCH();
    CK(cudaFreeAsync(acc, st));
    CK(cudaFreeAsync(members, st));
    return SYN_OK;
}

}  // extern "C"
