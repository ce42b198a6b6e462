/*
 * oracle_c.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Plain-C CPU restatement of the chunk_ccs / chunk_overlaps hot path of
 * nicholasturner1/Synaptor, used as the parity checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.
 *
 * Every function cites the reference file:line whose arithmetic it follows.
 * Arrays are C-contiguous [x][y][z] (z fastest), as they are at the seam
 * after np.ascontiguousarray (synaptor/proc/seg/conncomps.py:21).
 *
 * The 3-D labeler itself lives in a third-party dependency of the reference
 * (PyPI connected-components-3d, import name cc3d, UNPINNED in
 * requirements.txt:15 / setup.py:118, absent from /root/reference).  Its
 * published algorithm is restated in orc_ccl(): one raster pass assigning
 * provisional labels + union-find over the already-visited neighbours,
 * then a renumber pass that walks provisional labels in ascending order and
 * numbers each tree root the first time it is seen.  The consequence -- labels
 * 1..N in raster order of each component's first voxel -- is what the
 * reference's call site (conncomps.py:22) hands downstream.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* threshold: conncomps.py:17  `mask = d > thresh`  (strict, float32;   */
/* NaN compares false).  thresh arrives as the float32 the numpy ufunc  */
/* would compare against (a python float is cast to the array dtype).   */
ORC_API void orc_threshold(const float *d, int64_t n, float thresh, uint8_t *mask)
{
    for (int64_t i = 0; i < n; ++i)
        mask[i] = (uint8_t)(d[i] > thresh);
}

/* ------------------------------------------------------------------ */
/* union-find with growing storage                                     */
typedef struct {
    uint32_t *parent;
    int64_t cap;
    int64_t n; /* labels in use: 1..n */
} uf_t;

static int uf_init(uf_t *u, int64_t cap)
{
    u->parent = (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    u->cap = cap;
    u->n = 0;
    if (!u->parent)
        return -1;
    u->parent[0] = 0;
    return 0;
}

static uint32_t uf_new(uf_t *u)
{
    if (u->n + 1 >= u->cap) {
        int64_t nc = u->cap * 2;
        uint32_t *p = (uint32_t *)realloc(u->parent, (size_t)nc * sizeof(uint32_t));
        if (!p)
            return 0;
        u->parent = p;
        u->cap = nc;
    }
    u->n += 1;
    u->parent[u->n] = (uint32_t)u->n;
    return (uint32_t)u->n;
}

static uint32_t uf_root(uf_t *u, uint32_t x)
{
    uint32_t *p = u->parent;
    while (p[x] != x) {
        p[x] = p[p[x]];
        x = p[x];
    }
    return x;
}

static void uf_unite(uf_t *u, uint32_t a, uint32_t b)
{
    a = uf_root(u, a);
    b = uf_root(u, b);
    if (a == b)
        return;
    if (a < b)
        u->parent[b] = a;
    else
        u->parent[a] = b;
}

/*
 * orc_ccl: connected components of a binary mask (`key == NULL`) or of a
 * multi-label volume (`key != NULL`: voxels join only when their keys are
 * equal and nonzero -- conncomps.py:24-28, the overlap_seg branch).
 * connectivity is 6, 18 or 26 (the reference only ever passes 6:
 * conncomps.py:22,28).  Returns the number of components, or -1 on OOM.
 * labels (uint32, same shape) receives 0 for background and 1..N in raster
 * order of first voxel.
 */
ORC_API int64_t orc_ccl(const uint8_t *mask, const uint64_t *key,
                        int64_t nx, int64_t ny, int64_t nz,
                        int connectivity, uint32_t *labels)
{
    /* the 13 already-visited neighbours in raster order, tagged with the
       number of non-zero offsets (1 = face, 2 = edge, 3 = corner) */
    static const int nb[13][4] = {
        {0, 0, -1, 1},  {0, -1, 0, 1},  {-1, 0, 0, 1},
        {0, -1, -1, 2}, {0, -1, 1, 2},  {-1, 0, -1, 2}, {-1, 0, 1, 2},
        {-1, -1, 0, 2}, {-1, 1, 0, 2},
        {-1, -1, -1, 3}, {-1, -1, 1, 3}, {-1, 1, -1, 3}, {-1, 1, 1, 3}};
    int maxkind = connectivity == 6 ? 1 : (connectivity == 18 ? 2 : 3);
    int64_t n = nx * ny * nz;
    uf_t u;
    if (uf_init(&u, n / 8 + 1024) != 0)
        return -1;

    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                int64_t i = (x * ny + y) * nz + z;
                int fg = key ? (key[i] != 0) : (mask[i] != 0);
                if (!fg) {
                    labels[i] = 0;
                    continue;
                }
                uint32_t lab = 0;
                for (int k = 0; k < 13; ++k) {
                    if (nb[k][3] > maxkind)
                        continue;
                    int64_t xx = x + nb[k][0], yy = y + nb[k][1], zz = z + nb[k][2];
                    if (xx < 0 || yy < 0 || zz < 0 || yy >= ny || zz >= nz)
                        continue;
                    int64_t j = (xx * ny + yy) * nz + zz;
                    uint32_t lj = labels[j];
                    if (lj == 0)
                        continue;
                    if (key && key[j] != key[i])
                        continue;
                    if (lab == 0)
                        lab = lj;
                    else
                        uf_unite(&u, lab, lj);
                }
                if (lab == 0) {
                    lab = uf_new(&u);
                    if (lab == 0) {
                        free(u.parent);
                        return -1;
                    }
                }
                labels[i] = lab;
            }

    /* renumber: ascending provisional label, roots numbered on first sight */
    uint32_t *renum = (uint32_t *)calloc((size_t)u.n + 1, sizeof(uint32_t));
    if (!renum) {
        free(u.parent);
        return -1;
    }
    uint32_t next = 0;
    for (int64_t i = 1; i <= u.n; ++i) {
        uint32_t r = uf_root(&u, (uint32_t)i);
        if (renum[r] == 0)
            renum[r] = ++next;
        renum[i] = renum[r];
    }
    for (int64_t i = 0; i < n; ++i)
        labels[i] = renum[labels[i]];
    free(renum);
    free(u.parent);
    return (int64_t)next;
}

/* ------------------------------------------------------------------ */
/*
 * orc_segment_stats: one pass over a label volume whose ids lie in
 * [0, max_id].  For every id 1..max_id:
 *   size[id]              -- describe.py:74  np.unique(return_counts)
 *   sums[3*id + a]        -- _describe.cpp:30-51  (long accumulators)
 *   bbox[6*id + {0,1,2}]  -- min corner, bbox[6*id + {3,4,5}] -- exclusive max
 *                            (describe.py:49 ndimage.find_objects slices)
 * ids that do not occur keep size 0 and an empty box (min = INT64_MAX, max = 0).
 */
ORC_API void orc_segment_stats(const uint32_t *seg, int64_t nx, int64_t ny, int64_t nz,
                               int64_t max_id, int64_t *size, int64_t *sums, int64_t *bbox)
{
    for (int64_t id = 0; id <= max_id; ++id) {
        size[id] = 0;
        sums[3 * id] = sums[3 * id + 1] = sums[3 * id + 2] = 0;
        bbox[6 * id] = bbox[6 * id + 1] = bbox[6 * id + 2] = INT64_MAX;
        bbox[6 * id + 3] = bbox[6 * id + 4] = bbox[6 * id + 5] = 0;
    }
    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                uint32_t v = seg[(x * ny + y) * nz + z];
                if (v == 0 || (int64_t)v > max_id)
                    continue;
                size[v] += 1;
                sums[3 * v] += x;
                sums[3 * v + 1] += y;
                sums[3 * v + 2] += z;
                int64_t *b = bbox + 6 * (int64_t)v;
                if (x < b[0]) b[0] = x;
                if (y < b[1]) b[1] = y;
                if (z < b[2]) b[2] = z;
                if (x + 1 > b[3]) b[3] = x + 1;
                if (y + 1 > b[4]) b[4] = y + 1;
                if (z + 1 > b[5]) b[5] = z + 1;
            }
}

/*
 * orc_centroid: _describe.cpp:54-61.  `float sz = szs[id]` (long -> float32),
 * `xs[id] / sz` promotes the long sum to float32 and divides in float32,
 * `round()` is half-away-from-zero, the tuple element type is long.
 */
ORC_API int64_t orc_centroid(int64_t sum, int64_t size)
{
    volatile float sz = (float)size;
    volatile float s = (float)sum;
    volatile float q = s / sz;
    return (int64_t)roundf(q);
}

/* ------------------------------------------------------------------ */
/* relabel through a dense LUT: _relabel.cpp:19-38 (`if v in mapping:   */
/* v = mapping[v]`); has[v] says whether v is a key.                    */
ORC_API void orc_relabel_lut(uint32_t *d, int64_t n, const uint32_t *lut,
                             const uint8_t *has, int64_t lut_len)
{
    for (int64_t i = 0; i < n; ++i) {
        uint32_t v = d[i];
        if ((int64_t)v < lut_len && has[v])
            d[i] = lut[v];
    }
}

/* ------------------------------------------------------------------ */
/*
 * orc_count_overlaps: seg_utils/overlap.py:37-60.  Counts (seg1, seg2) pairs
 * over voxels where both are nonzero; entries come out in first-occurrence
 * raster order (the insertion order of collections.Counter, overlap.py:47).
 * Returns nnz (may exceed cap; only the first cap entries are stored), or -1
 * on OOM.  max1 / max2 receive the maxima over the WHOLE volumes
 * (overlap.py:27-28,41-42 use np.unique of each full volume).
 */
typedef struct {
    uint64_t b;
    uint32_t a;
    uint32_t used;
    int64_t idx;
} ovl_slot_t;

static uint64_t ovl_hash(uint32_t a, uint64_t b)
{
    uint64_t h = b * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)a * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

ORC_API int64_t orc_count_overlaps(const uint32_t *seg1, const uint64_t *seg2, int64_t n,
                                   uint32_t *rows, uint64_t *cols, int64_t *counts, int64_t cap,
                                   uint32_t *max1, uint64_t *max2)
{
    int64_t tcap = 1 << 16;
    ovl_slot_t *tab = (ovl_slot_t *)calloc((size_t)tcap, sizeof(ovl_slot_t));
    if (!tab)
        return -1;
    int64_t nnz = 0;
    uint32_t m1 = 0;
    uint64_t m2 = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t a = seg1[i];
        uint64_t b = seg2[i];
        if (a > m1) m1 = a;
        if (b > m2) m2 = b;
        if (a == 0 || b == 0)
            continue;
        if ((nnz + 1) * 2 > tcap) { /* grow + rehash */
            int64_t ncap = tcap * 2;
            ovl_slot_t *nt = (ovl_slot_t *)calloc((size_t)ncap, sizeof(ovl_slot_t));
            if (!nt) {
                free(tab);
                return -1;
            }
            for (int64_t s = 0; s < tcap; ++s)
                if (tab[s].used) {
                    uint64_t h = ovl_hash(tab[s].a, tab[s].b) & (uint64_t)(ncap - 1);
                    while (nt[h].used)
                        h = (h + 1) & (uint64_t)(ncap - 1);
                    nt[h] = tab[s];
                }
            free(tab);
            tab = nt;
            tcap = ncap;
        }
        uint64_t h = ovl_hash(a, b) & (uint64_t)(tcap - 1);
        while (tab[h].used && !(tab[h].a == a && tab[h].b == b))
            h = (h + 1) & (uint64_t)(tcap - 1);
        if (!tab[h].used) {
            tab[h].used = 1;
            tab[h].a = a;
            tab[h].b = b;
            tab[h].idx = nnz;
            if (nnz < cap) {
                rows[nnz] = a;
                cols[nnz] = b;
                counts[nnz] = 0;
            }
            nnz += 1;
        }
        if (tab[h].idx < cap)
            counts[tab[h].idx] += 1;
    }
    free(tab);
    *max1 = m1;
    *max2 = m2;
    return nnz;
}

/* ------------------------------------------------------------------ */
/*
 * orc_dilate_by_k: _seg_utils.cpp:25-82 (manhattan_dist2d) followed by
 * :180-200 (dilate_by_k), instantiated <unsigned int, unsigned int>.  Two
 * raster passes over axes 1 and 2 only; the update rule is
 * `if (neighbour_dist < my_dist) { my_dist = neighbour_dist + 1; my_label =
 * neighbour_label; }`.  seg is modified in place; dists is scratch.
 */
ORC_API void orc_dilate_by_k(uint32_t *seg, uint32_t *dists,
                             int64_t nx, int64_t ny, int64_t nz, uint32_t k)
{
    const uint32_t inf = UINT32_MAX;
#define AT(i, j, kk) (((i)*ny + (j)) * nz + (kk))
    for (int64_t i = 0; i < nx; ++i)
        for (int64_t j = 0; j < ny; ++j)
            for (int64_t kk = 0; kk < nz; ++kk) {
                int64_t c = AT(i, j, kk);
                dists[c] = seg[c] != 0 ? 0 : inf;
                if (j > 0 && dists[AT(i, j - 1, kk)] < dists[c]) {
                    dists[c] = dists[AT(i, j - 1, kk)] + 1;
                    seg[c] = seg[AT(i, j - 1, kk)];
                }
                if (kk > 0 && dists[AT(i, j, kk - 1)] < dists[c]) {
                    dists[c] = dists[AT(i, j, kk - 1)] + 1;
                    seg[c] = seg[AT(i, j, kk - 1)];
                }
            }
    for (int64_t i = nx - 1; i >= 0; --i)
        for (int64_t j = ny - 1; j >= 0; --j)
            for (int64_t kk = nz - 1; kk >= 0; --kk) {
                int64_t c = AT(i, j, kk);
                if (j < ny - 1 && dists[AT(i, j + 1, kk)] < dists[c]) {
                    dists[c] = dists[AT(i, j + 1, kk)] + 1;
                    seg[c] = seg[AT(i, j + 1, kk)];
                }
                if (kk < nz - 1 && dists[AT(i, j, kk + 1)] < dists[c]) {
                    dists[c] = dists[AT(i, j, kk + 1)] + 1;
                    seg[c] = seg[AT(i, j, kk + 1)];
                }
            }
    for (int64_t c = 0; c < nx * ny * nz; ++c)
        if (dists[c] > k)
            seg[c] = 0;
#undef AT
}
