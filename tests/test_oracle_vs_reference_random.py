"""
The oracle against the EXECUTED reference on random inputs (CPU only).  The golden fixtures pin the oracle on fourteen
fixed cases; here the unmodified reference (baseline/_ref or /root/reference through oracle/refshim.py: MagicMock
for the absent IO packages, scipy.ndimage.label for the absent cc3d, scipy's csgraph for the absent igraph -- the
stand-ins of tests/golden/make_golden.py) is run live next to both oracle flavours on seeded random volumes: random
shapes (degenerate extents included), densities, thresholds, dust sizes, offsets, dilation radii, base segmentations
with ids beyond 2^32, and 2-/3-chunk grids for the merge.  Skipped where no copy of the reference is present.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as orc  # noqa: E402
from oracle import refshim  # noqa: E402
from util import flatten_conts, table_of  # noqa: E402

pytestmark = pytest.mark.skipif(refshim.available() is None, reason="no reference copy (baseline/_ref or /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return refshim.load()[0]


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def random_volume(rng, maxside=26):
    shape = tuple(int(rng.integers(1, maxside)) for _ in range(3))
    if rng.random() < 0.5:
        pred = orc.synth_pred(shape, seed=int(rng.integers(1 << 30)), sigma=float(rng.choice([0.8, 1.5])),
                              fg=float(rng.choice([0.05, 0.2, 0.5])))
    else:
        pred = rng.random(shape, dtype=np.float32) + np.float32(rng.choice([-0.4, -0.2, 0.0, 0.3]))
    return shape, pred


def canon_conts(c):
    meta, xy = flatten_conts(c)
    out, q = [], 0
    for row in meta:
        out.append((tuple(int(v) for v in row), xy[q:q + row[3]].tobytes()))
        q += row[3]
    return sorted(out)


def same_info(a, b):
    ia, ta = table_of(a)
    ib, tb = table_of(b)
    return np.array_equal(ia, ib) and np.array_equal(ta, tb) and list(a.columns) == list(b.columns)


@pytest.mark.parametrize("seed", range(6))
def test_cc_task_and_overlaps(ref, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        shape, pred = random_volume(rng)
        thr = float(rng.choice([0.3, 0.5, 0.7]))
        dust = int(rng.choice([0, 2, 10, 40]))
        off = tuple(int(v) for v in rng.integers(-30, 300, size=3))
        want = quiet(ref.proc.tasks.cc_task, pred.copy(), thr, dust, offset=off)
        for fl in (orc.cc_task_c, orc.cc_task_np):
            got = fl(pred, thr, dust, offset=off)
            assert np.array_equal(got[0], want[0]) and got[0].dtype == want[0].dtype, (fl.__name__, shape, thr, dust)
            assert canon_conts(got[1]) == canon_conts(want[1])
            assert same_info(got[2], want[2])
        base = orc.synth_base(shape, seed=int(rng.integers(1 << 30)), block=int(rng.choice([1, 3, 8])), bg=0.2)
        ov = quiet(ref.proc.tasks.overlap_task, want[0], base)
        for fl in (orc.count_overlaps_c, orc.count_overlaps_np):
            r, c, v, shp = fl(want[0], base)
            assert tuple(ov.shape) == tuple(shp), (fl.__name__, shape)
            assert np.array_equal(np.asarray(ov.row, dtype=np.int64), np.asarray(r, dtype=np.int64))
            assert np.array_equal(np.asarray(ov.col, dtype=np.uint64), np.asarray(c, dtype=np.uint64))
            assert np.array_equal(np.asarray(ov.data, dtype=np.int64), np.asarray(v, dtype=np.int64))
        if ov.nnz:
            mo = ref.proc.overlap.find_max_overlaps(ov)
            mr, mc, mv = orc.find_max_overlaps(np.asarray(ov.row), np.asarray(ov.col), np.asarray(ov.data))
            assert np.array_equal(np.asarray(mo.row, dtype=np.int64), np.asarray(mr, dtype=np.int64))
            assert np.array_equal(np.asarray(mo.col, dtype=np.uint64), np.asarray(mc, dtype=np.uint64))
            assert np.array_equal(np.asarray(mo.data, dtype=np.int64), np.asarray(mv, dtype=np.int64))


@pytest.mark.parametrize("seed", range(3))
def test_dilated_components_and_descriptors(ref, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(5):
        shape, pred = random_volume(rng, maxside=22)
        k = int(rng.integers(1, 6))
        want = ref.proc.seg.dilated_components(pred.copy(), k, 0.5)
        for backend in ("c", "scipy"):
            got = orc.dilated_components(pred, k, 0.5, backend=backend)
            assert np.array_equal(got, want), (backend, shape, k)
        seg = want
        if seg.max() == 0:
            continue
        off = tuple(int(v) for v in rng.integers(0, 50, size=3))
        su = ref.seg_utils
        assert dict(su.segment_sizes(seg)) == dict(orc.segment_sizes_np(seg))
        wc = su.centers_of_mass(seg, offset=off)
        gc = orc.centers_of_mass_np(seg, offset=off)
        assert {int(i): tuple(int(x) for x in v) for i, v in wc.items()} == \
            {int(i): tuple(int(x) for x in v) for i, v in gc.items()}
        wb = su.bounding_boxes(seg, offset=off)
        gb = orc.bounding_boxes_np(seg, offset=off)
        assert {int(i): tuple(int(x) for x in b.min()) + tuple(int(x) for x in b.max()) for i, b in wb.items()} == \
            {int(i): tuple(int(x) for x in b) for i, b in gb.items()}        # (bx, by, bz, ex, ey, ez), exclusive end


@pytest.mark.parametrize("seed", range(8))
def test_merge_ccs_and_remap(ref, seed):
    rng = np.random.default_rng(3000 + seed)
    grid = tuple(int(v) for v in rng.choice([1, 2, 3], size=3))
    cshape = tuple(int(v) for v in rng.integers(6, 14, size=3))
    # ragged last chunks (chunking.py:47-69) on the odd seeds
    vshape = tuple(g * c - (int(rng.integers(0, c // 2)) if (seed % 2 and g > 1) else 0) for g, c in zip(grid, cshape))
    vol = orc.synth_pred(vshape, seed=int(rng.integers(1 << 30)), sigma=2.0, fg=0.15)
    dust, size_thr = int(rng.choice([0, 5])), int(rng.choice([0, 30, 80]))
    boxes = ref.types.bbox.chunk_bboxes(vshape, cshape)
    assert len(boxes) == int(np.prod(grid))
    r_cont, r_info = np.empty(grid, dtype=object), np.empty(grid, dtype=object)
    o_cont, o_info = np.empty(grid, dtype=object), np.empty(grid, dtype=object)
    segs = np.empty(grid, dtype=object)
    for gi, bb in zip(np.ndindex(grid), boxes):
        chunk = np.ascontiguousarray(vol[bb.index()])
        segs[gi], r_cont[gi], r_info[gi] = quiet(ref.proc.tasks.cc_task, chunk.copy(), 0.5, dust, offset=tuple(bb.min()))
        _, o_cont[gi], o_info[gi] = orc.cc_task_c(chunk, 0.5, dust, offset=tuple(bb.min()))
    mfs = (max(vshape),) * 2
    want_df, want_maps = quiet(ref.proc.tasks.merge_ccs_task, r_cont, r_info, size_thr, mfs)
    got_rows, got_maps = orc.merge_ccs_task(o_cont, o_info, size_thr, mfs)
    want_rows = {int(i): [int(v) for v in row] for i, row in zip(want_df.index, want_df[orc.SEG_INFO_COLUMNS].to_numpy())}
    assert {int(k): [int(v) for v in row] for k, row in got_rows.items()} == want_rows
    final_w, final_g = np.zeros(vshape, dtype=np.uint32), np.zeros(vshape, dtype=np.uint32)
    for gi, bb in zip(np.ndindex(grid), boxes):
        assert {int(a): int(b) for a, b in want_maps[gi].items()} == {int(a): int(b) for a, b in got_maps[gi].items()}
        final_w[bb.index()] = quiet(ref.proc.tasks.remap_ids_task, segs[gi].copy(), want_maps[gi], copy=False)
        final_g[bb.index()] = orc.remap_ids(segs[gi].copy(), got_maps[gi])
    assert np.array_equal(final_w, final_g)
    # the merged volume is the whole-volume labelling restricted to the survivors (1:1)
    whole = orc.label_c(vol > 0.5)[0]
    fg = final_g != 0
    pairs = np.unique(np.stack([final_g[fg], whole[fg]], 1), axis=0)
    assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))
