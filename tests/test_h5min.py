"""
h5min (synaptor_b200/proc/io/h5min.py): the HDF5 continuation files of synaptor/proc/io/continuation.py:94-118
without libhdf5.  CPU only.

Pinning: tests/golden/libhdf5_sample_matlab73.mat is a file WRITTEN BY LIBHDF5 (MATLAB 7.4's -v7.3 writer; the file
ships with scipy's test data as scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD licence) that uses the same
structures h5min emits -- superblock 0, symbol-table group (B-tree + SNOD + local heap), version-1 object headers.
The strict reader must parse it to the known content; the writer's files must pass the same reader.
"""
import os
import struct

import numpy as np
import pytest

from util import GOLD, gold

from synaptor_b200.proc.io import h5min

SAMPLE = os.path.join(GOLD, "libhdf5_sample_matlab73.mat")


def test_reader_on_a_file_written_by_libhdf5():
    tree, report = h5min.read_file(SAMPLE, with_report=True)
    assert list(tree) == ["testdouble"]
    assert tree["testdouble"].dtype == np.float64 and tree["testdouble"].shape == (9, 1)
    assert np.array_equal(tree["testdouble"][:, 0], np.linspace(0, 2 * np.pi, 9))      # what scipy's test data holds
    assert {("superblock", 0), ("object header", 1), ("B-tree node", 1), ("symbol table node", 1), ("local heap", 0),
            ("symbol table message", 0), ("dataspace message", 1), ("datatype class 1", 1)} <= report


def test_writer_emits_the_structures_libhdf5_wrote(tmp_path):
    """same structure kinds and versions as the libhdf5 file for everything both contain (the layout message is
    version 3 here -- what libhdf5 >= 1.6.3 and h5py write -- against version 2 in the 2008 sample)"""
    _, ref = h5min.read_file(SAMPLE, with_report=True)
    p = str(tmp_path / "t.h5")
    h5min.write_file(p, {"a": np.arange(6, dtype=np.float64).reshape(3, 2), "g": {"b": np.int64(3)}})
    _, ours = h5min.read_file(p, with_report=True)
    structural = {"superblock", "object header", "B-tree node", "symbol table node", "local heap",
                  "symbol table message", "dataspace message", "datatype class 1"}
    assert {r for r in ref if r[0] in structural} == {r for r in ours if r[0] in structural}
    assert ("layout message", 3) in ours and ("fill value message", 2) in ours and ("datatype class 0", 1) in ours
    # byte-level spot checks against the sample: the float64 datatype message and the scalar dataspace are
    # byte-identical to what libhdf5 wrote there
    raw = open(SAMPLE, "rb").read()
    assert h5min._datatype_message(np.float64) in raw
    assert raw[512:520] == open(p, "rb").read()[:8] == h5min.SIG


@pytest.mark.parametrize("n", [0, 1, 8, 9, 256, 257, 3000])
def test_group_sizes_roundtrip(tmp_path, n):
    """0 entries (empty B-tree node), one full symbol-table node, one full level-0 B-tree node (256 links), two
    B-tree levels"""
    rng = np.random.default_rng(n)
    ids = rng.choice(10 ** 6, size=n, replace=False) + 1
    coords = {str(i): rng.integers(0, 1024, size=(int(rng.integers(1, 9)), 2), dtype=np.int64) for i in ids}
    p = str(tmp_path / "f.h5")
    h5min.write_file(p, {"face_axis": np.asarray(2), "hi_face": np.asarray(True), "face_coords": coords})
    assert os.path.getsize(p) % 8 == 0
    back = h5min.read_file(p)
    assert back["face_axis"].shape == () and back["face_axis"].dtype == np.int64 and back["face_axis"][()] == 2
    assert back["hi_face"].dtype == np.bool_ and back["hi_face"][()] is np.True_
    assert list(back["face_coords"]) == sorted(coords, key=lambda s: s.encode())
    for k, v in coords.items():
        assert back["face_coords"][k].dtype == np.int64 and np.array_equal(back["face_coords"][k], v)


def test_dtypes_and_shapes(tmp_path):
    p = str(tmp_path / "d.h5")
    vals = {"i8": np.int8(-3), "u2": np.arange(5, dtype=np.uint16), "i4": np.arange(24, dtype=np.int32).reshape(2, 3, 4),
            "u8": np.array([2 ** 63 + 5], dtype=np.uint64), "f4": np.array([[1.5, -2.25]], dtype=np.float32),
            "f8": np.float64(np.pi), "b": np.array([True, False, True]), "empty": np.zeros((0, 2), dtype=np.int64),
            "fortran": np.asfortranarray(np.arange(6, dtype=np.int64).reshape(2, 3)),
            "be": np.arange(3, dtype=">i4")}
    h5min.write_file(p, vals)
    back = h5min.read_file(p)
    for k, v in vals.items():
        v = np.asarray(v)
        assert back[k].shape == v.shape and np.array_equal(back[k], v), k
        assert back[k].dtype == v.dtype.newbyteorder("=") or back[k].dtype == v.dtype, k
    with pytest.raises(TypeError):
        h5min.write_file(p, {"s": np.array(["x"])})
    with pytest.raises(ValueError):
        h5min.write_file(p, {"a/b": np.int64(1)})


def test_enum_datatype_bytes():
    """h5py's boolean: datatype class 8 version 1, 2 members, size 1; base int8; names padded to 8; values 0, 1
    (HDF5 spec IV.A.2.d, enumeration properties)"""
    m = h5min._datatype_message(np.bool_)
    assert m == (bytes([0x18, 2, 0, 0]) + struct.pack("<I", 1) +
                 bytes([0x10, 0x08, 0, 0]) + struct.pack("<IHH", 1, 0, 8) +
                 b"FALSE\0\0\0" + b"TRUE\0\0\0\0" + bytes([0, 1]))
    assert h5min._datatype_message(np.int64) == bytes([0x10, 0x08, 0, 0]) + struct.pack("<IHH", 8, 0, 64)


def test_reader_is_strict(tmp_path):
    p = str(tmp_path / "s.h5")
    h5min.write_file(p, {"face_axis": np.asarray(1), "hi_face": np.asarray(False),
                         "face_coords": {str(i): np.full((2, 2), i, dtype=np.int64) for i in range(1, 40)}})
    good = open(p, "rb").read()
    h5min.read_file(p)

    def broken(edit):
        b = bytearray(good)
        edit(b)
        q = str(tmp_path / "bad.h5")
        open(q, "wb").write(b)
        with pytest.raises(h5min.H5FormatError):
            h5min.read_file(q)

    broken(lambda b: b.__setitem__(0, 0))                                   # signature
    broken(lambda b: b.__setitem__(8, 2))                                   # superblock version
    broken(lambda b: b.__delitem__(slice(len(b) - 8, len(b))))              # shorter than the end-of-file address
    snod = good.find(b"SNOD")
    broken(lambda b: b.__setitem__(snod + 4, 9))                            # symbol-table node version
    # swap the names of the first two entries of a node: names no longer ascending
    broken(lambda b: (b.__setitem__(slice(snod + 8, snod + 16), good[snod + 48:snod + 56]),
                      b.__setitem__(slice(snod + 48, snod + 56), good[snod + 8:snod + 16])))
    tree = good.find(b"TREE")
    broken(lambda b: b.__setitem__(tree + 5, 3))                            # B-tree level
    heap = good.find(b"HEAP")
    broken(lambda b: b.__setitem__(slice(heap + 16, heap + 24), struct.pack("<Q", 4)))   # free-list head unaligned


def test_g1_continuations_through_face_files(tmp_path):
    """the continuations of the executed reference's cc_task (fixture g1) written as the six face files and read
    back: same (face, segid, coordinates) content"""
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.proc.seg.continuation import Continuation, Face
    from synaptor_b200.types.bbox import BBox3d
    from util import flatten_conts
    g = gold("g1_cc_task.npz")
    meta, coords = g["cont_meta"], g["cont_coords"]
    conts = {f: [] for f in Face.all_faces()}
    p = 0
    for (axis, hi, segid, n) in meta:
        conts[Face(int(axis), bool(hi))].append(Continuation(int(segid), Face(int(axis), bool(hi)), coords[p:p + n]))
        p += n
    assert len(meta) > 20
    bb = BBox3d((0, 0, 0), tuple(int(s) for s in g["ccs"].shape)) if "ccs" in g.files else BBox3d((0, 0, 0), (64, 64, 64))
    taskio.write_chunk_continuations(conts, str(tmp_path), bb)
    back = taskio.read_chunk_continuations(str(tmp_path), bb)

    def canon(c):
        m, xy = flatten_conts(c)
        out, q = [], 0
        for row in m:
            out.append((tuple(int(v) for v in row), xy[q:q + row[3]].tobytes()))
            q += row[3]
        return sorted(out)

    assert canon(back) == canon(conts)


class _ShimFile:
    """the three h5py calls the reference's continuation IO makes (File as a context manager, `f[path][()]`,
    `group.keys()`, `create_dataset`, `create_group`), backed by h5min"""

    class _Node:
        def __init__(self, v):
            self.v = v

        def __getitem__(self, key):
            if isinstance(key, tuple) and key == ():
                return self.v[()]
            node = self.v
            for part in key.split("/"):
                node = node[part]
            return _ShimFile._Node(node)

        def keys(self):
            return self.v.keys()

        def create_dataset(self, name, data=None):
            self.v[name] = np.asarray(data)

        def create_group(self, name):
            self.v[name] = {}
            return _ShimFile._Node(self.v[name])

    def __init__(self, fname, mode):
        self.fname, self.mode = fname, mode
        self.root = self._Node(h5min.read_file(fname) if mode == "r" else {})

    def __enter__(self):
        return self.root

    def __exit__(self, *exc):
        if self.mode == "w" and exc[0] is None:
            h5min.write_file(self.fname, self.root.v)
        return False


def test_reference_reader_and_writer_code_on_our_files(tmp_path):
    """The reference's OWN `_read_face_file` / `_write_face_file` (proc/io/continuation.py:53-69, 94-118), executed
    unmodified with an `h5py` stand-in backed by h5min: its reader gives our writer's continuations back, and what
    its writer produces through the stand-in our reader reads.  (Exercises the reference's access pattern -- scalar
    `[()]` reads, `face_coords/<segid>` paths, `keys()` order; the bytes are pinned by the tests above.)"""
    import sys
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), ".."))
    from oracle import refshim
    if refshim.available() is None:
        pytest.skip("no reference copy (baseline/_ref or /root/reference)")
    syn, _ = refshim.load()
    import importlib
    rc = importlib.import_module("synaptor.proc.io.continuation")
    shim = types.SimpleNamespace(File=_ShimFile)
    old, rc.h5py = rc.h5py, shim
    try:
        from synaptor_b200.proc import io as taskio
        from synaptor_b200.proc.seg.continuation import Continuation, Face
        rng = np.random.default_rng(5)
        f = Face(1, False)
        conts = [Continuation(int(s), f, rng.integers(0, 99, size=(int(rng.integers(1, 6)), 2), dtype=np.int64))
                 for s in rng.choice(5000, size=300, replace=False) + 1]
        ours = str(tmp_path / "continuations_0_0_0-8_8_8_f1_lo.h5")
        taskio._write_face_file(conts, ours, f)
        got = rc._read_face_file(ours)                       # reference reader, our file
        want = {c.segid: c.face_coords for c in conts}
        assert [c.segid for c in got] == sorted(want, key=lambda s: str(s).encode())
        assert all(np.array_equal(c.face_coords, want[c.segid]) and c.face.axis == 1 and not c.face.hi_index for c in got)
        assert rc.face_from_filename(ours).axis == 1
        theirs = str(tmp_path / "theirs.h5")
        RC = syn.proc.seg.continuation
        rconts = [RC.Continuation(c.segid, RC.Face(1, False), c.face_coords) for c in conts]
        rc._write_face_file(rconts, theirs)                  # reference writer (through the stand-in), our reader
        back = taskio._read_face_file(theirs)
        assert [c.segid for c in back] == [c.segid for c in got]
        assert all(np.array_equal(c.face_coords, want[c.segid]) for c in back)
        assert open(theirs, "rb").read() == open(ours, "rb").read()
    finally:
        rc.h5py = old


def test_random_trees_roundtrip(tmp_path):
    """property test: random nested groups / names / dtypes / shapes survive write -> strict read"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from hypothesis.extra import numpy as hnp

    names = st.text(alphabet=st.characters(min_codepoint=33, max_codepoint=126, blacklist_characters="/"), min_size=1, max_size=20)
    dtypes = st.sampled_from([np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64,
                              np.float32, np.float64, np.bool_])
    arrays = dtypes.flatmap(lambda dt: hnp.arrays(dt, hnp.array_shapes(min_dims=0, max_dims=3, min_side=0, max_side=5),
                                                  elements=hnp.from_dtype(np.dtype(dt), allow_nan=False)))
    trees = st.recursive(st.dictionaries(names, arrays, max_size=12),
                         lambda children: st.dictionaries(names, st.one_of(arrays, children), max_size=6), max_leaves=40)
    counter = [0]

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(trees)
    def run(tree):
        counter[0] += 1
        p = str(tmp_path / f"t{counter[0]}.h5")
        h5min.write_file(p, tree)
        back = h5min.read_file(p)

        def same(a, b):
            if isinstance(a, dict):
                assert isinstance(b, dict) and list(b) == sorted(a, key=lambda s: s.encode("utf-8"))
                for k in a:
                    same(a[k], b[k])
            else:
                a = np.asarray(a)
                assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
        same(tree, back)

    run()


def test_face_file_bytes_do_not_drift(tmp_path):
    """tests/golden/h5min_face_file_sample.h5 is a face file written by this module when it was validated (three
    continuations on the high z face); the writer must keep producing exactly these bytes"""
    from synaptor_b200.proc import io as taskio
    from synaptor_b200.proc.seg.continuation import Continuation, Face
    f = Face(2, True)
    conts = [Continuation(7, f, np.array([[1, 2], [1, 3]], dtype=np.int64)),
             Continuation(12, f, np.array([[5, 5]], dtype=np.int64)),
             Continuation(100, f, np.array([[0, 0], [9, 9], [10, 9]], dtype=np.int64))]
    p = str(tmp_path / "continuations_0_0_0-8_8_8_f2_hi.h5")
    taskio._write_face_file(conts, p, f)
    assert open(p, "rb").read() == open(os.path.join(GOLD, "h5min_face_file_sample.h5"), "rb").read()
    back = taskio._read_face_file(os.path.join(GOLD, "h5min_face_file_sample.h5"))
    assert [(c.segid, c.face_coords.tolist()) for c in back] == \
        [(100, [[0, 0], [9, 9], [10, 9]]), (12, [[5, 5]]), (7, [[1, 2], [1, 3]])]
    assert back[0].face == f
