"""
h5min -- the small part of HDF5 that a continuation file needs, written and read without libhdf5.

The reference stores the continuations of one chunk face with h5py (synaptor/proc/io/continuation.py:94-118:
`f.create_dataset("face_axis", data=int)`, `f.create_dataset("hi_face", data=bool)`, `f.create_group("face_coords")`
and one `(n, 2)` int64 dataset per segment id inside it) and reads them back with `f[name][()]` / `group.keys()`
(`:53-69`).  Neither h5py nor libhdf5 exists in this environment, so this module emits the file format itself, in
the variant every libhdf5 since 1.6 reads and h5py's default (`libver="earliest"`) produces:

  superblock version 0 . groups as symbol tables (version-1 B-tree of `SNOD` symbol-table nodes + a local heap for
  the link names) . version-1 object headers . dataspace message version 1 . datatype message version 1
  (fixed-point, IEEE float, and the two-member int8 enum {FALSE = 0, TRUE = 1} that h5py maps numpy bool to) .
  fill-value message version 2 (default fill value, late allocation) . data-layout message version 3, contiguous.

Field layouts follow the "HDF5 File Format Specification, version 1.1 / 2.0" (sections III.A.1 superblock, III.B
B-trees, III.C symbol table nodes / entries, III.D local heaps, IV.A.1.a version-1 object header prefix, IV.A.2.b/d/
f/i/r dataspace / datatype / fill value / layout / symbol-table messages).

How it is pinned without libhdf5: the READER in this file is strict (every signature, version, reserved byte,
alignment, size, ordering and address range is checked) and is tested against the one file in this image that
libhdf5 itself wrote (a MATLAB 7.3 file in scipy's test data: superblock 0, symbol-table group, version-1 object
headers -- the same structures).  The WRITER's output must then pass that same reader, and
tests/test_h5min.py compares the structure kinds / versions of both files.  libhdf5 has never opened a file written
by this module; DESIGN.md says so.
"""
import os
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4              # a symbol-table node holds at most 2 * LEAF_K entries
INTERNAL_K = 16         # a group B-tree node holds at most 2 * INTERNAL_K children
SNOD_ENTRIES = 2 * LEAF_K
TREE_CHILDREN = 2 * INTERNAL_K
SNOD_SIZE = 8 + SNOD_ENTRIES * 40
TREE_SIZE = 24 + TREE_CHILDREN * 8 + (TREE_CHILDREN + 1) * 8
HEAP_FREE_NULL = 1      # "no (further) free block" in a local heap

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_STAB, MSG_MTIME = 0xC, 0x10, 0x11, 0x12


class H5FormatError(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# =====================================================================================================================
# writer
# =====================================================================================================================
def _datatype_message(dt):
    """Datatype message, version 1 (spec IV.A.2.d).  Returns the unpadded message body."""
    dt = np.dtype(dt)
    if dt == np.bool_:
        # enum over int8 with members FALSE = 0, TRUE = 1 (h5py's representation of numpy bool): class 8, the number
        # of members in the class bit field, then base type, names (NUL-terminated, padded to 8), values
        base = _datatype_message(np.int8)
        names = b"".join(n + b"\0" * (_pad8(len(n) + 1) - len(n)) for n in (b"FALSE", b"TRUE"))
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.byteorder == ">":
        raise TypeError("big-endian data is not supported")
    if dt.kind in "iu":
        flags0 = 0x08 if dt.kind == "i" else 0x00            # bit 0 byte order (LE), bits 1-2 padding, bit 3 signed
        return struct.pack("<BBBBIHH", 0x10, flags0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        size = dt.itemsize
        exp_loc, exp_size, man_size, bias = (23, 8, 23, 127) if size == 4 else (52, 11, 52, 1023)
        # bit field: byte order LE, mantissa normalisation 2 (msb implied) in bits 4-5, sign location in byte 1
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * size - 1, 0, size, 0, 8 * size, exp_loc, exp_size, 0,
                           man_size, bias)
    raise TypeError(f"dtype {dt} is not supported by h5min")


def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages):
    """Version-1 object header (spec IV.A.1.a): 12-byte prefix + 4 bytes of padding, then 8-aligned messages."""
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


class _Image:
    def __init__(self):
        self.buf = bytearray(96)                   # the superblock goes here at the end

    def put(self, blob):
        addr = _pad8(len(self.buf))
        self.buf.extend(b"\0" * (addr - len(self.buf)))
        self.buf.extend(blob)
        return addr


def _write_dataset(img, value):
    arr = np.asarray(value)
    if arr.dtype.kind not in "biuf":
        raise TypeError(f"dtype {arr.dtype} is not supported by h5min")
    if arr.dtype.itemsize > 1:
        arr = arr.astype(arr.dtype.newbyteorder("<"), copy=False)
    raw = arr.tobytes(order="C")
    data_addr = img.put(raw) if len(raw) else UNDEF
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    msgs = [
        _message(MSG_DATASPACE, space),
        _message(MSG_DATATYPE, _datatype_message(arr.dtype), flags=1),
        _message(MSG_FILL, struct.pack("<BBBBi", 2, 2, 2, 1, 0), flags=1),      # late allocation, fill if set, default value
        _message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),  # version 3, contiguous
    ]
    return img.put(_object_header(msgs))


def _write_group(img, members):
    """members: {name: dict (sub-group) | array-like}.  Returns (object header address, B-tree address, heap address)."""
    names = sorted((n.encode("utf-8") if isinstance(n, str) else bytes(n)) for n in members)
    if len(set(names)) != len(names):
        raise ValueError("duplicate link names")
    by_name = {(n.encode("utf-8") if isinstance(n, str) else bytes(n)): v for n, v in members.items()}
    entries = []                                   # (name, header address, cache type, scratch)
    for name in names:
        if not name or b"/" in name or b"\0" in name:
            raise ValueError(f"bad link name {name!r}")
        v = by_name[name]
        if isinstance(v, dict):
            hdr, bt, hp = _write_group(img, v)
            entries.append((name, hdr, 1, struct.pack("<QQ", bt, hp)))
        else:
            entries.append((name, _write_dataset(img, v), 0, b"\0" * 16))
    # local heap: "" at offset 0 (the key left of everything), then the names, then one free block
    heap_data = bytearray(8)
    offs = []
    for name, *_ in entries:
        offs.append(len(heap_data))
        heap_data.extend(name + b"\0" * (_pad8(len(name) + 1) - len(name)))
    free_off = len(heap_data)
    heap_data.extend(struct.pack("<QQ", HEAP_FREE_NULL, 16))
    heap_addr = _pad8(len(img.buf))
    img.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_addr + 32) + bytes(heap_data))
    # symbol-table nodes, full ones first
    children = []                                  # (address, heap offset of the largest name inside)
    for i in range(0, len(entries), SNOD_ENTRIES):
        part = entries[i:i + SNOD_ENTRIES]
        blob = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
        for j, (name, hdr, ctype, scratch) in enumerate(part):
            blob.extend(struct.pack("<QQII", offs[i + j], hdr, ctype, 0) + scratch)
        blob.extend(b"\0" * (SNOD_SIZE - len(blob)))
        children.append((img.put(bytes(blob)), offs[i + len(part) - 1]))
    # B-tree levels bottom-up; an empty group is one level-0 node without children
    level = 0
    while True:
        groups = [children[i:i + TREE_CHILDREN] for i in range(0, len(children), TREE_CHILDREN)] or [[]]
        addrs = []
        base = _pad8(len(img.buf))
        for gi in range(len(groups)):
            addrs.append(base + gi * TREE_SIZE)    # nodes of a level are laid out back to back (TREE_SIZE % 8 == 0)
        parents = []
        for gi, grp in enumerate(groups):
            left = addrs[gi - 1] if gi > 0 else UNDEF
            right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
            first_key = groups[gi - 1][-1][1] if gi > 0 else 0
            blob = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right))
            blob.extend(struct.pack("<Q", first_key))
            for (caddr, ckey) in grp:
                blob.extend(struct.pack("<QQ", caddr, ckey))
            blob.extend(b"\0" * (TREE_SIZE - len(blob)))
            got = img.put(bytes(blob))
            assert got == addrs[gi]
            parents.append((got, grp[-1][1] if grp else 0))
        if len(parents) == 1:
            btree_addr = parents[0][0]
            break
        children = parents
        level += 1
    hdr = img.put(_object_header([_message(MSG_STAB, struct.pack("<QQ", btree_addr, heap_addr)), _message(MSG_NIL, b"")]))
    return hdr, btree_addr, heap_addr


def write_file(fname, members):
    """Writes the nested dict `members` ({name: array-like | dict}) as an HDF5 file."""
    img = _Image()
    hdr, bt, hp = _write_group(img, members)
    eof = _pad8(len(img.buf))
    img.buf.extend(b"\0" * (eof - len(img.buf)))
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, hp)      # root group symbol-table entry
    assert len(sb) == 96
    img.buf[:96] = sb
    tmp = f"{fname}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(img.buf)
    os.replace(tmp, fname)


# =====================================================================================================================
# strict reader
# =====================================================================================================================
class _Reader:
    def __init__(self, data):
        self.b = data
        self.report = set()                        # (structure, version) pairs met while reading
        for off in [0] + [512 << i for i in range(16)]:
            if data[off:off + 8] == SIG:
                self.sb = off
                break
        else:
            raise H5FormatError("no HDF5 signature")
        self._superblock()

    def need(self, cond, what):
        if not cond:
            raise H5FormatError(what)

    def at(self, addr, n, what):
        """n bytes at file-relative address addr"""
        a = addr + self.base
        self.need(addr != UNDEF and a + n <= self.eof_abs and a + n <= len(self.b), f"{what}: address {addr:#x} + {n} outside the file")
        return self.b[a:a + n]

    def _superblock(self):
        b, o = self.b, self.sb
        self.need(len(b) >= o + 96, "truncated superblock")
        (v_sb, v_fs, v_root, r0, v_shm, so, sl, r1, leaf_k, int_k, _flags) = struct.unpack_from("<BBBBBBBBHHI", b, o + 8)
        self.need(v_sb == 0 and v_fs == 0 and v_root == 0 and v_shm == 0, "only superblock version 0 is supported")
        self.need(r0 == 0 and r1 == 0, "superblock: reserved bytes not zero")
        self.need(so == 8 and sl == 8, "only 8-byte offsets and lengths are supported")
        self.need(leaf_k >= 1 and int_k >= 1, "superblock: bad K values")
        self.leaf_k, self.int_k = leaf_k, int_k
        base, fs_addr, eof, drv = struct.unpack_from("<QQQQ", b, o + 24)
        self.need(fs_addr == UNDEF and drv == UNDEF, "free-space / driver blocks are not supported")
        # libhdf5 1.6 wrote the end-of-file address of a file with a user block as an absolute address
        self.base = base
        self.eof_abs = eof if (base and eof == len(b)) else base + eof
        self.need(self.eof_abs <= len(b), "file shorter than its end-of-file address")
        self.need(base == self.sb, "base address differs from the superblock position")
        name_off, hdr, ctype, rsv = struct.unpack_from("<QQII", b, o + 56)
        self.need(name_off == 0 and rsv == 0, "root entry: bad name offset / reserved field")
        self.root_hdr = hdr
        self.root_scratch = struct.unpack_from("<QQ", b, o + 80) if ctype == 1 else None
        self.need(ctype in (0, 1), "root entry: bad cache type")
        self.report.add(("superblock", 0))

    # ---- object headers -------------------------------------------------------------------------------------------
    def messages(self, addr):
        self.need(addr % 8 == 0, "object header not 8-aligned")
        ver, rsv, nmsg, refcnt, size = struct.unpack_from("<BBHII", self.at(addr, 16, "object header"))
        self.need(ver == 1 and rsv == 0, "only version-1 object headers are supported")
        self.need(refcnt >= 1, "object header: reference count 0")
        self.report.add(("object header", 1))
        out = []
        blocks = [(addr + 16, size)]
        while blocks:
            baddr, bsize = blocks.pop(0)
            blk = self.at(baddr, bsize, "object header block")
            p = 0
            while p < bsize:
                self.need(p + 8 <= bsize, "object header: message header cut off")
                mtype, msize, mflags, r0, r1, r2 = struct.unpack_from("<HHBBBB", blk, p)
                self.need(msize % 8 == 0 and p + 8 + msize <= bsize, f"message {mtype:#x}: bad size {msize}")
                self.need(r0 == 0 and r1 == 0 and r2 == 0, "message header: reserved bytes not zero")
                body = blk[p + 8:p + 8 + msize]
                if mtype == MSG_CONTINUATION:
                    caddr, clen = struct.unpack_from("<QQ", body)
                    blocks.append((caddr, clen))
                out.append((mtype, mflags, body))
                p += 8 + msize
        self.need(len(out) == nmsg, f"object header: {len(out)} messages found, {nmsg} declared")
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------
    def heap(self, addr):
        h = self.at(addr, 32, "local heap")
        self.need(h[:4] == b"HEAP" and h[4] == 0 and h[5:8] == b"\0\0\0", "bad local heap header")
        size, free, daddr = struct.unpack_from("<QQQ", h, 8)
        data = self.at(daddr, size, "local heap data")
        seen = 0
        while free != HEAP_FREE_NULL:              # the free list must stay inside the data segment and end
            self.need(free % 8 == 0 and free + 16 <= size and seen < 1 << 20, "local heap: bad free list")
            nxt, fsz = struct.unpack_from("<QQ", data, free)
            self.need(fsz >= 16 and free + fsz <= size, "local heap: bad free block")
            free, seen = nxt, seen + 1
        self.report.add(("local heap", 0))
        return data

    def heap_name(self, data, off):
        self.need(off < len(data) and off % 8 == 0, "link name offset outside the heap / unaligned")
        end = data.find(b"\0", off)
        self.need(end >= 0, "link name not terminated")
        return bytes(data[off:end])

    def group_entries(self, btree, heap_addr):
        data = self.heap(heap_addr)
        self.need(self.heap_name(data, 0) == b"", "local heap: offset 0 is not the empty name")
        entries = []

        def node(addr, level_expected, lo_key, left_expected):
            """returns (largest key, address of the right sibling)"""
            tsize = 24 + 2 * self.int_k * 8 + (2 * self.int_k + 1) * 8
            t = self.at(addr, tsize, "B-tree node")
            self.need(addr % 8 == 0 and t[:4] == b"TREE", "bad B-tree node")
            ntype, level, used, left, right = struct.unpack_from("<BBHQQ", t, 4)
            self.need(ntype == 0, "B-tree node is not a group node")
            self.need(level_expected is None or level == level_expected, "B-tree: level mismatch")
            self.need(used <= 2 * self.int_k, "B-tree: too many entries")
            self.need(left == left_expected, "B-tree: left sibling mismatch")
            self.report.add(("B-tree node", 1))
            keys = struct.unpack_from(f"<{2 * used + 1}Q", t, 24)     # key0 child0 key1 ... key_used
            self.need(self.heap_name(data, keys[0]) == lo_key, "B-tree: first key differs from the left neighbour's last")
            prev_child = UNDEF
            for i in range(used):
                k_lo = self.heap_name(data, keys[2 * i])
                k_hi = self.heap_name(data, keys[2 * i + 2])
                self.need(k_lo < k_hi, "B-tree: keys not ascending")
                child = keys[2 * i + 1]
                if level > 0:
                    got_hi, _ = node(child, level - 1, k_lo, prev_child)
                    prev_child = child
                else:
                    got_hi = snod(child, k_lo)
                self.need(got_hi == k_hi, "B-tree: key is not the largest name of its child")
            return self.heap_name(data, keys[2 * used]), right

        def snod(addr, lo_key):
            s = self.at(addr, 8 + 2 * self.leaf_k * 40, "symbol table node")
            self.need(addr % 8 == 0 and s[:4] == b"SNOD" and s[4] == 1 and s[5] == 0, "bad symbol table node")
            n = struct.unpack_from("<H", s, 6)[0]
            self.need(1 <= n <= 2 * self.leaf_k, "symbol table node: bad entry count")
            self.report.add(("symbol table node", 1))
            last = lo_key
            for i in range(n):
                noff, hdr, ctype, rsv = struct.unpack_from("<QQII", s, 8 + 40 * i)
                scratch = s[8 + 40 * i + 24:8 + 40 * i + 40]
                name = self.heap_name(data, noff)
                self.need(name > last, "symbol table: names not strictly ascending")
                self.need(rsv == 0 and ctype in (0, 1), "symbol table entry: bad cache type / reserved field")
                entries.append((name, hdr, struct.unpack("<QQ", scratch) if ctype == 1 else None))
                last = name
            return last

        node(btree, None, b"", UNDEF)
        return entries

    def read_object(self, hdr, cached=None):
        msgs = self.messages(hdr)
        by = {}
        for (t, _f, body) in msgs:
            by.setdefault(t, []).append(body)
        if MSG_STAB in by:
            self.need(len(by[MSG_STAB]) == 1 and MSG_LAYOUT not in by, "object is both a group and a dataset")
            bt, hp = struct.unpack_from("<QQ", by[MSG_STAB][0])
            self.need(cached is None or tuple(cached) == (bt, hp), "cached symbol-table addresses differ from the header's")
            self.report.add(("symbol table message", 0))
            out = {}
            for (name, h, cache) in self.group_entries(bt, hp):
                out[name.decode("utf-8")] = self.read_object(h, cache)
            return out
        self.need(cached is None, "a dataset entry carries group scratch data")
        for t in (MSG_DATASPACE, MSG_DATATYPE, MSG_LAYOUT):
            self.need(len(by.get(t, [])) == 1, f"dataset: message {t:#x} missing or repeated")
        shape = self.dataspace(by[MSG_DATASPACE][0])
        dt = self.datatype(by[MSG_DATATYPE][0])[0]
        for body in by.get(MSG_FILL, []):
            self.need(body[0] in (1, 2) and body[1] in (1, 2, 3) and body[2] in (0, 1, 2) and body[3] in (0, 1), "bad fill-value message")
            self.report.add(("fill value message", body[0]))
        for body in by.get(MSG_ATTRIBUTE, []):
            self.attribute(body)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        raw = self.layout(by[MSG_LAYOUT][0], nbytes, shape, dt)
        return np.frombuffer(raw, dtype=dt).reshape(shape).copy()

    # ---- dataset messages -----------------------------------------------------------------------------------------
    def dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        self.need(ver == 1 and flags in (0, 1) and body[3:8] == b"\0" * 5, "only dataspace version 1 is supported")
        self.need(len(body) >= 8 + 8 * rank * (2 if flags else 1), "dataspace message too short")
        self.report.add(("dataspace message", 1))
        return tuple(struct.unpack_from(f"<{rank}Q", body, 8))

    def datatype(self, body):
        """returns (numpy dtype, bytes consumed)"""
        cv, f0, f1, f2, size = struct.unpack_from("<BBBBI", body)
        ver, cls = cv >> 4, cv & 15
        self.need(ver == 1, "only datatype version 1 is supported")
        self.report.add((f"datatype class {cls}", ver))
        if cls == 0:
            off, prec = struct.unpack_from("<HH", body, 8)
            self.need((f0 & 1) == 0 and (f0 & 6) == 0 and f1 == 0 and f2 == 0, "fixed-point: big-endian / padded types unsupported")
            self.need(off == 0 and prec == 8 * size and size in (1, 2, 4, 8), "fixed-point: bad precision")
            return np.dtype(("<i" if f0 & 8 else "<u") + str(size)), 12
        if cls == 1:
            off, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", body, 8)
            ok32 = (size, f1, eloc, esz, mloc, msz, bias) == (4, 31, 23, 8, 0, 23, 127)
            ok64 = (size, f1, eloc, esz, mloc, msz, bias) == (8, 63, 52, 11, 0, 52, 1023)
            self.need(f0 == 0x20 and f2 == 0 and off == 0 and prec == 8 * size and (ok32 or ok64), "float: only IEEE LE 32 / 64")
            return np.dtype(f"<f{size}"), 20
        if cls == 3:
            self.need((f0 & 15) in (0, 1, 2) and (f0 >> 4) in (0, 1), "string: bad padding / charset")
            return np.dtype(f"S{size}"), 8
        if cls == 8:
            nmemb = f0 | (f1 << 8)
            base, used = self.datatype(body[8:])
            self.need(f2 == 0 and base.kind in "iu" and base.itemsize == size, "enum: bad base type")
            p = 8 + used
            names = []
            for _ in range(nmemb):
                end = body.find(b"\0", p)
                self.need(end >= 0, "enum: member name not terminated")
                names.append(bytes(body[p:end]))
                p += _pad8(end + 1 - p)
            values = np.frombuffer(body[p:p + nmemb * size], dtype=base)
            self.need(len(values) == nmemb, "enum: values cut off")
            p += nmemb * size
            if names == [b"FALSE", b"TRUE"] and list(values) == [0, 1] and size == 1:
                return np.dtype(np.bool_), p       # h5py's boolean
            return base, p
        raise H5FormatError(f"datatype class {cls} is not supported")

    def layout(self, body, nbytes, shape, dt):
        ver = body[0]
        if ver == 3:
            self.need(body[1] == 1, "only contiguous layout is supported")
            addr, size = struct.unpack_from("<QQ", body, 2)
        elif ver in (1, 2):
            ndim, cls = body[1], body[2]
            self.need(cls == 1 and body[3:8] == b"\0" * 5, "only contiguous layout is supported")
            addr = struct.unpack_from("<Q", body, 8)[0]
            dims = struct.unpack_from(f"<{ndim}I", body, 16)
            self.need(dims == tuple(shape) + (dt.itemsize,), "layout: dimensions differ from the dataspace")
            size = nbytes
        else:
            raise H5FormatError(f"layout version {ver} is not supported")
        self.report.add(("layout message", ver))
        self.need(size == nbytes, "layout: size differs from dataspace x datatype")
        if nbytes == 0:
            return b""
        self.need(addr % 8 == 0, "raw data not 8-aligned")
        return self.at(addr, nbytes, "raw data")

    def attribute(self, body):
        ver, rsv, nsz, tsz, ssz = struct.unpack_from("<BBHHH", body)
        self.need(ver == 1 and rsv == 0, "only attribute version 1 is supported")
        p = 8
        name = bytes(body[p:p + nsz]).rstrip(b"\0")
        p += _pad8(nsz)
        dt, _ = self.datatype(body[p:p + tsz])
        p += _pad8(tsz)
        shape = self.dataspace(body[p:p + _pad8(ssz)])
        p += _pad8(ssz)
        n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        self.need(p + n <= len(body), "attribute: data cut off")
        self.report.add(("attribute message", 1))
        return name, np.frombuffer(body[p:p + n], dtype=dt).reshape(shape)


def read_file(fname, with_report=False):
    """Reads an HDF5 file of the supported subset into a nested dict {name: numpy array | dict}.  Scalars come back
    as 0-d arrays (`arr[()]` is what `h5py_dataset[()]` returns)."""
    with open(fname, "rb") as f:
        data = f.read()
    r = _Reader(data)
    tree = r.read_object(r.root_hdr, r.root_scratch)
    if not isinstance(tree, dict):
        raise H5FormatError("the root object is not a group")
    return (tree, r.report) if with_report else tree
