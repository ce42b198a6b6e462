"""
BBox3d / Vec3d / chunk_bboxes -- the value types that appear in the outputs of the
chunk tasks.  Interface follows synaptor/types/bbox/bbox.py:29-166 (half-open
boxes, `astuple()` = (bx, by, bz, ex, ey, ez), `translate`, `index`, `min`/`max`),
synaptor/types/bbox/vector.py:14 and synaptor/types/bbox/chunking.py:13-69.
Written from the documented behaviour, not from the reference source.
"""
import itertools


class Vec3d(tuple):
    """Immutable 3-vector with elementwise arithmetic against vectors or scalars."""
    __slots__ = ()

    def __new__(cls, x_or_triple, y=None, z=None):
        if y is None:
            x_or_triple, y, z = x_or_triple[0], x_or_triple[1], x_or_triple[2]
        return tuple.__new__(cls, (x_or_triple, y, z))

    x = property(lambda s: s[0])
    y = property(lambda s: s[1])
    z = property(lambda s: s[2])

    def _zip(self, o, f):
        if hasattr(o, "__getitem__"):
            return Vec3d(f(self[0], o[0]), f(self[1], o[1]), f(self[2], o[2]))
        return Vec3d(f(self[0], o), f(self[1], o), f(self[2], o))

    def __add__(self, o): return self._zip(o, lambda a, b: a + b)
    __radd__ = __add__
    def __sub__(self, o): return self._zip(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._zip(o, lambda a, b: b - a)
    def __mul__(self, o): return self._zip(o, lambda a, b: a * b)
    __rmul__ = __mul__
    def __floordiv__(self, o): return self._zip(o, lambda a, b: a // b)
    def __truediv__(self, o): return self._zip(o, lambda a, b: a / b)
    def __neg__(self): return Vec3d(-self[0], -self[1], -self[2])
    def __abs__(self): return Vec3d(abs(self[0]), abs(self[1]), abs(self[2]))
    def __eq__(self, o): return hasattr(o, "__getitem__") and len(o) == 3 and tuple(self) == (o[0], o[1], o[2])
    def __ne__(self, o): return not self.__eq__(o)
    def __hash__(self): return tuple.__hash__(self)
    def __repr__(self): return "Vec3d(%s, %s, %s)" % tuple(self)

    def all_le(self, o): return all(a <= b for (a, b) in zip(self, Vec3d(o)))
    def all_gt(self, o): return all(a > b for (a, b) in zip(self, Vec3d(o)))
    def all_ge(self, o): return all(a >= b for (a, b) in zip(self, Vec3d(o)))


def minimum(a, b):
    return Vec3d(min(a[0], b[0]), min(a[1], b[1]), min(a[2], b[2]))


def maximum(a, b):
    return Vec3d(max(a[0], b[0]), max(a[1], b[1]), max(a[2], b[2]))


class BBox3d(object):
    """Half-open 3-D box [min, max).  Built from a box, (begin, end) vectors or 3 slices."""
    __slots__ = ("_min", "_max")

    def __init__(self, v1_or_bbox, v2=None, v3=None):
        if v3 is not None:
            self._set_slices(v1_or_bbox, v2, v3)
        elif v2 is not None:
            self._min, self._max = Vec3d(v1_or_bbox), Vec3d(v2)
        elif isinstance(v1_or_bbox, BBox3d):
            self._min, self._max = v1_or_bbox._min, v1_or_bbox._max
        elif len(v1_or_bbox) == 2:
            self._min, self._max = Vec3d(v1_or_bbox[0]), Vec3d(v1_or_bbox[1])
        elif len(v1_or_bbox) == 3:
            self._set_slices(*v1_or_bbox)
        else:
            raise ValueError("unexpected argument specified")

    def _set_slices(self, x, y, z):
        self._min = Vec3d(x.start, y.start, z.start)
        self._max = Vec3d(x.stop, y.stop, z.stop)

    def index(self):
        return tuple(slice(b, e) for (b, e) in zip(self._min, self._max))

    def min(self): return self._min
    def max(self): return self._max
    def shape(self): return self._max - self._min
    def astuple(self): return tuple(self._min) + tuple(self._max)
    def translate(self, v): return BBox3d(self._min + v, self._max + v)
    def merge(self, other): return BBox3d(minimum(self._min, other._min), maximum(self._max, other._max))
    def intersect(self, other): return BBox3d(maximum(self._min, other._min), minimum(self._max, other._max))
    def transpose(self): return BBox3d(Vec3d(self._min[::-1]), Vec3d(self._max[::-1]))
    def contains(self, x): return self._min.all_le(x) and self._max.all_gt(x)
    def grow_by(self, v): return BBox3d(self._min - Vec3d(v), self._max + Vec3d(v))
    def shrink_by(self, v): return BBox3d(self._min + Vec3d(v), self._max - Vec3d(v))
    def __eq__(self, o): return self._min == o.min() and self._max == o.max()
    def __ne__(self, o): return not self == o
    def __hash__(self): return hash(self.astuple())
    def __repr__(self): return "<BBox3d {}-{}>".format(self._min, self._max)
    def __str__(self): return "BBox3d({}<->{})".format(tuple(self._min), tuple(self._max))


def bounds1D(stop, step_size, start=0):
    """1-D chunk slices; the last one is the remainder (chunking.py:47-69)."""
    assert step_size > 0, f"invalid step_size: {step_size}"
    edges = list(range(start, stop, step_size)) or [start]
    return [slice(b, min(b + step_size, stop)) for b in edges]


def chunk_bboxes(vol_size, chunk_size, offset=None, mip=0):
    """Chunk grid in itertools.product(x, y, z) order (chunking.py:13-44)."""
    if mip > 0:
        f = 2 ** mip
        vol_size = (vol_size[0] // f, vol_size[1] // f, vol_size[2])
        chunk_size = (chunk_size[0] // f, chunk_size[1] // f, chunk_size[2])
        if offset is not None:
            offset = (offset[0] // f, offset[1] // f, offset[2])
    bnds = [bounds1D(vol_size[a], chunk_size[a]) for a in range(3)]
    boxes = [BBox3d(x, y, z) for (x, y, z) in itertools.product(*bnds)]
    if offset is not None:
        boxes = [b.translate(offset) for b in boxes]
    return boxes
