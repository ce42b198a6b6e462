"""
Hashing of values to a bucket index for the distributed (sharded-by-hash) merge.
Reference: synaptor/proc/hashing.py:17-100.  A value is packed into an integer from
the little-endian bytes of its decimal text (tuples: the comma-joined texts), then
mapped by the affine hash (a*v + b) mod PRIME, mod maxval.  a and b are the first
two draws of Python's Mersenne twister seeded with 54321 (the reference reseeds the
global `random` on every call; a private generator gives the same two numbers
without that side effect).

Note: the reference's pack_many tests `collections.Iterable`, which Python >= 3.10
no longer has; the fixtures of tests/golden/g9 were generated with that name aliased
to `collections.abc.Iterable`.  `hashtuple` / `hashcolumns2` cannot run in the
reference (`map` is given a generator as its function) and are not mirrored.
"""
import collections.abc
import random

import pandas as pd

HASHED_INDEX_NAME = "hashed_index"
PRIME = 4839472903831
_COEFFS = {}


def _coeffs(seed):
    if seed not in _COEFFS:
        rng = random.Random(seed)
        a = rng.randint(1, PRIME - 1)
        b = rng.randint(0, PRIME - 1)
        _COEFFS[seed] = (a, b)
    return _COEFFS[seed]


def basehash(v, seed=54321, verbose=False):
    a, b = _coeffs(seed)
    ret = (a * v + b) % PRIME
    if verbose:
        print((v, a, b, ret))
    return ret


def pack_many(v, verbose=False):
    if isinstance(v, collections.abc.Iterable):
        text = ",".join(map(str, v))
        if verbose:
            print(text)
        return int.from_bytes(text.encode(), byteorder="little")
    return int.from_bytes(str(v).encode(), byteorder="little")


def hashval(v, maxval):
    """Hash of a value (or tuple of values) to an index below maxval."""
    return basehash(pack_many(v)) % maxval


def hashcolumns(dframe, columns, maxval):
    """Hash of each row's tuple of the given columns."""
    columnvals = list(zip(*(dframe[col] for col in columns)))
    return list(hashval(v, maxval) for v in columnvals)


def add_hashed_index(dframe, columns, maxval, indexname=HASHED_INDEX_NAME, null_fillval=None):
    dframe[indexname] = hashcolumns(dframe, columns, maxval)
    if null_fillval is not None:
        dframe = fill_nulls(dframe, columns, null_fillval, indexname=indexname)
    return dframe


def fill_nulls(dframe, columns, fillval, indexname=HASHED_INDEX_NAME):
    for col in columns:
        dframe.loc[pd.isnull(dframe[col]), indexname] = fillval
    return dframe
