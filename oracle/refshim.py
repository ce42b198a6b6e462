"""
Imports the UNMODIFIED reference package for the CPU arm of bench.py (`--impl reference`) -- test / measurement
infrastructure only, never on the product path.

Where it comes from: `baseline/_ref/` (git-ignored; `python -m pip install --no-index --no-build-isolation
--no-deps --target baseline/_ref <copy of /root/reference>`, done by __graft_entry__.build() in the build
container; travels to the GPU box) or, in the build container, /root/reference itself with the three extensions
compiled by oracle/Makefile into oracle/_ref.

Dependencies of the reference that are not installed in this image and are not on the chunk_ccs / chunk_overlaps /
merge_ccs path are replaced by MagicMock; two that ARE on the path get minimal stand-ins, the same ones the golden
fixtures were generated with (tests/golden/make_golden.py):
  cc3d.connected_components -> scipy.ndimage.label (same partition, same raster-first numbering)
  igraph.Graph              -> scipy.sparse.csgraph.connected_components
"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def available():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "synaptor")):
            return c
    return None


def load():
    """Returns (synaptor module, where it was loaded from).  Raises ImportError when no copy is present."""
    where = available()
    if where is None:
        raise ImportError("no reference copy: baseline/_ref/synaptor and /root/reference are both absent")
    if "synaptor" in sys.modules and getattr(sys.modules["synaptor"], "__refshim__", None):
        return sys.modules["synaptor"], sys.modules["synaptor"].__refshim__
    from scipy import ndimage
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components as cg_cc

    for name in ["h5py", "cloudvolume", "taskqueue", "sqlalchemy", "sqlalchemy.sql", "psycopg2", "boto3",
                 "fastremap", "zstandard", "future", "future.standard_library"]:
        sys.modules.setdefault(name, mock.MagicMock())
    try:
        import google.cloud  # namespace package in this image
        storage = mock.MagicMock()
        sys.modules["google.cloud.storage"] = storage
        google.cloud.storage = storage
    except ImportError:
        sys.modules.setdefault("google", mock.MagicMock())
        sys.modules.setdefault("google.cloud", mock.MagicMock())
        sys.modules.setdefault("google.cloud.storage", mock.MagicMock())

    cc3d = types.ModuleType("cc3d")

    def connected_components(arr, connectivity=26, **kw):
        struct = ndimage.generate_binary_structure(3, {6: 1, 18: 2, 26: 3}[connectivity])
        arr = np.asarray(arr)
        if arr.dtype != bool:
            raise NotImplementedError("the stand-in labels boolean masks only (conncomps.py:22)")
        return ndimage.label(arr, structure=struct)[0].astype(np.uint32)

    cc3d.connected_components = connected_components
    sys.modules["cc3d"] = cc3d

    igraph = types.ModuleType("igraph")

    class _VS(dict):
        def __init__(self, n):
            super().__init__()
            self._n = n

        def __len__(self):
            return self._n

    class Graph:
        def __init__(self, n):
            self.n, self.vs, self.edges = n, _VS(n), []

        def add_edges(self, es):
            self.edges.extend(list(es))

        def components(self):
            if self.n == 0:
                return []
            if self.edges:
                e = np.array(self.edges, dtype=np.int64)
                m = coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(self.n, self.n))
            else:
                m = coo_matrix((self.n, self.n))
            _, lab = cg_cc(m, directed=False)
            out = {}
            for (i, l) in enumerate(lab.tolist()):
                out.setdefault(l, []).append(i)
            return list(out.values())

    igraph.Graph = Graph
    sys.modules["igraph"] = igraph

    if where == "/root/reference":      # source tree: its extensions were compiled into oracle/_ref
        refdir = os.path.join(ROOT, "oracle", "_ref")
        for name in ["_describe", "_relabel", "_seg_utils"]:
            f = [x for x in os.listdir(refdir) if x.startswith(name + ".")][0]
            full = "synaptor.seg_utils." + name
            spec = importlib.util.spec_from_file_location(full, os.path.join(refdir, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules[full] = mod
    sys.path.insert(0, where)
    import synaptor
    synaptor.__refshim__ = where
    # pandas probes `zstandard` on its first to_csv / read_csv and cannot take a MagicMock for it; the reference holds
    # its own reference to the stand-in, so it can go from sys.modules again (make_golden.py does the same)
    if isinstance(sys.modules.get("zstandard"), mock.MagicMock):
        sys.modules.pop("zstandard")
    return synaptor, where
