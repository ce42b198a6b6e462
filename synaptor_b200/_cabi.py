"""
ctypes binding of libsynaptor_b200.so (include/synaptor_b200.h).

This is the only way Python reaches the CUDA kernels.  There is NO CPU fallback:
if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYNAPTOR_B200_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("SYNAPTOR_B200_LIB") or os.path.join(_HERE, "libsynaptor_b200.so")

SYN_OK = 0
SYN_ERR_INVALID = -1
SYN_ERR_CUDA = -2
SYN_ERR_CAPACITY = -3
SYN_ERR_WORKSPACE = -4

# counts[] indices (enum in synaptor_b200.h)
CNT_RUNS, CNT_COMPONENTS, CNT_KEPT, CNT_PAIRS, CNT_MAX_BASE, CNT_MAX_LABEL, CNT_ERRFLAGS, CNT_FG_VOXELS, \
    CNT_TABLE_SLOTS, CNT_ACTIVE_WORDS, CNT_TILE_RECORDS, CNT_FLAGGED_TILES, CNT_BND_WORDS, CNT_TICKET = range(14)
NCOUNTS = 32
CNT_DBG0 = 16
SEG_NCOLS = 11

EF_RUNS, EF_SEGS, EF_TABLE, EF_PAIRS = 1, 2, 4, 8

_i64, _i32, _f32, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); exactly the declarations of include/synaptor_b200.h
SIGNATURES = {
    "syn_version": (_i32, []),
    "syn_last_error": (ctypes.c_char_p, []),
    "syn_ccs_workspace_bytes": (_i64, [_i64] * 5),
    "syn_chunk_ccs": (_i32, [_vp, _i64, _i64, _i64, _f32, _i32, _i32, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp,
                             _i64, _vp, _i64, _i64, _vp, _vp]),
    "syn_chunk_ccs_multilabel": (_i32, [_vp, _vp, _i64, _i64, _i64, _f32, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64,
                                        _vp, _vp]),
    "syn_read_counts": (_i32, [_vp, _vp, _vp]),
    "syn_overlap_workspace_bytes": (_i64, [_i64] * 4),
    "syn_count_overlaps": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "syn_extract_faces": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "syn_transpose_zyx": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "syn_relabel_lut": (_i32, [_vp, _i64, _vp, _i64, _vp]),
    "syn_segment_stats": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "syn_max_u32": (_i32, [_vp, _i64, _vp, _vp]),
    "syn_face_pair_edges": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "syn_union_min_map": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "syn_threshold_dilate": (_i32, [_vp, _i64, _i64, _i64, _f32, _i32, _vp, _vp]),
    "syn_unique_u64": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "syn_ids_to_lut": (_i32, [_vp, _i64, _i64, ctypes.c_uint32, _vp, _i64, _vp]),
    "syn_merge_seg_tables": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "syn_merge_ccs_grid": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp]),
    "syn_packed_slot_bytes": (_i64, [_i64, _i64]),
    "syn_chunk_ccs_begin": (_i32, [_vp, _i64, _i64, _i64, _f32, _i32, _i32, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64,
                                   _i32, _vp, _i64, _i64, _vp]),
    "syn_chunk_ccs_finish": (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp,
                                    _i64, _vp, _i64, _i64, _vp, _vp]),
    "syn_push_slots": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _i32, _vp, _vp]),
    "syn_merge_packed_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "syn_merge_ccs_table": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32,
                                   _vp]),
    "syn_merge_ccs_packed": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp,
                                    _vp]),
    "syn_merge_overlaps": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "syn_launch_count": (ctypes.c_longlong, []),
    "syn_profile_enable": (_i32, [_i32]),
    "syn_profile_read": (_i32, [_vp, _i32]),
    "syn_profile_name": (ctypes.c_char_p, [_i32]),
}
PROF_MAX = 40


class SynaptorB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsynaptor_b200 error {code}: {msg}")
        self.code = code


class CapacityError(SynaptorB200Error):
    """A caller-provided capacity was too small; `counts` holds the needed sizes."""

    def __init__(self, code, msg, counts):
        super().__init__(code, msg)
        self.counts = counts


_lib = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C synaptor_b200/csrc).  synaptor_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for (name, (res, args)) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def profile_read():
    """{kernel name: elapsed ms} of the last syn_chunk_ccs call recorded with syn_profile_enable on (ordered)."""
    import numpy as np
    buf = np.zeros(PROF_MAX, dtype=np.float32)
    k = lib().syn_profile_read(buf.ctypes.data, PROF_MAX)
    if k < 0:
        check(k)
    out = {}
    for i in range(k):
        name = lib().syn_profile_name(i).decode()
        out[name] = out.get(name, 0.0) + float(buf[i])
    return out


def last_error():
    return lib().syn_last_error().decode("utf-8", "replace")


def check(rc, counts=None):
    if rc == SYN_OK:
        return
    msg = last_error()
    if rc == SYN_ERR_CAPACITY:
        raise CapacityError(rc, msg, counts)
    raise SynaptorB200Error(rc, msg)
