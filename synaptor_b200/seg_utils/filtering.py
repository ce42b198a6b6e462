"""Size / id filters -- signatures of synaptor/seg_utils/filtering.py:7-60."""
from . import describe
from . import relabel


def filter_segs_by_size(seg, thresh, szs=None, to_ignore=None, copy=True):
    """
    Remove segments with size < thresh except those listed in to_ignore.
    Returns (filtered volume, {id: size} of the remaining segments); when nothing
    is removed the SAME array object comes back (filtering.py:36-40).
    """
    if szs is None:
        szs = describe.segment_sizes(seg)
    to_remove = set(segid for (segid, sz) in szs.items() if sz < thresh)
    if to_ignore is not None:
        to_remove = to_remove.difference(to_ignore)
    remaining = {segid: sz for (segid, sz) in szs.items() if segid not in to_remove}
    if len(to_remove) > 0:
        return filter_segs_by_id(seg, to_remove, copy=copy), remaining
    return seg, remaining


def filter_segs_by_id(seg, ids, copy=True):
    """Remove (zero) the segments with the given ids (filtering.py:43-60)."""
    mapping = {int(v): 0 for v in ids}
    if len(mapping) > 0:
        return relabel.relabel_data(seg, mapping, copy=copy)
    return seg
