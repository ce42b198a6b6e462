"""Segment overlap (mirror of synaptor/proc/overlap/__init__.py)."""
from . import overlap
from .overlap import count_overlaps, find_max_overlaps, add_overlapping_seg, convert_to_dict
from . import merge
