from . import merge_overlaps
from .merge_overlaps import consolidate_overlaps, apply_chunk_id_maps
