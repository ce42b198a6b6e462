"""Cross-chunk merge of connected components (mirror of synaptor/proc/seg/merge/__init__.py)."""
from . import assign_ids
from .assign_ids import assign_unique_ids_serial
from .assign_ids import apply_chunk_id_maps, update_chunk_id_maps
from .assign_ids import apply_id_map

from . import merge_ccs
from .merge_ccs import find_connected_continuations, merge_continuations
from .merge_ccs import match_continuations, reconstruct_face
from .merge_ccs import pair_continuation_files

from . import merge_df
from .merge_df import merge_seginfo_df, enforce_size_threshold, add_new_ids

from . import misc
from .misc import expand_id_map
