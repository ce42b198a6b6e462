"""
Segmentation utilities on the chunk_ccs / chunk_overlaps path, with the call
signatures of synaptor/seg_utils (reference: synaptor/seg_utils/__init__.py:25-41)
and CUDA implementations behind them.  No CPU fallback.
"""
from . import describe
from .describe import nonzero_unique_ids, centers_of_mass, bounding_boxes, segment_sizes

from . import filtering
from .filtering import filter_segs_by_size, filter_segs_by_id

from . import overlap
from .overlap import count_overlaps

from . import relabel
from .relabel import relabel_data, relabel_data_1N

from . import misc
from .misc import dilate_by_k
