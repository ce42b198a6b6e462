"""
synaptor_b200 -- B200-native (sm_100a) implementation of Synaptor's chunk_ccs /
chunk_overlaps / merge_ccs / remap_ids hot path behind the reference's own call
signatures (synaptor/__init__.py:1-12).  All array work runs in
libsynaptor_b200.so (hand-written CUDA, ctypes C-ABI); torch is only the buffer
interchange.  There is no CPU fallback.
"""
from . import types
from .types import bbox
from .types.bbox import BBox3d, chunk_bboxes

from . import seg_utils
from .seg_utils import filter_segs_by_size, centers_of_mass, bounding_boxes

from . import proc
from . import io
from .proc.seg import connected_components, dilated_components

__version__ = "0.1.0"
