"""Segmentation steps of the chunk tasks (mirror of synaptor/proc/seg/__init__.py:1-13)."""
from . import misc
from .misc import make_seg_info_dframe

from . import conncomps
from .conncomps import connected_components, dilated_components

from . import continuation
from .continuation import Continuation, Face, ContinFile, extract_all_continuations
from .continuation import hash_chunk_faces

from . import merge
