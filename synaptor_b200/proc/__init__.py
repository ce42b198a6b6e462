"""Processing tasks of the chunk_ccs / chunk_overlaps / merge_ccs path (mirror of synaptor/proc)."""
from . import colnames
from . import seg
from . import overlap
from . import utils
from . import hashing
from . import io

from . import tasks
from . import tasks_w_io
from .tasks import cc_task
from .tasks import merge_ccs_task
from .tasks import remap_ids_task
from .tasks import overlap_task
from .tasks import merge_overlaps_task
from .tasks import match_continuations_task, seg_graph_cc_task, merge_seginfo_task
