"""Value types on the chunk_ccs path (mirror of synaptor/types)."""
from . import bbox
from .bbox import BBox3d, Vec3d, chunk_bboxes, bounds1D
