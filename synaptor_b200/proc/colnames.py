"""Column names of the chunk-task tables (reference: synaptor/proc/colnames.py:6-47)."""
seg_id = "cleft_segid"

centroid_x, centroid_y, centroid_z = "centroid_x", "centroid_y", "centroid_z"
centroid_cols = [centroid_x, centroid_y, centroid_z]

bbox_bx, bbox_by, bbox_bz = "bbox_bx", "bbox_by", "bbox_bz"
bbox_ex, bbox_ey, bbox_ez = "bbox_ex", "bbox_ey", "bbox_ez"
bbox_cols = [bbox_bx, bbox_by, bbox_bz, bbox_ex, bbox_ey, bbox_ez]

size = "size"
size_cols = [size]

seg_info_cols = size_cols + centroid_cols + bbox_cols

src_id = "src_id"
dst_id = "dst_id"
dst_id_hash = "dst_id_hash"

ovl_segid = "max_overlap"
rows = "row_id"
cols = "col_id"
vals = "vals"

contin_filename = "filename"
facehash = "facehash"
graph_id1 = "graph_id1"
graph_id2 = "graph_id2"
