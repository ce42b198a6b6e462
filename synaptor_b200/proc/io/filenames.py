"""File storage conventions of the chunk tasks (reference: synaptor/proc/io/filenames.py:3-42)."""
seginfo_dirname = "seg_infos"
seginfo_fmtstr = "seg_info_{tag}.df"
merged_seginfo_fname = "merged_cleft_info.df"

idmap_dirname = "id_maps"
idmap_fmtstr = "id_map_{tag}.df"

overlaps_dirname = "overlaps"
overlaps_fmtstr = "chunk_overlap_{tag}.df"
max_overlaps_fname = "max_overlaps.df"
