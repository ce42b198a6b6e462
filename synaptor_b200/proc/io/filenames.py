"""File storage conventions of the chunk tasks (reference: synaptor/proc/io/filenames.py:3-42)."""
seginfo_dirname = "seg_infos"
seginfo_fmtstr = "seg_info_{tag}.df"
merged_seginfo_fname = "merged_cleft_info.df"
merged_seginfo_fmtstr = "merged_cleft_info_{tag}.df"

contin_dirname = "continuations"
contin_fmtstr = "continuations_{chunk_tag}_{face_tag}.h5"

idmap_dirname = "id_maps"
idmap_fmtstr = "id_map_{tag}.df"

uniquemap_dirname = "unique_ids"
uniquemap_fmtstr = "unique_ids_{tag}.df"

dup_map_fname = "dup_id_map.df"
tagged_dup_fname = "dup_id_map_{i}.df"

overlaps_dirname = "overlaps"
overlaps_fmtstr = "chunk_overlap_{tag}.df"
max_overlaps_fname = "max_overlaps.df"

timing_dirname = "task_durations"
timing_fmtstr = "{tag}"
