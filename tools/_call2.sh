mkdir -p gpurun_out
timeout 400 python tools/variant_ab.py build_variants/libsyn_base.so build_variants/libsyn_v1.so build_variants/libsyn_v2.so build_variants/libsyn_v3.so build_variants/libsyn_v5.so build_variants/libsyn_v4.so > gpurun_out/ab.jsonl 2> gpurun_out/ab.err
tail -c 600 gpurun_out/ab.err
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d); continue
    s=d['stage_ms']; print(d['lib'], d['config'], round(d['ms_per_step'],4), round(d['ms_per_step_two_chunks_in_flight'],4), 'tile', s.get('tile_ccl'), 'stats', s.get('stats_merge'), 'parity', d['parity'])
PY
