mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 200 python tools/variant_ab.py default build_variants/libsyn_sv1.so build_variants/libsyn_sv2.so > gpurun_out/ab2.jsonl 2> gpurun_out/ab2.err
python - <<'PY'
import json
for l in open('gpurun_out/ab2.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d); continue
    s=d['stage_ms']; print(d['lib'], d['config'], round(d['ms_per_step'],4), round(d['ms_per_step_two_chunks_in_flight'],4), 'stats', s.get('stats_merge'), 'parity', d['parity'])
PY
