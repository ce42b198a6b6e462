mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_final.log 2>&1
tail -4 gpurun_out/gpu_tests_final.log
( time timeout 100 python tools/fuzz_parity.py 250 2026 ) > gpurun_out/fuzz_final.log 2>&1
tail -3 gpurun_out/fuzz_final.log
( time timeout 420 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
tail -c 300 gpurun_out/bench_n1_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['parity']['bit_exact_vs_oracle'], {k:(v['ms_per_step'], v.get('ms_per_step_two_chunks_in_flight')) for k,v in d['config']['other_configs'].items()})"
