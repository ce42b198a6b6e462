set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time SYNAPTOR_B200_LIB=$PWD/build_variants/libsyn_timing.so timeout 150 python tools/phase_timing.py 1024 1024 128 1.0 0.10 ) > gpurun_out/c3_phase.txt 2>&1
( time timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:k_tile_ccl|k_sector_labels|k_stats_merge' --launch-skip 6 --launch-count 3 -f -o gpurun_out/c3_full python tools/phase_timing.py 1024 1024 128 1.0 0.10 ) > gpurun_out/c3_ncu.log 2>&1
( time timeout 120 python bench.py --config c3 --steps 10 --warmup 3 ) > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -3 gpurun_out/c3_phase.txt; tail -3 gpurun_out/c3_ncu.log; ls -la gpurun_out
