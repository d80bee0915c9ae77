#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiles.py -m gpu -x -q -k "marking or search_radii or fractional or tiles" 2>&1 | tail -15 | tee gpurun_out/pytest_mark.log
for nw in 4 1 8 2; do BCD_B200_MARKING_NW=$nw timeout 120 python tools/mark_probe.py; done 2>&1 | tee gpurun_out/mark_probe.log
timeout 60 python tools/mark_trace.py 1920 290 2>&1 | tee gpurun_out/mark_trace_290.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_sys.json 2> gpurun_out/bench_s3_sys.err
timeout 200 python bench.py --steps 5 --warmup 3 --scales 1 --no-cpu-baseline > gpurun_out/bench_s1_sys.json 2> gpurun_out/bench_s1_sys.err
timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_4k_sys.json 2> gpurun_out/bench_4k_sys.err
