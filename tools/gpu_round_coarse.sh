#!/bin/bash
# streamed selection of the coarser levels: parity + e2e A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_coarse_on.json 2> gpurun_out/bench_coarse_on.err
BCD_B200_STREAM_COARSE=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_coarse_off.json 2> gpurun_out/bench_coarse_off.err
timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_coarse_on_4k.json 2> gpurun_out/bench_coarse_on_4k.err
BCD_B200_TRACE=1 timeout 100 python tools/trace_e2e.py > gpurun_out/trace_e2e_coarse.txt 2>&1
ls -la gpurun_out | tail -5
