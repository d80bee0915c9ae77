#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_m1_4k_1gpu.json 2> gpurun_out/bench_m1_4k_1gpu.err
ls -la gpurun_out
