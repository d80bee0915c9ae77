#!/bin/bash
# last check of the round: full GPU suite + bench A/B (K1b packed accumulation, K3b/K3c warp-level syncs; K3e 8 lanes/matrix)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
BCD_B200_EIG_GROUP=8 timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last_eig8.json 2> gpurun_out/bench_last_eig8.err
