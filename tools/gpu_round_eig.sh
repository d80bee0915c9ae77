#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for g in 16 8; do
  BCD_B200_EIG_GROUP=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eig$g.json 2> gpurun_out/bench_eig$g.err
done
ls -la gpurun_out | tail -4
