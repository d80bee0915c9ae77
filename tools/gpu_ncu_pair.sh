#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pair_terms -s 0 -c 1 -f -o gpurun_out/prof_pair \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1
ls -la gpurun_out | tail -3
