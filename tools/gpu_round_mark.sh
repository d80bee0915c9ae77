#!/bin/bash
# marking-kernel round: parity of both K2 kernels, timing of each at 1080p (s=1, s=3) and 4096^2, ncu of the new one
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "marking or search_radii or fractional" 2>&1 | tail -15 | tee gpurun_out/pytest_mark.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for k in sys v4; do
  BCD_B200_MARKING=$k timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_$k.json 2> gpurun_out/bench_s3_$k.err
  BCD_B200_MARKING=$k timeout 200 python bench.py --steps 5 --warmup 3 --scales 1 --no-cpu-baseline > gpurun_out/bench_s1_$k.json 2> gpurun_out/bench_s1_$k.err
  BCD_B200_MARKING=$k timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_4k_$k.json 2> gpurun_out/bench_4k_$k.err
done
ncu --set full --clock-control none --import-source on -k regex:k_marking -s 0 -c 1 -f -o gpurun_out/prof_mark_sys \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_mark_sys.log 2>&1
ls -la gpurun_out
