#!/bin/bash
# K3e layout comparison: tests, then the bench with 2 / 4 / 1 matrices per warp, then one ncu capture of the new kernel.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
for g in 16 8 32; do
  BCD_B200_EIG_GROUP=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eig$g.json 2> gpurun_out/bench_eig$g.err
done
ncu --set full --clock-control none --import-source on -k regex:k_eig_chain -s 4 -c 1 -f -o gpurun_out/prof_eigsub \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_eigsub.log 2>&1
ls -la gpurun_out
