#!/bin/bash
# One gpurun call = tests + smoke + bench + ncu evidence (batched: a call costs ~5 min of box time by itself).
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_m1.json 2> gpurun_out/bench_m1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_m1.csv \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_terms -s 0 -c 1 -f -o gpurun_out/prof_pair \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eig_chain -s 0 -c 1 -f -o gpurun_out/prof_eig \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_eig.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_marking -s 0 -c 1 -f -o gpurun_out/prof_mark \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_mark.log 2>&1
ls -la gpurun_out
