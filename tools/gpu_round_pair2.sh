#!/bin/bash
# r01c: sparse pair kernel v2 (warp-per-pixel staging, offset table, uniform-2^k shortcut) — parity, A/B bench, ncu
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_terms or selection or histogram_depths" 2>&1 | tail -15 | tee gpurun_out/pytest_pair2.log
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_pair2.json 2> gpurun_out/bench_s3_pair2.err
BCD_B200_PAIR=v1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_pair1.json 2> gpurun_out/bench_s3_pair1.err
timeout 200 python bench.py --steps 5 --warmup 3 --scales 1 --no-cpu-baseline > gpurun_out/bench_s1_pair2.json 2> gpurun_out/bench_s1_pair2.err
ncu --set full --clock-control none --import-source on -k regex:k_pair_terms -s 0 -c 1 -f -o gpurun_out/prof_pair2 \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_pair2.log 2>&1
BCD_B200_TRACE=1 timeout 100 python tools/trace_e2e.py > gpurun_out/trace_e2e_pair2.txt 2>&1
ls -la gpurun_out | tail -5
