#!/bin/bash
# r01c: full GPU suite (incl. the CLI end-to-end tests), smoke, benches, launch list, ncu captures of the main kernels
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c_s3.json 2> gpurun_out/bench_r1c_s3.err
timeout 200 python bench.py --steps 5 --warmup 3 --scales 1 --no-cpu-baseline > gpurun_out/bench_r1c_s1.json 2> gpurun_out/bench_r1c_s1.err
timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_r1c_4k.json 2> gpurun_out/bench_r1c_4k.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1c_reference.json 2> gpurun_out/bench_r1c_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
for k in k_pair_terms k_eig_chain k_marking k_filter_apply; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 2 -f -o gpurun_out/prof_r1c_$k \
    timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
BCD_B200_TRACE=1 timeout 100 python tools/trace_e2e.py > gpurun_out/trace_e2e_r1c.txt 2>&1
ls -la gpurun_out | tail -8
