#!/bin/bash
# round-2 profiling pass (1 GPU): launch list of two frames, full captures of the pair-term and eigen kernels, the C5 sweep
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r02_launches.csv python tools/prof_frame.py 3 > gpurun_out/r02_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_terms_sparse -s 3 -c 1 -o gpurun_out/r02_prof_pair python tools/prof_frame.py 2 > gpurun_out/r02_prof_pair.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eig_chain_sub -s 6 -c 2 -o gpurun_out/r02_prof_eig python tools/prof_frame.py 2 > gpurun_out/r02_prof_eig.out 2>&1
python tools/sweep.py --m0 --steps 2 > gpurun_out/r02_sweep_1080p.jsonl 2> gpurun_out/r02_sweep.err
python bench.py --steps 10 --warmup 3 --m 0 --no-plugin --alt-spp 0 --cpu-crop 384 > gpurun_out/r02_bench_m0.json 2> gpurun_out/r02_bench_m0.err
ls -la gpurun_out | tail -12
