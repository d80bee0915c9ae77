#!/bin/bash
# scaling measurement at N = $NGPU GPUs (row tiles over NCCL): 1080p and 4096^2, s=3 m=1
set -x
mkdir -p gpurun_out
N=${NGPU:-4}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r1c_${N}gpu.json 2> gpurun_out/bench_r1c_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 3 --warmup 3 --size 4096x4096 > gpurun_out/bench_r1c_4k_${N}gpu.json 2> gpurun_out/bench_r1c_4k_${N}gpu.err
tail -c 600 gpurun_out/bench_r1c_${N}gpu.err
ls -la gpurun_out | tail -4
