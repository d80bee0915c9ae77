#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | grep -E "dist_parity|DIST_PARITY|Error|error" | tail -20 | tee gpurun_out/dist_parity_${N}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_m1_${N}gpu.json 2> gpurun_out/bench_m1_${N}gpu.err
tail -3 gpurun_out/bench_m1_${N}gpu.err
ls -la gpurun_out | tail -8
