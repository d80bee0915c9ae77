#!/bin/bash
# N-GPU check of the row-tile path: tile tests on GPU 0, NCCL parity, bench at 1080p and 4096^2
set -x
mkdir -p gpurun_out
nvidia-smi -L
N=${NGPU:-2}
timeout 300 python -m pytest tests/test_gpu_tiles.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_tiles.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | grep -E "dist_parity|DIST_PARITY|Error|error" | tail -20 | tee gpurun_out/dist_parity_${N}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_m1_${N}gpu.json 2> gpurun_out/bench_m1_${N}gpu.err
tail -3 gpurun_out/bench_m1_${N}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --size 4096x4096 > gpurun_out/bench_m1_4k_${N}gpu.json 2> gpurun_out/bench_m1_4k_${N}gpu.err
tail -3 gpurun_out/bench_m1_4k_${N}gpu.err
ls -la gpurun_out | tail -8
