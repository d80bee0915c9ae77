#!/bin/bash
set -x
mkdir -p gpurun_out
N=${NGPU:-4}
nvidia-smi -L | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | grep -E "dist_parity|DIST_PARITY|Error|error" | tail -20 | tee gpurun_out/dist_parity_${N}.log
for n in 2 $N; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_m1_${n}gpu.json 2> gpurun_out/bench_m1_${n}gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --size 4096x4096 > gpurun_out/bench_m1_4k_${n}gpu.json 2> gpurun_out/bench_m1_4k_${n}gpu.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --size 4096x4096 --no-cpu-baseline > gpurun_out/bench_m1_4k_1gpu.json 2> gpurun_out/bench_m1_4k_1gpu.err
ls -la gpurun_out | tail -12
