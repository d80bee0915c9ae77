#!/bin/bash
# 8-GPU measurements of the row-tile path: 1080p and 4096^2, tiles per GPU swept. Results land in gpurun_out/.
N=${1:-8}
run() { # name, extra args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('value %.1f ms %.3f e2e %.1f tiles %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('tiles'), (d.get('parity') or {}).get('rel_linf_vs_1gpu')))
except Exception as e:
    print('no line', e)
")"
}
run s8_1080_t1 --steps 20 --warmup 5 --tiles-per-gpu 1
run s8_1080_t2 --steps 20 --warmup 5 --tiles-per-gpu 2
run s8_1080_t3 --steps 20 --warmup 5 --tiles-per-gpu 3
BCD_B200_TRACE=1 run s8_1080_t2_trace --steps 1 --warmup 3 --tiles-per-gpu 2
run s8_4096_t1 --size 4096x4096 --steps 5 --warmup 2 --tiles-per-gpu 1
run s8_4096_t3 --size 4096x4096 --steps 5 --warmup 2 --tiles-per-gpu 3
