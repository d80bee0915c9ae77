#!/bin/bash
# 2-GPU checks of the cyclic multi-tile path at 4096-wide frames + 1080p tile-count A/B; N=1 4096^2 baseline
run() { # name, nproc, extra args...
  name=$1; n=$2; shift; shift
  if [ $n -gt 1 ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('value %.1f ms %.3f e2e_ms %.3f tiles %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('tiles'), (d.get('parity') or {}).get('rel_linf_vs_1gpu', (d.get('parity') or {}).get('rel_linf'))))
except Exception as e:
    print('no line', e)
")"
}
run s2_4096x2048_t4 2 --size 4096x2048 --steps 5 --warmup 2 --tiles-per-gpu 4
run s2_4096x2048_t1 2 --size 4096x2048 --steps 5 --warmup 2 --tiles-per-gpu 1
run s2_1080_t1 2 --steps 10 --warmup 3 --tiles-per-gpu 1
run s2_1080_t2 2 --steps 10 --warmup 3 --tiles-per-gpu 2
run s2_1080_auto 2 --steps 10 --warmup 3
run s1_4096 1 --size 4096x4096 --steps 5 --warmup 2 --no-plugin --alt-spp 0 --no-cpu-baseline
