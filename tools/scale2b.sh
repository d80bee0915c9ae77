#!/bin/bash
source <(sed -n '/^run() {/,/^}/p' tools/scale2.sh)
run s2b_4096x2048_t4 2 --size 4096x2048 --steps 5 --warmup 2 --tiles-per-gpu 4
run s2b_4096x2048_t2 2 --size 4096x2048 --steps 5 --warmup 2 --tiles-per-gpu 2
