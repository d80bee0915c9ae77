#!/bin/bash
# 4-GPU check of the final code: 1080p default, 4096^2 with the automatic tile count (4 per GPU) and with 1 per GPU
source <(sed -n '/^run() {/,/^}/p' tools/scale2.sh)
run s4_1080_auto 4 --steps 20 --warmup 5
run s4_4096_auto 4 --size 4096x4096 --steps 5 --warmup 2
run s4_4096_t1 4 --size 4096x4096 --steps 5 --warmup 2 --tiles-per-gpu 1
