#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --m0 > gpurun_out/sweep_1080p.jsonl 2> gpurun_out/sweep_1080p.err
tail -3 gpurun_out/sweep_1080p.err
wc -l gpurun_out/sweep_1080p.jsonl
