#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
timeout 600 python tools/sweep.py --m0 > gpurun_out/sweep_1080p.jsonl 2> gpurun_out/sweep_1080p.err
ls -la gpurun_out | tail -4
