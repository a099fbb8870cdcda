#!/bin/bash
mkdir -p gpurun_out
timeout 300 bash tools/reference_default_bake.sh > gpurun_out/reference_default_bake.log 2>&1; tail -n 4 gpurun_out/reference_default_bake.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "c3 lbvh" 2>&1 | tail -1
timeout 300 python tools/bake_probe.py --probes 7x7x7 --dirs 3141x1000 --order 3 --tris 262144 --reps 2 --tag "7x7x7 x 3141x1000 on the atrium" 2>&1 | tail -1
