#!/bin/bash
for v in 1 8 16 20 24 28; do
  VLB_BAKE_VIS_REFILL_MIN=$v timeout 600 python tools/c4_bench.py --reps 1 --tag "vis_refill_min $v" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
VLB_BAKE_VIS_REFILL_MIN=20 timeout 600 python -m pytest tests/test_gpu_gather.py tests/test_gpu_named_configs.py -m gpu -x -q -k "gather or c4" 2>&1 | tail -2
