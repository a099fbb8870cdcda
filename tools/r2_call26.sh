#!/bin/bash
for ord in 0 1; do for rm in 16 20 24; do for nm in 6 8; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=$ord VLB_BAKE_REFILL_MIN=$rm VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 2 --tag "order $ord refill_min $rm node_min $nm" 2>&1 | tail -1
done; done; done
VLB_BAKE_REFILL_ORDER=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bake" 2>&1 | tail -2
