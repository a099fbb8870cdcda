#!/bin/bash
for rm in 20 24; do for nm in 1 2 4 6 8; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_MIN=$rm VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "refill_min $rm node_min $nm" 2>&1 | tail -1
done; done
