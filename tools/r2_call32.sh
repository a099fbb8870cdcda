#!/bin/bash
for nm in 4 6 8 12; do for lm in 1 4 8 12 16; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_NODE_MIN=$nm VLB_BAKE_LEAF_MIN=$lm timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 2 --tag "node_min $nm leaf_min $lm" 2>&1 | tail -1 | cut -c1-120
done; done
