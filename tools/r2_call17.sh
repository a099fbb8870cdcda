#!/bin/bash
mkdir -p gpurun_out
C3="--probes 64x32x64 --dirs 64x64"
: > gpurun_out/ab13.log
for ml in 2 3 4; do for r in 16 24; do
  env VLB_BVH_MAX_LEAF=$ml VLB_BVH_BUILDER=ploc VLB_PLOC_RADIUS=$r timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "ploc r$r max_leaf $ml" >> gpurun_out/ab13.log 2>&1
done; done
for nm in 12 16 20; do
  env VLB_BAKE_NODE_MIN=$nm VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "ploc r16 node_min $nm" >> gpurun_out/ab13.log 2>&1
done
cat gpurun_out/ab13.log
