#!/bin/bash
for tw in 8 16 32 -1; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_TAIL_WAVES=$tw timeout 300 python tools/bake_probe.py --probes 64x32x8 --dirs 64x64 --reps 4 --tag "1/8 share (64x32x8) tail_waves $tw" 2>&1 | tail -1
done
for tw in 8 32 -1; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_TAIL_WAVES=$tw timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "C3 tail_waves $tw" 2>&1 | tail -1
done
