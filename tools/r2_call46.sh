#!/bin/bash
# last micro A/Bs: shadow-ray queue records as three float4s (_sq), 1 / direction in the direction table (_idir), both
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
for l in "" _sq _idir _both ""; do
  VLB_BVH_BUILDER=ploc VLB_LIB=$P/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib$l" 2>&1 | tail -1
done
for l in "" _idir; do VLB_LIB=$P/libvlb_bake$l.so timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 lib$l" 2>&1 | tail -1; done
