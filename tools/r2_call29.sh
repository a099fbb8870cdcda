#!/bin/bash
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
for tag in "" sq96 sq128 sq256; do
  lib=$P/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "order 2 lib=${tag:-cap64}" 2>&1 | tail -1
  VLB_LIB=$lib VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__" | tr '\n' ' '; echo
done
