#!/bin/bash
mkdir -p gpurun_out
P=vulkan-light-bakery_b200
C3="--probes 64x32x64 --dirs 64x64 --reps 3"
: > gpurun_out/ab2.log
for tag in "" nohq base; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  for nm in 8 16; do
    VLB_BAKE_NODE_MIN=$nm VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --tag "lib=${tag:-default} node_min=$nm" >> gpurun_out/ab2.log 2>&1
  done
done
for tag in "" base; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py --tag "c2 lib=${tag:-default}" >> gpurun_out/ab2.log 2>&1
done
cat gpurun_out/ab2.log
