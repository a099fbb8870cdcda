#!/bin/bash
# with the instruction cache out of the way the kernel is back on the L1 wavefront wall (86.8 %): re-test the variants that trade L1 wavefronts or occupancy
C3="--probes 64x32x64 --dirs 64x64"
for l in "" _q8 _mb7 _mb6 ""; do
  VLB_BVH_BUILDER=ploc VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib$l" 2>&1 | tail -1
done
for l in "" _q8; do
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
for nm in 4 8 10; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "node_min $nm" 2>&1 | tail -1
done
for rm in 16 24; do
  VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_MIN=$rm timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "refill_min $rm" 2>&1 | tail -1
done
