#!/bin/bash
# same-box A/B: product (cold div / sqrt out of line) vs _s1 (inlined, otherwise identical)
C3="--probes 64x32x64 --dirs 64x64"
for l in "" _s1 "" _s1; do
  VLB_BVH_BUILDER=ploc VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib$l" 2>&1 | tail -1
done
for l in "" _s1; do
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 lib$l" 2>&1 | tail -1
done
