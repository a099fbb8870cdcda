#!/bin/bash
# per-direction tables (ray direction, projection weights) hoisted out of the kernel: product vs _prev; VLB_BAKE_DIR_TABLES=0 = same code, compute path
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "tables" 2>&1 | tail -1
VLB_BVH_BUILDER=ploc VLB_LIB=$P/libvlb_bake_prev.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "prev" 2>&1 | tail -1
VLB_BVH_BUILDER=ploc VLB_BAKE_DIR_TABLES=0 timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "tables off (same code)" 2>&1 | tail -1
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "tables" 2>&1 | tail -1
for l in "" _prev; do
VLB_LIB=$P/libvlb_bake$l.so timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 lib$l" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
