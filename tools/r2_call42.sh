#!/bin/bash
# gather kernel: one shade_prelude per hit (product) vs two (_pre2); odd direction grids test; direct kernel unchanged check
for l in "" _pre2 ""; do
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "c3 product" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
