#!/bin/bash
# back on the known-good register allocation: product vs powf out of line; order rule on 8-tile chunks (C4); new policy test
C3="--probes 64x32x64 --dirs 64x64"
for l in "" _powout ""; do
  VLB_BVH_BUILDER=ploc VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib$l" 2>&1 | tail -1
done
timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 default order rule" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_powout.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 powout" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
timeout 300 python tools/bake_probe.py --reps 5 --tag "c2" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ray_slot or whole_probe" 2>&1 | tail -2
