#!/bin/bash
# any-hit node steps without the distance sort (VLB_BAKE_ANY_UNORDERED) + refill threshold of phase B
C3="--probes 64x32x64 --dirs 64x64"
export VLB_BVH_BUILDER=ploc
for u in 0 1; do
  VLB_BAKE_ANY_UNORDERED=$u timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "any_unordered $u" 2>&1 | tail -1
  VLB_BAKE_ANY_UNORDERED=$u VLB_BAKE_COUNTERS=2 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters any_unordered $u" 2>&1 | tail -3 | cut -c1-260
done
for u in 0 1; do for b in 8 12 16 24 28; do
  VLB_BAKE_ANY_UNORDERED=$u VLB_BAKE_REFILL_MIN_B=$b timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "any_unordered $u refill_min_b $b" 2>&1 | tail -1
done; done
for l in "" _powin; do
  VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "powf outlined (product) / inlined (_powin): lib$l" 2>&1 | tail -1
done
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "again: product" 2>&1 | tail -1
# chunk length vs refill order: tiles per item 8 / 4 / 2 (atrium 262k)
for pr in 32x16x32 32x16x16 16x16x16; do for o in 1 3; do
  VLB_BAKE_REFILL_ORDER=$o timeout 300 python tools/bake_probe.py --probes $pr --dirs 64x64 --reps 3 --tag "probes $pr order $o" 2>&1 | tail -1
done; done
unset VLB_BVH_BUILDER
for u in 0 1; do
VLB_BAKE_ANY_UNORDERED=$u timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 any_unordered $u" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ray_slot or whole_probe" 2>&1 | tail -2
