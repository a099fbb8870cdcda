#!/bin/bash
C3="--probes 64x32x64 --dirs 64x64"
for tag in "" c8 c32; do
  lib=$PWD/vulkan-light-bakery_b200/libvlb_bake${tag:+_$tag}.so
  VLB_BVH_BUILDER=ploc VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "new defaults lib=${tag:-default(c16)}" 2>&1 | tail -1
done
VLB_BVH_BUILDER=ploc VLB_BAKE_COUNTERS=2 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters" 2>&1 | tail -3 | cut -c1-330
timeout 300 python tools/bake_probe.py --reps 5 --tag "c2" 2>&1 | tail -1
VLB_BAKE_NODE_MIN=16 VLB_BAKE_REFILL_MIN=1 timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 old policy" 2>&1 | tail -1
timeout 600 python tools/c4_bench.py --tag new-policy 2>&1 | tail -1 | cut -c1-420
VLB_BAKE_NODE_MIN=16 VLB_BAKE_REFILL_MIN=1 timeout 600 python tools/c4_bench.py --tag old-policy 2>&1 | tail -1 | cut -c1-420
timeout 300 python tools/bake_probe.py --probes 7x7x7 --dirs 3141x1000 --order 3 --tris 262144 --reps 2 --tag "7x7x7 x 3141x1000 atrium" 2>&1 | tail -1
VLB_BAKE_NODE_MIN=16 VLB_BAKE_REFILL_MIN=1 timeout 300 python tools/bake_probe.py --probes 7x7x7 --dirs 3141x1000 --order 3 --tris 262144 --reps 2 --tag "7x7x7 x 3141x1000 atrium old policy" 2>&1 | tail -1
timeout 300 bash tools/reference_default_bake.sh 2>&1 | grep baked | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py tests/test_gpu_named_configs.py -m gpu -x -q 2>&1 | tail -2
