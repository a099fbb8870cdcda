#!/bin/bash
C3="--probes 64x32x64 --dirs 64x64 --reps 2"
P=$PWD/vulkan-light-bakery_b200
VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=1 timeout 300 python tools/bake_probe.py $C3 --tag "order 1 cap 64" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake_sq128.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=1 timeout 300 python tools/bake_probe.py $C3 --tag "order 1 cap 128" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=1 timeout 300 python tools/bake_probe.py $C3 --tag "order 1 cap 544" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 timeout 300 python tools/bake_probe.py $C3 --tag "order 2 (phases) cap 544" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 VLB_BAKE_REFILL_MIN=24 timeout 300 python tools/bake_probe.py $C3 --tag "order 2 cap 544 refill 24" 2>&1 | tail -1
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 VLB_BAKE_REFILL_MIN=28 timeout 300 python tools/bake_probe.py $C3 --tag "order 2 cap 544 refill 28" 2>&1 | tail -1
