#!/bin/bash
# round 2, GPU call 1: parity tests (incl. named configs) + A/B of the shared-memory stack / hit-queue variants on C3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
P=vulkan-light-bakery_b200
C3="--probes 64x32x64 --dirs 64x64 --reps 3"
: > gpurun_out/ab.log
for tag in "" base s8 s16 s24 nohq mb6; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --tag "lib=${tag:-default}" >> gpurun_out/ab.log 2>&1
done
for nm in 4 12 16 24; do VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py $C3 --tag "node_min=$nm" >> gpurun_out/ab.log 2>&1; done
VLB_BVH_CUBIC=1 timeout 300 python tools/bake_probe.py $C3 --tag "cubic-morton" >> gpurun_out/ab.log 2>&1
VLB_BAKE_TAIL_WAVES=-1 timeout 300 python tools/bake_probe.py $C3 --tag "all-partials" >> gpurun_out/ab.log 2>&1
VLB_BAKE_TAIL_WAVES=0 timeout 300 python tools/bake_probe.py $C3 --tag "all-whole" >> gpurun_out/ab.log 2>&1
VLB_BAKE_COUNTERS=2 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters" >> gpurun_out/ab.log 2>&1
VLB_BAKE_COUNTERS=2 VLB_BVH_CUBIC=1 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters-cubic" >> gpurun_out/ab.log 2>&1
VLB_BAKE_COUNTERS=2 VLB_LIB=$PWD/$P/libvlb_bake_base.so timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters-base" >> gpurun_out/ab.log 2>&1
timeout 300 python tools/bake_probe.py --tag "c2-default" >> gpurun_out/ab.log 2>&1
VLB_LIB=$PWD/$P/libvlb_bake_base.so timeout 300 python tools/bake_probe.py --tag "c2-base" >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
