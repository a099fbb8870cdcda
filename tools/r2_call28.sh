#!/bin/bash
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
for tag in c8p sq512 c32p; do
  VLB_LIB=$P/libvlb_bake_$tag.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "phases lib=$tag" 2>&1 | tail -1
  VLB_LIB=$P/libvlb_bake_$tag.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__|l1tex__|smsp__"
done
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 VLB_BAKE_COUNTERS=2 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters phases" 2>&1 | tail -3 | cut -c1-300
for nm in 4 6 8 12; do
VLB_LIB=$P/libvlb_bake_sq512.so VLB_BVH_BUILDER=ploc VLB_BAKE_REFILL_ORDER=2 VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "phases node_min $nm" 2>&1 | tail -1
done
