#!/bin/bash
# deferred shading (refill order 3: all closest-hit rays of a chunk first, hits queued; then shading + any-hit batches)
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
M=dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
export VLB_BVH_BUILDER=ploc VLB_BAKE_L2_PERSIST=0
timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "product order 1" 2>&1 | tail -1
for tag in d3 d3c32; do
  VLB_LIB=$P/libvlb_bake_$tag.so VLB_BAKE_REFILL_ORDER=3 timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "order 3 lib=$tag" 2>&1 | tail -1
  VLB_LIB=$P/libvlb_bake_$tag.so VLB_BAKE_REFILL_ORDER=3 timeout 300 ncu --metrics $M --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__|l1tex__|smsp__" | tr -s ' ' | tr '\n' ';'; echo
done
VLB_LIB=$P/libvlb_bake_d3.so VLB_BAKE_REFILL_ORDER=1 timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "order 1 lib=d3 (global hit queue)" 2>&1 | tail -1
for rm in 16 24 28; do
  VLB_LIB=$P/libvlb_bake_d3.so VLB_BAKE_REFILL_ORDER=3 VLB_BAKE_REFILL_MIN=$rm timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "order 3 d3 refill_min $rm" 2>&1 | tail -1
done
for nm in 4 8 12; do
  VLB_LIB=$P/libvlb_bake_d3.so VLB_BAKE_REFILL_ORDER=3 VLB_BAKE_NODE_MIN=$nm timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "order 3 d3 node_min $nm" 2>&1 | tail -1
done
VLB_LIB=$P/libvlb_bake_d3.so VLB_BAKE_REFILL_ORDER=3 VLB_BAKE_COUNTERS=2 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters order 3" 2>&1 | tail -3 | cut -c1-300
