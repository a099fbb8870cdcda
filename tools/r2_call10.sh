#!/bin/bash
mkdir -p gpurun_out
C3="--probes 64x32x64 --dirs 64x64"
: > gpurun_out/ab9.log
for tag in "" bvh8 bvh8s6; do
  lib=$PWD/vulkan-light-bakery_b200/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib=${tag:-default}" >> gpurun_out/ab9.log 2>&1
  VLB_BAKE_COUNTERS=2 VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters lib=${tag:-default}" 2>&1 | tail -3 >> gpurun_out/ab9.log
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py --reps 3 --tag "c2 lib=${tag:-default}" >> gpurun_out/ab9.log 2>&1
done
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_bvh8.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py tests/test_gpu_named_configs.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/ab9.log
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_bvh8.so timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "l1tex__|sm__|smsp__" >> gpurun_out/ab9.log
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_bvh8.so timeout 600 python tools/c4_bench.py --tag bvh8 2>&1 | tail -1 | cut -c1-600 >> gpurun_out/ab9.log
cat gpurun_out/ab9.log
