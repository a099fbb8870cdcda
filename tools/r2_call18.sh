#!/bin/bash
mkdir -p gpurun_out
C3="--probes 64x32x64 --dirs 64x64"
: > gpurun_out/ab14.log
for tag in "" l1h; do
  lib=$PWD/vulkan-light-bakery_b200/libvlb_bake${tag:+_$tag}.so
  VLB_BVH_BUILDER=ploc VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib=${tag:-default}" >> gpurun_out/ab14.log 2>&1
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py --reps 3 --tag "c2 lib=${tag:-default}" >> gpurun_out/ab14.log 2>&1
  VLB_BVH_BUILDER=ploc VLB_LIB=$lib timeout 300 ncu --metrics l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "l1tex__|lts__|dram__" >> gpurun_out/ab14.log
done
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_l1h.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py tests/test_textures.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/ab14.log
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_l1h.so timeout 600 python tools/c4_bench.py --tag l1h 2>&1 | tail -1 | cut -c1-420 >> gpurun_out/ab14.log
cat gpurun_out/ab14.log
