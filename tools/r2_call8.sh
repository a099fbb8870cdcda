#!/bin/bash
mkdir -p gpurun_out
C3="--probes 64x32x64 --dirs 64x64"
: > gpurun_out/ab7.log
for tag in "" nodisc; do
  lib=$PWD/vulkan-light-bakery_b200/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib=${tag:-default(discard)}" >> gpurun_out/ab7.log 2>&1
  VLB_LIB=$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__|lts__" >> gpurun_out/ab7.log
done
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_named_configs.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/ab7.log
cat gpurun_out/ab7.log
