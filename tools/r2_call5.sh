#!/bin/bash
# round 2, GPU call 5 (1 GPU): comm tests on one GPU, A/B of FFMA2 / branch-free pushes / chunk size on C3, ncu of the new default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_c5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c5.log
tail -n 3 gpurun_out/pytest_gpu_c5.log
P=vulkan-light-bakery_b200
C3="--probes 64x32x64 --dirs 64x64 --reps 3"
: > gpurun_out/ab5.log
for tag in "" base f2 fp c4 c16 c32; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --tag "lib=${tag:-default}" >> gpurun_out/ab5.log 2>&1
done
for tag in "" base c16 c32; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  VLB_LIB=$lib timeout 300 python tools/bake_probe.py --tag "c2 lib=${tag:-default}" >> gpurun_out/ab5.log 2>&1
done
for tag in "" base c32; do
  lib=$PWD/$P/libvlb_bake${tag:+_$tag}.so
  VLB_BAKE_COUNTERS=2 VLB_LIB=$lib timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters lib=${tag:-default}" >> gpurun_out/ab5.log 2>&1
done
cat gpurun_out/ab5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c3_r2 -f python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 1 > gpurun_out/ncu_bake_c3_r2.log 2>&1
ls -la gpurun_out/prof_bake_c3_r2.ncu-rep
