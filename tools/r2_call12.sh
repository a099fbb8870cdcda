#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sky_single4.log
timeout 200 python tools/skybox_single_bench.py --tag "lanes4(default)" >> gpurun_out/sky_single4.log 2>&1
L8=$PWD/vulkan-light-bakery_b200/libvlb_bake_lanes8.so
VLB_LIB=$L8 timeout 200 python tools/skybox_single_bench.py --tag "lanes8" >> gpurun_out/sky_single4.log 2>&1
VLB_LIB=$L8 VLB_PROJ_STAGES=2 timeout 200 python tools/skybox_single_bench.py --tag "lanes8 stages2" >> gpurun_out/sky_single4.log 2>&1
VLB_LIB=$L8 VLB_PROJ_LANES=6 timeout 200 python tools/skybox_single_bench.py --tag "lanes6" >> gpurun_out/sky_single4.log 2>&1
VLB_LIB=$L8 VLB_PROJ_LANES=8 VLB_PROJ_PDL=0 timeout 200 python tools/skybox_single_bench.py --tag "lanes8 nopdl" >> gpurun_out/sky_single4.log 2>&1
cat gpurun_out/sky_single4.log
