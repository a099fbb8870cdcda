#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sky_single2.log
run() { env "$@" timeout 200 python tools/skybox_single_bench.py --cpu --graph --tag "$*" >> gpurun_out/sky_single2.log 2>&1; }
run VLB_PROJ_PDL=1
run VLB_PROJ_PDL=0
run VLB_PROJ_PDL=1 VLB_PROJ_STAGES=5
run VLB_PROJ_PDL=1 VLB_PROJ_STAGES=2
cat gpurun_out/sky_single2.log
