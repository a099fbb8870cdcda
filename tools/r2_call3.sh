#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "skybox or envmap" > gpurun_out/pytest_sky.log 2>&1; tail -3 gpurun_out/pytest_sky.log
: > gpurun_out/sky_single.log
run() { env "$@" timeout 200 python tools/skybox_single_bench.py --tag "$*" >> gpurun_out/sky_single.log 2>&1; }
run VLB_PROJ_PDL=1
run VLB_PROJ_PDL=0
run VLB_PROJ_PDL=1 VLB_PROJ_STAGES=5
run VLB_PROJ_PDL=1 VLB_PROJ_STAGES=4
run VLB_PROJ_PDL=1 VLB_PROJ_STAGES=2
run VLB_PROJ_PDL=1 VLB_PROJ_CTAS_PER_SM=2 VLB_PROJ_STAGES=3
run VLB_PROJ_PDL=1 VLB_PROJ_CTAS_PER_SM=2 VLB_PROJ_STAGES=5
run VLB_PROJ_PDL=0 VLB_PROJ_CTAS_PER_SM=2 VLB_PROJ_STAGES=5
env VLB_PROJ_PDL=1 timeout 200 python tools/skybox_single_bench.py --order 3 --tag "L3" >> gpurun_out/sky_single.log 2>&1
env VLB_PROJ_PDL=1 timeout 200 python tools/skybox_single_bench.py --wh 4096x2048 --tag "big" >> gpurun_out/sky_single.log 2>&1
env VLB_PROJ_PDL=1 timeout 200 python tools/skybox_single_bench.py --wh 512x256 --tag "small" >> gpurun_out/sky_single.log 2>&1
cat gpurun_out/sky_single.log
