#!/bin/bash
# round 2, GPU call 7 (1 GPU): full GPU suite (gather batches, graph replay), C4 with batched visibility rays, skybox modes, bake C3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python tools/c4_bench.py --tag r2-vis-batches > gpurun_out/c4_r2.jsonl 2> gpurun_out/c4_r2.err; cat gpurun_out/c4_r2.jsonl | cut -c1-700; tail -3 gpurun_out/c4_r2.err
: > gpurun_out/sky_single3.log
env VLB_PROJ_PDL=1 timeout 200 python tools/skybox_single_bench.py --graph --tag "graph-cache" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_GRAPH=0 timeout 200 python tools/skybox_single_bench.py --tag "no-graph-cache" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_LANES=2 timeout 200 python tools/skybox_single_bench.py --tag "lanes2" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_LANES=3 timeout 200 python tools/skybox_single_bench.py --tag "lanes3" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_LANES=1 timeout 200 python tools/skybox_single_bench.py --tag "lanes1" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_STAGES=2 timeout 200 python tools/skybox_single_bench.py --tag "stages2" >> gpurun_out/sky_single3.log 2>&1
env VLB_PROJ_STAGES=4 timeout 200 python tools/skybox_single_bench.py --tag "stages4" >> gpurun_out/sky_single3.log 2>&1
cat gpurun_out/sky_single3.log
timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "c3 default(chunk16)" 2>&1 | tail -2
