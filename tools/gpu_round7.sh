#!/bin/bash
# full verification of the current default build + C3 ncu capture + C4 measurement (fp32 and 8-bit nodes)
mkdir -p gpurun_out
L=vulkan-light-bakery_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c3 -f python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 1 > gpurun_out/ncu_bake_c3.log 2>&1
timeout 600 python tools/c4_bench.py --tag fp32-nodes > gpurun_out/c4.jsonl 2> gpurun_out/c4.err
VLB_LIB=$PWD/$L/libvlb_bake_q8.so timeout 600 python tools/c4_bench.py --tag q8-nodes >> gpurun_out/c4.jsonl 2>> gpurun_out/c4.err
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_n1.json; cat gpurun_out/c4.jsonl; tail -n 5 gpurun_out/c4.err
