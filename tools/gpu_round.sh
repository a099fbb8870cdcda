#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c2 -f python tools/bake_probe.py --reps 1 > gpurun_out/ncu_bake_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c3 -f python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 1 > gpurun_out/ncu_bake_c3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_project -s 4 -c 2 -o gpurun_out/prof_skybox -f python tools/skybox_bench.py > gpurun_out/ncu_skybox.log 2>&1
timeout 600 python tools/c4_bench.py --tag fp32-nodes > gpurun_out/c4.jsonl 2> gpurun_out/c4.err
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_q8.so timeout 600 python tools/c4_bench.py --tag q8-nodes >> gpurun_out/c4.jsonl 2>> gpurun_out/c4.err
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_n1.json; cat gpurun_out/c4.jsonl | cut -c1-500
