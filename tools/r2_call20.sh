#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gather.py tests/test_gpu_named_configs.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/ab15.log
timeout 600 python tools/c4_bench.py --tag corner-major 2>&1 | tail -1 | cut -c1-420 >> gpurun_out/ab15.log
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_hitmajor.so timeout 600 python tools/c4_bench.py --tag hit-major 2>&1 | tail -1 | cut -c1-420 >> gpurun_out/ab15.log
cat gpurun_out/ab15.log
