#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gather.py tests/test_gpu_named_configs.py tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/ab10.log
timeout 600 python tools/c4_bench.py --tag cell-roots 2>&1 | tail -1 | cut -c1-700 >> gpurun_out/ab10.log
VLB_BAKE_CELL_ROOTS=0 timeout 600 python tools/c4_bench.py --tag no-cell-roots 2>&1 | tail -1 | cut -c1-700 >> gpurun_out/ab10.log
cat gpurun_out/ab10.log
