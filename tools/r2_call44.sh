#!/bin/bash
# gather kernel: rolled corner loop + hit-lane table instead of __fns (product) vs the build before (_prev)
for l in "" _prev ""; do
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
done
timeout 600 python -m pytest tests/test_gpu_gather.py tests/test_gpu_named_configs.py tests/test_gpu_diag.py -m gpu -x -q 2>&1 | tail -2
