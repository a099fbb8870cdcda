#!/bin/bash
python - <<'PY'
import importlib, sys, json
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200")
with vlb.Context(0) as c:
    for i in range(3):
        print(json.dumps(c.cache_peaks()))
PY
timeout 300 python -m pytest tests/test_gpu_diag.py -m gpu -x -q 2>&1 | tail -2
timeout 300 compute-sanitizer --tool memcheck python -c "
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module('vulkan-light-bakery_b200')
with vlb.Context(0) as c: print(c.cache_peaks())
" 2>&1 | tail -3
cd vulkan-light-bakery_b200 >/dev/null; cd ..
VLB_BVH_BUILDER=ploc VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake_l1h.so timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 2 --tag "L1 hints variant" 2>&1 | tail -1
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 2 --tag "product" 2>&1 | tail -1
