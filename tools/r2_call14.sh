#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ploc" 2>&1 | tail -3 > gpurun_out/ab12.log
for bps in 2 4; do
for r in 8 16; do
  env VLB_PLOC_BLOCKS_PER_SM=$bps VLB_BVH_BUILDER=ploc VLB_PLOC_RADIUS=$r timeout 300 python - >> gpurun_out/ab12.log 2>&1 <<PY
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
for n in (262144, 3 * (1 << 20)):
    c = vlb.Context(0); c.set_scene(scenes.atrium(n, seed=7)); c.build_bvh(); b = c.build_bvh(); b = c.build_bvh()
    print("blocks/SM $bps radius $r: build %d tris: %.3f ms (sort %.3f), %d nodes" % (n, b.build_ms, b.sort_ms, b.n_nodes)); c.close()
PY
done
done
cat gpurun_out/ab12.log
