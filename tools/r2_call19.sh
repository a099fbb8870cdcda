#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gltf.py -m gpu -x -q 2>&1 | tail -2
for b in lbvh ploc; do
env VLB_BVH_BUILDER=$b timeout 300 python - <<PY
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
for n in (262144, 3 * (1 << 20)):
    c = vlb.Context(0); c.set_scene(scenes.atrium(n, seed=7)); c.build_bvh(); b = c.build_bvh(); b = c.build_bvh()
    print("$b: build %d tris: %.3f ms (sort %.3f), %d nodes, bounds %s" % (n, b.build_ms, b.sort_ms, b.n_nodes, list(b.bounds))); c.close()
PY
done
