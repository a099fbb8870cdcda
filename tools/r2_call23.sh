#!/bin/bash
for t in 128 256 512 1024 2048 4096; do
env VLB_PLOC_TAIL=$t VLB_BVH_BUILDER=ploc timeout 300 python - <<PY
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
for n in (262144,):
    c = vlb.Context(0); c.set_scene(scenes.atrium(n, seed=7)); c.build_bvh(); b = c.build_bvh(); b = c.build_bvh()
    print("tail $t: build %d tris: %.3f ms, %d nodes" % (n, b.build_ms, b.n_nodes)); c.close()
PY
done
