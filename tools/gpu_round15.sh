#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
python - <<'PY'
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
ctx = vlb.Context(0)
for n, seed in ((262144, 7), (3 * (1 << 20), 11)):
    ctx.set_scene(scenes.atrium(n, seed=seed))
    for i in range(3):
        b = ctx.build_bvh(); print("tris", n, "build", i, "ms", round(b.build_ms, 3), "sort", round(b.sort_ms, 3), "nodes", b.n_nodes)
PY
python tools/bake_probe.py --probes 32x16x32 --dirs 64x64 --reps 3 --tag "c3q"
timeout 600 python tools/c4_bench.py --tag fp32-nodes
