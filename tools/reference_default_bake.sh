#!/bin/bash
# The reference's own workload: `baker default_blender_cube.gltf` = 7x7x7 probes x 3141x1000 rays, 16 coefficients
# (light_baker.cpp:38,65,294), here through vlb_baker on the same cube written by scenes.default_cube().
mkdir -p gpurun_out
python - <<'PY'
import importlib, sys
sys.path.insert(0, '.')
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
scenes.write_gltf(scenes.default_cube(), "gpurun_out/default_cube.gltf")
PY
for i in 1 2; do vulkan-light-bakery_b200/vlb_baker gpurun_out/default_cube.gltf; done 2>&1 | tee gpurun_out/reference_default_bake.log
