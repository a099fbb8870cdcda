#!/bin/bash
# vlb_baker --devices 0,1 (one process, vlb_bake_probes_multi over two GPUs) against the single-GPU run: same file.
mkdir -p gpurun_out
python - <<'PY'
import importlib, sys
sys.path.insert(0, '.')
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
scenes.write_gltf(scenes.atrium(32768, seed=7), "gpurun_out/atrium32k.gltf", index_dtype=__import__("numpy").uint32)
PY
B=vulkan-light-bakery_b200/vlb_baker
$B gpurun_out/atrium32k.gltf --tight-bounds --probes 32x16x32 --dirs 64x64 --order 2 --light 15,11,9 --out gpurun_out/one.gltf
$B gpurun_out/atrium32k.gltf --tight-bounds --probes 32x16x32 --dirs 64x64 --order 2 --light 15,11,9 --devices 0,1 --out gpurun_out/two.gltf
python - <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200")
a, _ = vlb.deserialize_gltf("gpurun_out/one.gltf"); b, _ = vlb.deserialize_gltf("gpurun_out/two.gltf")
print("two-GPU file bit-identical to the one-GPU file:", bool(np.array_equal(a, b)), a.shape)
PY
rm -f gpurun_out/one.gltf gpurun_out/two.gltf gpurun_out/atrium32k.gltf
