#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ploc" 2>&1 | tail -5 > gpurun_out/ab11.log
C3="--probes 64x32x64 --dirs 64x64"
for r in 0 8 16 32; do
  if [ $r = 0 ]; then E="VLB_BVH_BUILDER=lbvh"; else E="VLB_BVH_BUILDER=ploc VLB_PLOC_RADIUS=$r"; fi
  env $E VLB_BAKE_COUNTERS=1 timeout 300 python tools/bake_probe.py $C3 --reps 1 --tag "counters $E" 2>&1 | tail -2 >> gpurun_out/ab11.log
  env $E timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "$E" >> gpurun_out/ab11.log 2>&1
  env $E timeout 300 python tools/bake_probe.py --reps 3 --tag "c2 $E" >> gpurun_out/ab11.log 2>&1
  env $E timeout 300 python - >> gpurun_out/ab11.log 2>&1 <<PY
import importlib, sys
sys.path.insert(0, '.')
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
for n in (262144, 3 * (1 << 20)):
    c = vlb.Context(0); c.set_scene(scenes.atrium(n, seed=7)); c.build_bvh(); b = c.build_bvh(); b = c.build_bvh()
    print("   build %d tris: %.3f ms (sort %.3f), %d nodes" % (n, b.build_ms, b.sort_ms, b.n_nodes)); c.close()
PY
done
env VLB_BVH_BUILDER=ploc VLB_PLOC_RADIUS=16 timeout 600 python tools/c4_bench.py --tag ploc16 2>&1 | tail -1 | cut -c1-500 >> gpurun_out/ab11.log
cat gpurun_out/ab11.log
