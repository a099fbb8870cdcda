#!/bin/bash
# instruction-cache diet, step 2: cold divisions / square roots out of line, sky texel wrap without integer division (product); + normalize_exact out of line (_nrm); _inl = step 1 inlined forms
C3="--probes 64x32x64 --dirs 64x64"
for l in "" _nrm ""; do
  VLB_BVH_BUILDER=ploc VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "lib$l" 2>&1 | tail -1
done
for l in "" _nrm; do
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 lib$l" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
VLB_LIB=$PWD/vulkan-light-bakery_b200/libvlb_bake$l.so timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 lib$l" 2>&1 | tail -1
done
timeout 300 bash tools/reference_default_bake.sh 2>&1 | grep "baked" | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
VLB_BVH_BUILDER=ploc timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c3_diet -f python tools/bake_probe.py $C3 --reps 1 > gpurun_out/ncu_bake_c3_diet.log 2>&1
