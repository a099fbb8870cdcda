#!/bin/bash
C3="--probes 64x32x64 --dirs 64x64"
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "final policy (order 1, refill 20, node_min 6)" 2>&1 | tail -1
VLB_BVH_BUILDER=ploc timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__" | tr '\n' ' '; echo
timeout 300 python tools/bake_probe.py --reps 5 --tag "c2" 2>&1 | tail -1
timeout 600 python tools/c4_bench.py --tag final-policy 2>&1 | tail -1 | cut -c1-420
timeout 300 bash tools/reference_default_bake.sh 2>&1 | grep baked | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
