#!/bin/bash
# adoption check of VLB_SQ_PACKED + VLB_IDIR_TAB: C3, C2, C4, the whole GPU suite
VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 3 --tag "c3 product" 2>&1 | tail -1
timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 product" 2>&1 | tail -1
timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 product" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
