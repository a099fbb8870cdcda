#!/bin/bash
# L2 access-policy window A/B on C3 (time + DRAM traffic) and C4
C3="--probes 64x32x64 --dirs 64x64"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
run() {  # tag, env...
  tag=$1; shift
  env "$@" VLB_BVH_BUILDER=ploc timeout 300 python tools/bake_probe.py $C3 --reps 3 --tag "$tag" 2>&1 | tail -1
  env "$@" VLB_BVH_BUILDER=ploc timeout 300 ncu --metrics $M --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__|lts__|l1tex__" | tr -s ' ' | tr '\n' ';'; echo
}
run "persist off" VLB_BAKE_L2_PERSIST=0
run "persist 48 MB scratch" VLB_BAKE_L2_PERSIST=48
run "persist 64 MB scratch" VLB_BAKE_L2_PERSIST=64
run "persist 32 MB scratch (ratio<1)" VLB_BAKE_L2_PERSIST=32
run "persist 32 MB nodes" VLB_BAKE_L2_PERSIST=32 VLB_BAKE_L2_WINDOW=1
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persist max", getattr(p, "persisting_l2_cache_max_size", None), "window max", getattr(p, "access_policy_max_window_size", None))
PY
for v in 0 48; do
VLB_BAKE_L2_PERSIST=$v timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 persist $v" 2>&1 | tail -1 | cut -c1-300
done
