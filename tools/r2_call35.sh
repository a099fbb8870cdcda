#!/bin/bash
# per-direction slots (hit record, then radiance) + deferred shading (refill order 3)
C3="--probes 64x32x64 --dirs 64x64"
P=$PWD/vulkan-light-bakery_b200
M=dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
export VLB_BVH_BUILDER=ploc VLB_BAKE_L2_PERSIST=0
run() { tag=$1; shift
  env "$@" timeout 300 python tools/bake_probe.py $C3 --reps 2 --tag "$tag" 2>&1 | tail -1
  env "$@" timeout 300 ncu --metrics $M --clock-control none -k regex:k_bake_stream -s 2 -c 1 python tools/bake_probe.py $C3 --reps 1 2>&1 | grep -E "dram__|l1tex__|smsp__" | tr -s ' ' | tr '\n' ';'; echo
}
run "slots order 1" VLB_BAKE_REFILL_ORDER=1
run "slots order 3" VLB_BAKE_REFILL_ORDER=3
run "slots order 3 chunk 32" VLB_BAKE_REFILL_ORDER=3 VLB_LIB=$P/libvlb_bake_c32.so
run "slots order 1 chunk 32" VLB_BAKE_REFILL_ORDER=1 VLB_LIB=$P/libvlb_bake_c32.so
unset VLB_BVH_BUILDER
for o in 1 3; do
VLB_BAKE_REFILL_ORDER=$o timeout 300 python tools/bake_probe.py --reps 5 --tag "c2 order $o" 2>&1 | tail -1
VLB_BAKE_REFILL_ORDER=$o timeout 600 python tools/c4_bench.py --reps 1 --tag "c4 order $o" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], d['pass_kernel_ms'], d['checksum'])"
VLB_BAKE_REFILL_ORDER=$o timeout 300 bash tools/reference_default_bake.sh 2>&1 | grep baked | tail -1
done
VLB_BAKE_REFILL_ORDER=3 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
