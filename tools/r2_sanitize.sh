#!/bin/bash
# compute-sanitizer over the GPU tests that drive this round's new device code (shared-memory stack / hit queue,
# visibility-ray batches, cell roots, L2 discard, un-interleave, PDL-chained / graph-replayed projections)
mkdir -p gpurun_out
T="tests/test_gpu_gather.py tests/test_gpu_parity.py"
K="room or gather or chained or graph or ptrs or multi or slab"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "$K" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_gather.py tests/test_gpu_parity.py -m gpu -x -q -k "room_vs_oracle or gather_pass_device or chained" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
for f in memcheck synccheck racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|rc=|Race reported|hazard" gpurun_out/sanitizer_$f.log | sort | uniq -c | head -12; done
