#!/bin/bash
# compute-sanitizer over the device code this session added: per-direction slots + 16-bit hit queue, deferred shading, direction
# tables, hit-lane table / entry nodes of the visibility batches, cache-peak micro-kernels. Time-boxed.
mkdir -p gpurun_out
T="tests/test_gpu_parity.py tests/test_gpu_gather.py"
timeout 130 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "(room_vs_oracle and base) or odd_direction or gather_pass_device or multibounce" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 110 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "(room_vs_oracle and base) or gather_pass_device" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
timeout 90 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "(room_vs_oracle and base) or odd_direction or gather_pass_device" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/sanitizer_initcheck.log
for f in memcheck racecheck initcheck; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|rc=|RACECHECK SUMMARY" gpurun_out/sanitizer_$f.log | sort | uniq -c | head -8; done
