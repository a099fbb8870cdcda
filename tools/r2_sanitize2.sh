#!/bin/bash
mkdir -p gpurun_out
K="ploc or index_past or room or gather or chained or graph or ptrs or multi or slab or comm or sharded"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py tests/test_gpu_comm.py -m gpu -x -q -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py -m gpu -x -q -k "$K" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_gather.py tests/test_gpu_parity.py -m gpu -x -q -k "room_vs_oracle or gather_pass_device or chained or ploc or index_past" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ploc or room_vs_oracle or index_past" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/sanitizer_initcheck.log
for f in memcheck synccheck racecheck initcheck; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|rc=|RACECHECK SUMMARY" gpurun_out/sanitizer_$f.log | sort | uniq -c | head -8; done
