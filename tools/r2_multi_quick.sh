#!/bin/bash
# round 2, final build on N GPUs: comm tests, one-process-per-GPU check, multi-device CLI, bench at N (all-gather inside the library)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_parity.py -m gpu -x -q -k "comm or ranks or multi" > gpurun_out/pytest_comm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_comm.log
tail -n 3 gpurun_out/pytest_comm.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py > gpurun_out/comm_check_n$N.log 2>&1; echo "comm_check rc=$?"
grep "rank" gpurun_out/comm_check_n$N.log | sort | tail -n 8 | cut -c1-200
timeout 300 bash tools/multi_device_cli_check.sh > gpurun_out/multi_device_cli.log 2>&1; tail -n 4 gpurun_out/multi_device_cli.log
bash tools/r2_benchN.sh $N
