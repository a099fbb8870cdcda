#!/bin/bash
# round 2 (2+ GPUs): the library's own NCCL communicator: in-process ranks (pytest), one process per GPU (torchrun), bench at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_parity.py -m gpu -x -q -k "comm or ranks or multi" > gpurun_out/pytest_comm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_comm.log
tail -n 5 gpurun_out/pytest_comm.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py > gpurun_out/comm_check_n$N.log 2>&1; echo "comm_check rc=$?"
grep "rank" gpurun_out/comm_check_n$N.log | sort | tail -n 40
timeout 300 bash tools/multi_device_cli_check.sh > gpurun_out/multi_device_cli.log 2>&1; tail -n 4 gpurun_out/multi_device_cli.log
for g in abi torch; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --gather $g > gpurun_out/bench_n${N}_$g.json 2> gpurun_out/bench_n${N}_$g.err; echo "bench $g rc=$?"
  tail -c 300 gpurun_out/bench_n${N}_$g.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n${N}_$g.json").read().strip().splitlines() if l.startswith("{")][-1])
print("$g", {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
PY
done
