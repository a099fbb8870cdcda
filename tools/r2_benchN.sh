#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_abi.json 2> gpurun_out/bench_n${N}_abi.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_n${N}_abi.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n${N}_abi.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "bvh_builder")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("host_grid_equals_device_grid"))
PY
