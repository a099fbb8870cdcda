#!/bin/bash
# round 2, GPU call 4 (1 GPU): full GPU test suite + bench N=1 with the comm-capable library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["skybox"]["modes"])
PY
