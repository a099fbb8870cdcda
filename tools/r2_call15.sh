#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "bvh_builder")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("bvh", d["bvh"]); print("c2", d["c2"]["value"], d["c2"]["e2e"]["value"], d["c2"]["e2e"]["ms_per_step"])
print("roofline", d["roofline"]["nodes_per_ray"], d["roofline"]["frac"], d["roofline"]["binding_pipe"])
print("skybox", {k: round(v["frac"], 3) for k, v in d["skybox"]["modes"].items()}); print("c4", d["c4"].get("pass_kernel_ms"), "c5 min", d["c5"].get("min_frac"))
PY
