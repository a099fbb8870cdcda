#!/bin/bash
# round 2: the whole record on one GPU -- GPU suite, smoke, both bench arms, ncu launch list of the bench, ncu --set full of the two hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( time timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
VLB_BVH_BUILDER=ploc timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bake_stream -s 2 -c 1 -o gpurun_out/prof_bake_c3_r2 -f python tools/bake_probe.py --probes 64x32x64 --dirs 64x64 --reps 1 > gpurun_out/ncu_bake_c3_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_project -s 6 -c 2 -o gpurun_out/prof_skybox_r2 -f python tools/skybox_bench.py > gpurun_out/ncu_skybox_r2.log 2>&1
timeout 300 bash tools/reference_default_bake.sh > gpurun_out/reference_default_bake.log 2>&1; tail -n 3 gpurun_out/reference_default_bake.log
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -c 400 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "bvh_builder")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "traffic")}, d["roofline"]["hbm"], d["roofline"]["binding_pipe"])
print("skybox", {k: round(v["frac"], 3) for k, v in d["skybox"]["modes"].items()})
print("c2", d.get("c2", {}).get("value"), d.get("c2", {}).get("e2e"))
print("c4", {k: d["c4"].get(k) for k in ("pass_kernel_ms", "value", "error")})
print("c5", [(o["W"], o["sh_order"], o["maps"], round(o["frac"], 3)) for o in d["c5"].get("sizes", [])], d["c5"].get("error"))
print("cpu", d["cpu_baseline"]); print("bvh", d["bvh"])
r = json.loads([l for l in open("gpurun_out/bench_ref.json").read().strip().splitlines() if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"]["cores"], r.get("vulkan_probe_missing"))
PY
