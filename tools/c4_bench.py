#!/usr/bin/env python
"""BASELINE configs[3] (C4) on one GPU: procedural atrium at ~3 M triangles (seed 11), 32x16x32 probes,
64x64 directions per probe, L3 SH (16 coefficients), direct pass + 3 gather passes
(vlb_bake_gather_device: every hit adds the reference's 8-probe visibility-weighted gather of the pass
before, shaders/main.rchit:124-163). Prints one JSON line. Device-resident, CUDA events of the library."""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vlb = importlib.import_module("vulkan-light-bakery_b200")
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=3 * (1 << 20))
ap.add_argument("--probes", default="32x16x32")
ap.add_argument("--dirs", default="64x64")
ap.add_argument("--bounces", type=int, default=3)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tag", default="")
a = ap.parse_args()
probes = tuple(int(x) for x in a.probes.split("x")); dirs = tuple(int(x) for x in a.dirs.split("x"))
scene = scenes.atrium(a.tris, seed=11); sky = scenes.hdr_sky(2048, 1024, seed=1)
s = scenes.atrium_settings(probes=probes, dirs=dirs, order=a.order, bounds=(0, 0, 0) + tuple(scenes.HALL))
s.indirect_gain = 1.0
ctx = vlb.Context(0)
ctx.set_scene(scene); bvh = ctx.build_bvh(); bvh = ctx.build_bvh(); ctx.set_skybox(sky)
bufs = [torch.zeros((s.n_probes, 48), device="cuda") for _ in range(2)]
best = None
for rep in range(a.reps + 1):
    ms, shadow = [], []
    prev = 0
    for p in range(1 + a.bounces):
        out = bufs[p & 1]
        ctx.bake_gather_device(s, prev, out.data_ptr()); ctx.synchronize()
        st = ctx.last_bake_stats()
        ms.append(st.kernel_ms); shadow.append(int(st.n_shadow_rays))
        prev = out.data_ptr()
    if rep > 0 and (best is None or sum(ms) < sum(best)):
        best = ms
rays = s.n_probes * dirs[0] * dirs[1]
res = {"workload": "C4: atrium seed 11, %d triangles; %s probes x %d rays; L%d SH; direct + %d gather passes" % (a.tris, a.probes, dirs[0] * dirs[1], a.order, a.bounces),
       "tag": a.tag, "lib": os.path.basename(vlb.LIB_PATH), "bvh_build_ms": bvh.build_ms, "bvh_nodes": int(bvh.n_nodes),
       "pass_kernel_ms": [round(x, 3) for x in best], "total_ms": round(sum(best), 3),
       "primary_Grays_per_s_per_pass": [round(rays / x / 1e6, 3) for x in best],
       "primary_Grays_per_s_all_passes": round(rays * len(best) / sum(best) / 1e6, 3),
       "shadow_rays_per_pass": shadow, "checksum": float(bufs[a.bounces & 1].double().abs().sum())}
print(json.dumps(res))
