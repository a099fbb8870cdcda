#!/usr/bin/env python
"""Bake kernel micro-benchmark on C2 (or C3 with --c3): device-resident scene, kernel_ms from the library's events."""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vlb = importlib.import_module("vulkan-light-bakery_b200")
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--probes", default="16x8x16")
ap.add_argument("--dirs", default="32x32")
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--tris", type=int, default=262144)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tag", default="")
a = ap.parse_args()
probes = tuple(int(x) for x in a.probes.split("x")); dirs = tuple(int(x) for x in a.dirs.split("x"))
scene = scenes.atrium(a.tris, seed=7); sky = scenes.hdr_sky(2048, 1024, seed=1)
s = scenes.atrium_settings(probes=probes, dirs=dirs, order=a.order, bounds=(0, 0, 0) + tuple(scenes.HALL))
ctx = vlb.Context(0)
ctx.set_scene(scene); ctx.build_bvh(); ctx.set_skybox(sky)
out = torch.zeros((s.n_probes, 48), device="cuda")
ms = []
for i in range(a.reps + 2):
    ctx.bake_probes_device(s, out.data_ptr()); ctx.synchronize()
    ms.append(ctx.last_bake_stats().kernel_ms)
st = ctx.last_bake_stats()
best = min(ms[2:]); rays = st.n_primary_rays
print("%s probes %s dirs %s K%d: kernel %.3f ms (median %.3f) -> %.3f Grays/s primary, %.3f incl shadow; checksum %.6f" % (
    a.tag, a.probes, a.dirs, 9 if a.order == 2 else 16, best, float(np.median(ms[2:])), rays / best / 1e6, (rays + st.n_shadow_rays) / best / 1e6,
    float(out.double().abs().sum())))
if st.n_nodes_visited:
    nr = rays + st.n_shadow_rays
    print("   counters: %.2f nodes/ray, %.2f tris/ray (over primary+shadow rays)" % (st.n_nodes_visited / nr, st.n_tris_tested / nr))
