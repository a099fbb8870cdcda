#!/usr/bin/env python
"""Per-rank bake kernel time of the cyclic z-slice shares of C3 (run under torchrun): how even is the deal?"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
par = importlib.import_module("vulkan-light-bakery_b200.parallel")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
scene, sky, s = bench.workload(vlb, scenes, world, "c3")
ctx = vlb.Context(local)
ctx.set_scene(scene); ctx.set_bvh_builder("ploc"); ctx.build_bvh(); ctx.set_skybox(sky)
for mode in ("cyclic", "contiguous"):
    mine = par.shard_settings(s, rank, world, cyclic=(mode == "cyclic"))
    out = torch.zeros((mine.n_slab_probes, 48), device="cuda")
    ms = []
    for _ in range(4):
        ctx.bake_probes_device(mine, out.data_ptr()); ctx.synchronize(); ms.append(ctx.last_bake_stats().kernel_ms)
    st = ctx.last_bake_stats()
    print("rank %d/%d %-10s kernel %.3f ms, %d shadow rays" % (rank, world, mode, min(ms[1:]), st.n_shadow_rays), flush=True)
ctx.close()
