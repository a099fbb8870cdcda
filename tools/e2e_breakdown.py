#!/usr/bin/env python
"""Where the e2e step's overhead goes (run under torchrun): every phase of bench.py's e2e step timed on the host with a
synchronisation after it (so the phases do not overlap as they do in the real step), on C3 with the library's communicator."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
vlb = importlib.import_module("vulkan-light-bakery_b200"); scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
scene, sky, s = bench.workload(vlb, scenes, world, "c3")
ctx = vlb.Context(local)
if world > 1:
    uid = torch.zeros(vlb.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(vlb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init_rank(uid.cpu().numpy().tobytes(), rank, world)
    ctx.comm_sharded_uploads(True)
pins = {k: bench.pinned_like(torch, np.ascontiguousarray(scene[k])) for k in ("vertices", "indices", "instances", "materials")}
pscene = {k: v[0] for k, v in pins.items()}
psky, _k = bench.pinned_like(torch, sky)
ctx.set_bvh_builder(vlb.recommend_builder(bench.N_TRIS, s.n_probes * s.dir_w * s.dir_h // world))
grid = torch.empty((s.n_probes, 48), dtype=torch.float32, pin_memory=True)
full = torch.zeros((s.n_probes, 48), device=dev)
def sync():
    ctx.synchronize(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
phases = {}
def timed(name, fn):
    sync(); t0 = time.perf_counter(); fn(); ctx.synchronize(); torch.cuda.synchronize(); phases.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for rep in range(6):
    timed("skybox upload (blocking)", lambda: ctx.set_skybox(psky))
    timed("set_scene", lambda: ctx.set_scene(pscene))
    timed("build_bvh", lambda: ctx.build_bvh())
    timed("bake + gather (device)", lambda: ctx.bake_probes_sharded_device(s, 0, full.data_ptr()))
    timed("bake + gather + own rows to host", lambda: ctx.bake_probes_sharded_rows(s, grid.data_ptr()))
    def step():
        ctx.set_skybox_async(psky); ctx.set_scene(pscene); ctx.build_bvh(); ctx.bake_probes_sharded_rows(s, grid.data_ptr())
    timed("whole e2e step", step)
if rank == 0:
    for k, v in phases.items():
        print("%-36s %8.3f ms (min of %d)" % (k, min(v[1:]), len(v) - 1))
ctx.close()
if world > 1:
    dist.destroy_process_group()
