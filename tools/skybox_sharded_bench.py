#!/usr/bin/env python
"""BASELINE configs[4] across GPUs (launch with torch.distributed.run, one rank per GPU): n maps of one size are dealt
round-robin (parallel.project_maps_sharded), every rank projects its device-resident share in one batched launch, one
NCCL all-gather of 192 bytes per map. Prints one JSON line per size from rank 0: whole-job GB/s, max over ranks of the
CUDA-event time, all-gather included."""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
vlb = importlib.import_module("vulkan-light-bakery_b200")
par = importlib.import_module("vulkan-light-bakery_b200.parallel")
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--gb-per-gpu", type=float, default=4.0, help="resident map bytes per GPU and size")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = vlb.Context(local)
ctx.set_stream(stream.cuda_stream)
gen = torch.Generator(device=dev).manual_seed(1 + rank)
for (W, H) in ((512, 256), (1024, 512), (2048, 1024), (4096, 2048)):
    per_gpu = max(1, int(a.gb_per_gpu * 1e9 // (W * H * 16)))
    n_maps = per_gpu * world
    mine = par.map_share(n_maps, rank, world)
    maps = torch.rand((len(mine), H, W, 4), device=dev, generator=gen, dtype=torch.float32)
    stride = H * W * 16

    def project(ids, out):
        ctx.skybox_project_sh_device(maps.data_ptr(), stride, len(ids), vlb.FMT_RGBA32F, W, H, a.order, out.data_ptr())

    for _ in range(2):
        full = par.project_maps_sharded(project, n_maps, rank, world, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        full = par.project_maps_sharded(project, n_maps, rank, world, device=dev)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = n_maps * stride / (float(ms.item()) * 1e-3) / 1e9
        print(json.dumps({"W": W, "H": H, "maps": n_maps, "order": a.order, "n_gpus": world, "ms": round(float(ms.item()), 3),
                          "GBs_whole_job": round(gbs, 1), "GBs_per_gpu": round(gbs / world, 1), "gathered": list(full.shape)}), flush=True)
    del maps
ctx.close()
if world > 1:
    dist.destroy_process_group()
