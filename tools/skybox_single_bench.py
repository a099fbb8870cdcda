#!/usr/bin/env python
"""C1 (one 2048x1024 RGBA32F map -> L2 SH per launch), 8 distinct maps rotated (268 MB > L2), on the ctx's OWN stream:
back-to-back vlb_skybox_project_sh_device calls (chained with programmatic dependent launch unless VLB_PROJ_PDL=0),
the multi-pointer entry point (lanes) and the 8 maps as one batched launch. Prints one line per mode."""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vlb = importlib.import_module("vulkan-light-bakery_b200")
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--tag", default="")
ap.add_argument("--wh", default="2048x1024")
ap.add_argument("--order", type=int, default=2)
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
W, H = (int(x) for x in a.wh.split("x"))
n = 8
ctx = vlb.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
maps = torch.rand((n, H, W, 4), device="cuda")
outs = torch.zeros((n, 48), device="cuda")
torch.cuda.synchronize()
stride = H * W * 16
ptrs = [maps[i].data_ptr() for i in range(n)]
peak = 6545.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def single():
    for i in range(n):
        ctx.skybox_project_sh_device(ptrs[i], stride, 1, vlb.FMT_RGBA32F, W, H, a.order, outs[i].data_ptr())
def pipelined():
    ctx.skybox_project_sh_device_ptrs(ptrs, vlb.FMT_RGBA32F, W, H, a.order, outs.data_ptr())
def batched():
    ctx.skybox_project_sh_device(maps.data_ptr(), stride, n, vlb.FMT_RGBA32F, W, H, a.order, outs.data_ptr())

def make_graph(fn):
    """fn's launches captured from the ctx's own stream into a CUDA graph; returns a replay function."""
    from cuda.bindings import runtime as rt
    h = ctx.stream
    fn(); ctx.synchronize()                       # tables, scratch, tensor maps exist before the capture
    (err,) = rt.cudaStreamBeginCapture(h, rt.cudaStreamCaptureMode.cudaStreamCaptureModeThreadLocal); assert err == 0, err
    fn()
    err, graph = rt.cudaStreamEndCapture(h); assert err == 0, err
    err, ge = rt.cudaGraphInstantiate(graph, 0); assert err == 0, err
    def replay():
        (e,) = rt.cudaGraphLaunch(ge, h); assert e == 0, e
    return replay

modes = [("single", single), ("pipelined", pipelined), ("batched8", batched)]
if a.graph:
    modes += [("graph(single x8)", make_graph(single))]
res = {}
for name, fn in modes:
    for _ in range(3):
        fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        fn()
    e1.record(st)
    ctx.synchronize(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (a.reps * n)
    res[name] = (us, stride / us / 1e3, stride / us / 1e3 / peak)
print("%s %s order %d: " % (a.tag, a.wh, a.order) + " | ".join("%s %.2f us/map %.0f GB/s %.3f" % (k, *v) for k, v in res.items()))
if a.cpu:
    # host cost per call: wall clock of issuing 4000 calls (the GPU work of a 64x32 map is negligible), and of a no-op ABI call
    import time
    tiny = torch.rand((32, 64, 4), device="cuda")
    ctx.skybox_project_sh_device(tiny.data_ptr(), 32 * 64 * 16, 1, vlb.FMT_RGBA32F, 64, 32, a.order, outs[0].data_ptr()); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(4000):
        ctx.skybox_project_sh_device(tiny.data_ptr(), 32 * 64 * 16, 1, vlb.FMT_RGBA32F, 64, 32, a.order, outs[0].data_ptr())
    t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
    for _ in range(4000):
        ctx.launch_count
    t3 = time.perf_counter()
    print("   host: %.2f us per projection call issued (%.2f us incl. drain), %.2f us per no-op ABI call" % ((t1 - t0) / 4000 * 1e6, (t2 - t0) / 4000 * 1e6, (t3 - t2) / 4000 * 1e6))
